"""Compute transport maps between pairs of time points (CLI).

Flags, defaults and their mapping onto OTModel keywords follow the reference verbatim
(wot/commands/util.py:179-237 and :146-176; wot/commands/optimal_transport.py:12-30).  Additive flags:
--kernel {auto,stored,online}, --streams N (day-pairs in flight per GPU), and --format also accepts txt / npz
(usable without anndata / h5py).
Run one process per GPU under torchrun to shard the day-pairs across GPUs.
"""
from __future__ import annotations

import argparse
import logging


def add_ot_parameters_arguments(parser):
    parser.add_argument("--matrix", required=True, help="Gene expression matrix (cells on rows, genes on columns)")
    parser.add_argument("--cell_days", required=True, help='File with headers "id" and "day" (cell id, day)')
    parser.add_argument("--cell_growth_rates", help='File with headers "id" and "cell_growth_rate" (growth per day)')
    parser.add_argument("--parameters", help="Optional two column parameter file containing parameter name and value")
    parser.add_argument("--config", help="Configuration per timepoint or pair of timepoints")
    parser.add_argument("--transpose", action="store_true", help="Transpose the matrix")
    parser.add_argument("--local_pca", type=int, default=30,
                        help="Convert day pairs matrix to local PCA coordinates. Set to 0 to disable")
    parser.add_argument("--growth_iters", type=int, default=1,
                        help="Number of growth iterations for learning the growth rate.")
    parser.add_argument("--gene_filter", help="File with one gene id per line to use for computing cost matrices")
    parser.add_argument("--cell_filter", help="File with one cell id per line to include")
    parser.add_argument("--cell_day_filter", type=str, help="Comma separated list of days to include (e.g. 12,14,16)")
    parser.add_argument("--scaling_iter", type=int, default=3000, help="Number of scaling iterations for OT solver")
    parser.add_argument("--inner_iter_max", type=int, default=50, help="For OT solver")
    parser.add_argument("--epsilon", type=float, default=0.05, help="Controls the entropy of the transport map")
    parser.add_argument("--lambda1", type=float, default=1,
                        help="Regularization parameter that controls the fidelity of the constraints on p")
    parser.add_argument("--lambda2", type=float, default=50,
                        help="Regularization parameter that controls the fidelity of the constraints on q")
    parser.add_argument("--max_iter", type=int, default=1e7,
                        help="Maximum number of scaling iterations. Abort if convergence was not reached")
    parser.add_argument("--batch_size", type=int, default=5,
                        help="Number of scaling iterations to perform between duality gap check")
    # the reference declares type=int here (util.py:220), which rejects every non-integer value; float is what
    # the parameter means
    parser.add_argument("--tolerance", type=float, default=1e-8,
                        help="Maximal acceptable ratio between the duality gap and the primal objective value")
    parser.add_argument("--epsilon0", type=float, default=1, help="Warm starting value for epsilon")
    parser.add_argument("--tau", type=float, default=10000, help="For OT solver")
    parser.add_argument("--ncells", type=int, help="Number of cells to downsample from each timepoint and covariate")
    parser.add_argument("--ncounts", type=int, help="Sample ncounts from each cell")
    parser.add_argument("--solver", choices=["duality_gap", "fixed_iters"], default="duality_gap",
                        help="The solver to use to compute transport matrices")
    parser.add_argument("--cell_days_field", default="day", dest="day_field",
                        help="Field name in cell_days file that contains cell days")
    parser.add_argument("--cell_growth_rates_field", default="cell_growth_rate", dest="growth_rate_field",
                        help="Field name in cell_growth_rates file that contains growth rates")
    parser.add_argument("--verbose", action="store_true", help="Print progress information")


def initialize_ot_model_from_args(args):
    from .. import ot
    extra = {"kernel": args.kernel} if getattr(args, "kernel", None) else {}
    if getattr(args, "streams", None):
        extra["streams"] = args.streams
    return ot.initialize_ot_model(
        args.matrix, cell_days=args.cell_days, solver=args.solver, local_pca=args.local_pca,
        growth_rate_field=args.growth_rate_field, day_field=args.day_field,
        covariate_field=args.covariate_field if hasattr(args, "covariate_field") else None,
        growth_iters=args.growth_iters, epsilon=args.epsilon, lambda1=args.lambda1, lambda2=args.lambda2,
        epsilon0=args.epsilon0, tau=args.tau, config=args.config, parameters=args.parameters,
        cell_day_filter=args.cell_day_filter, cell_growth_rates=args.cell_growth_rates, gene_filter=args.gene_filter,
        cell_filter=args.cell_filter, scaling_iter=args.scaling_iter, inner_iter_max=args.inner_iter_max,
        ncells=args.ncells, ncounts=args.ncounts, transpose=args.transpose, max_iter=args.max_iter,
        batch_size=args.batch_size, tolerance=args.tolerance,
        covariate=args.covariate if hasattr(args, "covariate") else None, **extra)


def create_parser():
    parser = argparse.ArgumentParser(description="Compute transport maps between pairs of time points")
    add_ot_parameters_arguments(parser)
    parser.add_argument("--format", default="h5ad", choices=["h5ad", "loom", "txt", "npz"], help="Output file format")
    parser.add_argument("--no_overwrite", action="store_true",
                        help="Do not overwrite existing transport maps if they exist")
    parser.add_argument("--out", default="./tmaps", help="Prefix for output file names")
    parser.add_argument("--kernel", choices=["auto", "stored", "online"], default=None,
                        help="GPU kernel family: K kept in HBM, or recomputed from coordinates on tcgen05 + MUFU "
                             "(default auto: online unless the final epsilon is below 0.02 or local_pca > 46)")
    parser.add_argument("--streams", type=int, default=None,
                        help="Day-pairs kept in flight per GPU on separate CUDA streams (default 2; 1 = serial loop)")
    return parser


def main(args):
    import os
    if args.verbose:
        logger = logging.getLogger("wot")
        logger.setLevel(logging.DEBUG)
        logger.addHandler(logging.StreamHandler())
    ot_model = initialize_ot_model_from_args(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        from .. import parallel
        if not dist.is_initialized():
            dist.init_process_group("gloo")   # plumbing only: growth tables and a barrier
        parallel.compute_all_transport_maps(ot_model, tmap_out=args.out, overwrite=not args.no_overwrite,
                                            output_file_format=args.format)
    else:
        ot_model.compute_all_transport_maps(overwrite=not args.no_overwrite, output_file_format=args.format,
                                            tmap_out=args.out)
