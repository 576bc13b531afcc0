"""OTModel with the reference's API (reference: wot/ot/ot_model.py:19-326); the cost matrix and the
solver run on the GPU through libwot_b200.so.

Kept verbatim from the reference contract: constructor keywords and defaults (:37-114), day-pair
enumeration, file naming and --no_overwrite semantics of compute_all_transport_maps (:124-201), the
`g0..gN` obs columns (:319-325), the errors raised.  Additive keys in ot_config: kernel='stored'|'online',
out_dtype, device.
"""
from __future__ import annotations

import itertools
import logging
import os

import numpy as np
import pandas as pd
import scipy.sparse

from .. import io as _io
from .._anndata import AnnData
from . import optimal_transport as _ot
from .initializer import parse_configuration, parse_parameter_file
from .util import compute_pca

logger = logging.getLogger("wot")

_DEFAULTS = {"local_pca": 30, "growth_iters": 1, "epsilon": 0.05, "lambda1": 1, "lambda2": 50, "epsilon0": 1,
             "tau": 10000, "scaling_iter": 3000, "inner_iter_max": 50, "tolerance": 1e-8, "max_iter": 1e7,
             "batch_size": 5, "extra_iter": 1000}  # ot_model.py:85-87


def _rows_key(mask):
    """A boolean row mask as a slice when the selected rows are one contiguous run (matrices sorted by day): the
    day's expression rows are then a view, not a 150 MB copy."""
    idx = np.flatnonzero(mask)
    if len(idx) > 0 and idx[-1] - idx[0] + 1 == len(idx):
        return slice(int(idx[0]), int(idx[-1]) + 1)
    return mask


def _dense(x):
    return x.toarray() if scipy.sparse.isspmatrix(x) else np.asarray(x)


class OTModel:
    """Computes transport maps between the time points of an expression matrix.

    Parameters
    ----------
    matrix : AnnData
        Cells x genes expression matrix; obs must hold the day of every cell.
    day_field, covariate_field, growth_rate_field : str
        obs column names.
    **kwargs
        config, cell_filter, gene_filter, cell_day_filter, ncounts, ncells, solver, parameters, and any
        OT parameter (epsilon, lambda1, lambda2, growth_iters, local_pca, ...).
        Additive (not in the reference): `streams` (default 3) = day-pairs compute_all_transport_maps keeps in
        flight on the GPU (wot_b200.pipeline; 1 = the reference's serial loop), `kernel` = 'auto' | 'stored' |
        'online'.
    """

    def __init__(self, matrix, day_field="day", covariate_field="covariate", growth_rate_field="cell_growth_rate",
                 **kwargs):
        self.matrix = matrix
        self.day_field = day_field
        self.covariate_field = covariate_field
        self.cell_growth_rate_field = growth_rate_field
        self.day_pairs = parse_configuration(kwargs.pop("config", None))
        cell_filter = kwargs.pop("cell_filter", None)
        gene_filter = kwargs.pop("gene_filter", None)
        day_filter = kwargs.pop("cell_day_filter", None)
        ncounts = kwargs.pop("ncounts", None)
        ncells = kwargs.pop("ncells", None)
        self.streams = int(kwargs.pop("streams", 3))
        self.matrix = _io.filter_adata(self.matrix, obs_filter=cell_filter, var_filter=gene_filter)
        if day_filter is not None:
            keep_days = [float(t) for t in day_filter.split(",")] if isinstance(day_filter, str) else day_filter
            self.matrix = self.matrix[self.matrix.obs[self.day_field].isin(keep_days).values].copy()
        if ncells is not None:
            self._downsample_cells(ncells)
        if ncounts is not None:
            self._downsample_counts(ncounts)
        if self.matrix.X.shape[0] == 0:
            raise ValueError("No cells in matrix")

        self.ot_config = dict(_DEFAULTS)
        solver = kwargs.pop("solver", "duality_gap")
        if solver == "fixed_iters":
            self.solver = _ot.transport_stablev2
        elif solver == "duality_gap":
            self.solver = _ot.optimal_transport_duality_gap
        else:
            raise ValueError("Unknown solver")
        parameters_file = kwargs.pop("parameters", None)
        self.ot_config.update(kwargs)
        if parameters_file is not None:
            self.ot_config.update(parse_parameter_file(parameters_file))

        local_pca = self.ot_config["local_pca"]
        if local_pca > self.matrix.X.shape[1]:
            logger.warning("local_pca set to {}, above gene count of {}. Disabling PCA".format(
                local_pca, self.matrix.X.shape[1]))
            self.ot_config["local_pca"] = 0
        if self.day_field not in self.matrix.obs.columns:
            raise ValueError("Days information not available for matrix")
        missing_day = self.matrix.obs[self.day_field].isnull()
        if any(missing_day):
            self.matrix = self.matrix[(missing_day == False).values]  # noqa: E712
        self.timepoints = sorted(set(self.matrix.obs[self.day_field]))

    # The reference reads self.timepoints before assigning it (ot_model.py:58 vs :114), so its `ncells`
    # option raises AttributeError; here the option works as documented.
    def _downsample_cells(self, ncells):
        obs = self.matrix.obs
        covs = sorted(set(obs[self.covariate_field])) if self.covariate_field in obs else [None]
        picked = []
        for day in sorted(set(obs[self.day_field].dropna())):
            on_day = (obs[self.day_field] == day).values
            for cv in covs:
                idx = np.where(on_day if cv is None else on_day & (obs[self.covariate_field] == cv).values)[0]
                if len(idx) > ncells:
                    np.random.shuffle(idx)
                    idx = idx[:ncells]
                picked.append(idx)
        self.matrix = self.matrix[np.concatenate(picked)]

    def _downsample_counts(self, ncounts):
        X = self.matrix.X
        for i in range(X.shape[0]):
            row = _dense(X[i]).astype("float64").ravel()
            total = row.sum()
            if total > ncounts:
                X[i] = np.random.multinomial(ncounts, row / total, size=1)[0]

    def get_covariate_pairs(self):
        """All ordered pairs of covariate values present in the dataset."""
        if self.covariate_field not in self.matrix.obs.columns:
            raise ValueError("Covariate value not available in dataset")
        values = set(self.matrix.obs[self.covariate_field])
        return itertools.product(values, values)

    def compute_all_transport_maps(self, tmap_out="tmaps", overwrite=True, output_file_format="h5ad",
                                   with_covariates=False, cost_matrices=None):
        """Compute and save every transport map (ot_model.py:124-201).

        Files are '{dir}/{prefix}_{t0}_{t1}.{fmt}' ('_cv{a}_cv{b}' appended with covariates); existing files
        are skipped when overwrite is False; with growth_iters > 1 the learned growth columns of all pairs go
        to '{prefix}_g.txt'.  Returns None.
        """
        tmap_dir, tmap_prefix = os.path.split(tmap_out) if tmap_out is not None else (None, None)
        tmap_prefix = tmap_prefix or "tmaps"
        tmap_dir = tmap_dir or "."
        os.makedirs(tmap_dir, exist_ok=True)
        day_pairs = self.day_pairs
        if day_pairs is None or len(day_pairs) == 0:
            t = self.timepoints
            day_pairs = [(t[k], t[k + 1]) for k in range(len(t) - 1)]
        if with_covariates:
            day_pairs = [(*pair, cv) for pair, cv in itertools.product(day_pairs, self.get_covariate_pairs())]
        if not day_pairs:
            logger.info("No day pairs")
            return
        if cost_matrices is None:
            cost_matrices = [None] * len(day_pairs)
        growth_frames = []
        keep_growth = self.ot_config.get("growth_iters", 1) > 1
        todo = []
        for day_pair, cost_matrix in zip(day_pairs, cost_matrices):
            if not with_covariates:
                name = tmap_prefix + "_{}_{}".format(*day_pair)
            else:
                name = tmap_prefix + "_{}_{}_cv{}_cv{}".format(day_pair[0], day_pair[1], *day_pair[2])
            output_file = _io.check_file_extension(os.path.join(tmap_dir, name), output_file_format)
            if os.path.exists(output_file) and not overwrite:
                logger.info("Found existing tmap at " + output_file + ". ")
                continue
            todo.append((day_pair, cost_matrix, output_file))

        _io.check_output_format(output_file_format)      # fail before any GPU work, not after the first solve
        # files are written behind the solves: the map sits in a page-locked block that stays alive until its file
        # is complete, the worker goes on to its next day-pair (wot_b200/h5ad.py)
        writer = _io.TmapWriter(output_file_format)

        def one(day_pair, cost_matrix, output_file):
            tmap = self.compute_transport_map(*day_pair, cost_matrix=cost_matrix)
            if tmap is None:
                return None
            writer.write(tmap, output_file)
            return tmap.obs if keep_growth else None

        ours = self.solver in (_ot.optimal_transport_duality_gap, _ot.transport_stablev2)
        if ours and self.streams > 1 and len(todo) > 1:
            # the day-pairs are independent (the reference loops over them serially, :182-199): keep `streams`
            # solves on the SMs and one more context whose coupling is travelling to the host; local PCA of the
            # next pairs runs on the host meanwhile.  Files and the row order of '{prefix}_g.txt' are those of
            # the serial loop.
            from ..pipeline import Pipeline
            counts = self.matrix.obs[self.day_field].value_counts()
            costs = [float(counts.get(job[0][0], 0)) * float(counts.get(job[0][1], 0)) for job in todo]
            try:
                with Pipeline(streams=self.streams + 1, compute_slots=self.streams) as pipe:
                    frames = pipe.map(lambda ctx, job: one(*job), todo, costs=costs)
            except BaseException:
                writer.close(quiet=True)
                raise
        else:
            frames = [one(*job) for job in todo]
        writer.close()
        growth_frames = [f for f in frames if f is not None]
        if growth_frames:
            pd.concat(growth_frames).to_csv(os.path.join(tmap_dir, tmap_prefix + "_g.txt"), sep="\t",
                                            index_label="id")

    def compute_transport_map(self, t0, t1, covariate=None, cost_matrix=None):
        """Transport map from time t0 to time t1 as AnnData (rows: cells at t0, columns: cells at t1).

        Raises ValueError if the model was built with day_pairs and (t0, t1) is not among them.
        """
        if self.day_pairs is not None:
            if (t0, t1) not in self.day_pairs:
                raise ValueError("Transport map ({},{}) is not present in day_pairs".format(t0, t1))
            local_config = self.day_pairs[(t0, t1)]
        else:
            local_config = {}
        if covariate is None:
            logger.info("Computing transport map from {} to {}".format(t0, t1))
        else:
            logger.info("Computing transport map from {} {} to {} {}".format(t0, covariate[0], t1, covariate[1]))
        config = {**self.ot_config, **local_config, "t0": t0, "t1": t1, "covariate": covariate, "C": cost_matrix}
        return self.compute_single_transport_map(config)

    def compute_implicit_transport_map(self, t0, t1, covariate=None):
        """Additive (not in the reference): the transport map from t0 to t1 WITHOUT its I x J matrix, as a
        wot_b200.tmap.ImplicitTransportMap (coordinates, potentials, epsilon) that pushes populations forward and
        pulls them back on the GPU (the products of transport_map_model.py:290, :356).  `obs` carries the same
        g0..gN growth columns as compute_transport_map's result."""
        if self.day_pairs is not None:
            if (t0, t1) not in self.day_pairs:
                raise ValueError("Transport map ({},{}) is not present in day_pairs".format(t0, t1))
            local_config = self.day_pairs[(t0, t1)]
        else:
            local_config = {}
        if self.solver not in (_ot.optimal_transport_duality_gap, _ot.transport_stablev2):
            raise ValueError("implicit transport maps need one of the built-in solvers")
        config = {**self.ot_config, **local_config, "t0": t0, "t1": t1, "covariate": covariate, "C": None,
                  "implicit": True}
        return self.compute_single_transport_map(config)

    @staticmethod
    def compute_default_cost_matrix(a, b, eigenvals=None):
        """Median-normalised squared Euclidean cost (ot_model.py:242-253), computed on the GPU."""
        a, b = _dense(a), _dense(b)
        scale = None
        if eigenvals is not None:
            eigenvals = np.asarray(eigenvals)
            if eigenvals.ndim == 2 and np.count_nonzero(eigenvals - np.diag(np.diagonal(eigenvals))) == 0:
                scale = np.diagonal(eigenvals).astype(np.float64)
            else:  # a general matrix: apply it the way the reference does, then no per-dimension scale
                a, b = a.dot(eigenvals), b.dot(eigenvals)
        return _ot.default_cost_matrix(a, b, scale)

    def compute_single_transport_map(self, config):
        """One transport map from a fully merged config (t0, t1, covariate, C and the OT parameters)."""
        t0 = config.pop("t0", None)
        t1 = config.pop("t1", None)
        if t0 is None or t1 is None:
            raise ValueError("config must have both t0 and t1, indicating target timepoints")
        ds = self.matrix
        covariate = config.pop("covariate", None)
        sel0 = (ds.obs[self.day_field] == float(t0)).values
        sel1 = (ds.obs[self.day_field] == float(t1)).values
        if covariate is not None:
            sel0 = sel0 & (ds.obs[self.covariate_field] == covariate[0]).values
            sel1 = sel1 & (ds.obs[self.covariate_field] == covariate[1]).values
        p0 = ds[_rows_key(sel0), :]
        p1 = ds[_rows_key(sel1), :]
        if p0.shape[0] == 0:
            logger.info("No cells at {}".format(t0))
            return None
        if p1.shape[0] == 0:
            logger.info("No cells at {}".format(t1))
            return None

        local_pca = config.pop("local_pca", None)
        scale = None
        if local_pca is not None and local_pca > 0:
            p0_x, p1_x, pca, _ = compute_pca(p0.X, p1.X, local_pca)
            scale = np.asarray(pca.singular_values_, dtype=np.float64)  # diag of `eigenvals`, ot_model.py:301
        else:
            p0_x, p1_x = _dense(p0.X), _dense(p1.X)

        delta_days = t1 - t0
        if self.cell_growth_rate_field in p0.obs.columns:
            config["G"] = np.power(p0.obs[self.cell_growth_rate_field].values, delta_days)
        else:
            config["G"] = np.ones(p0.shape[0])

        ours = self.solver in (_ot.optimal_transport_duality_gap, _ot.transport_stablev2)
        if config.pop("implicit", False):
            from ..tmap import ImplicitTransportMap
            solver_id = _ot._SOLVER_IDS[self.solver]
            growth_iters = int(config["growth_iters"])
            keys = ("lambda1", "lambda2", "epsilon", "batch_size", "tolerance", "tau", "epsilon0", "max_iter",
                    "scaling_iter", "extra_iter", "inner_iter_max")
            _, learned = _ot.solve_coords(p0_x, p1_x, config["G"], solver_id, scale=scale, growth_iters=growth_iters,
                                          kernel=config.get("kernel", "auto"), want_tmap=False,
                                          **{k: config[k] for k in keys if k in config})
            last = _ot.last_solve_info()
            info = last["infos"][-1]
            obs_growth = {"g" + str(k): np.power(learned[k], 1.0 / delta_days) for k in range(growth_iters + 1)}
            return ImplicitTransportMap(p0_x, p1_x, last["f"], last["g"], last["median"], info["eps_final"],
                                        info["out_scale"], scale=scale,
                                        obs=pd.DataFrame(index=p0.obs.index, data=obs_growth),
                                        var=pd.DataFrame(index=p1.obs.index), t0=t0, t1=t1)
        if config["C"] is None and ours:
            # default cost, growth loop and final row sums in one library call; C never visits the host
            config["coords"] = (p0_x, p1_x, scale)
        elif config["C"] is None:
            config["C"] = OTModel.compute_default_cost_matrix(p0_x, p1_x, None if scale is None else np.diag(scale))
        tmap, learned_growth = _ot.compute_transport_matrix(solver=self.solver, **config)
        if ours:
            last = _ot.last_solve_info()
            final_rows = last["learned_growth"][-1]
            if logger.isEnabledFor(logging.INFO):
                # per-pair metrics: what the reference's progress line lacks
                iters = sum(i["iters"] for i in last["infos"])
                ms = sum(i["gpu_ms"] for i in last["infos"])
                logger.info("tmap {} -> {}: {} x {} cells, {} Sinkhorn iterations in {} solves (final-stage batches {}), "
                            "{:.1f} ms on the GPU, {:.2f} T kernel entries/s".format(
                                t0, t1, tmap.shape[0], tmap.shape[1], iters, len(last["infos"]),
                                [i["batches"][-1] for i in last["infos"]], ms,
                                2.0 * iters * tmap.shape[0] * tmap.shape[1] / max(ms, 1e-9) / 1e9))
        else:
            final_rows = tmap.sum(axis=1)
        learned_growth = list(learned_growth) + [final_rows]
        obs_growth = {"g" + str(k): np.power(g, 1.0 / delta_days) for k, g in enumerate(learned_growth)}
        obs = pd.DataFrame(index=p0.obs.index, data=obs_growth)
        return AnnData(tmap, obs, pd.DataFrame(index=p1.obs.index))
