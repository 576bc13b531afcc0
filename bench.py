#!/usr/bin/env python
"""Benchmark of the transport-map hot path (BASELINE.json metric: day-pair tmaps/s and Sinkhorn
iterations/s at 1/2/4/8 B200, fraction of the HBM or MUFU roofline).

Workload (config.workload): BASELINE.json configs[1], the reprogramming-atlas-shaped
compute_all_transport_maps: 39 day-pairs, 5-20k cells per day (seed 1), 30 local-PCA coordinates,
defaults eps=0.05 lambda1=1 lambda2=50, growth_iters=3.  A STEP is one day-pair transport map:
median-normalised cost + 3 cold-start duality-gap solves + coupling and growth row sums.  Steps walk the
39 pairs in order; with N GPUs rank r takes pair (step*N + r) mod 39 ("independent day-pairs shard one
per GPU", no data-path collective), so per-GPU work is fixed: weak scaling.  Every GPU keeps `--streams` (3)
day-pairs in flight on separate CUDA streams (wot_b200.pipeline; a step is still one day-pair).

  value  tmaps/s with every pair's coordinates already in HBM, timed with CUDA events on the library's
         streams (earliest start to latest end), max over ranks.
  e2e    the same through the host-buffer C-ABI call (wotb_transport_map_from_coords_host): coordinates
         copied from pinned host memory, the float64 coupling, growth rows and potentials copied back; one more
         context than `--streams` so that a coupling crosses PCIe while the solves keep the SMs (compute slots).
  roofline  the dominant kernel of the selected variant, measured live on the mean atlas shape:
            --kernel online (default): k_online_tc, one half-iteration pass that recomputes exp2 of all I*J
            entries from coordinates on tcgen05 + MUFU; bound = MUFU.EX2 throughput (BASELINE.json names the
            MUFU roofline for this variant), algorithmic work = I*J exp evaluations per launch, peak = the
            MUFU.EX2 rate measured on this GPU (wotb_bench_mufu_dev; MEASURED_PEAKS.json has no MUFU figure).
            --kernel stored: k_fused, one Sinkhorn iteration streaming K once; bound = HBM, algorithmic bytes
            = 4*I*ld per launch, peak = MEASURED_PEAKS.json hbm_gbs.  The other variant's figures ride along
            as roofline_stored / roofline_online.
  cpu_baseline  the UNMODIFIED reference solver (oracle/_ref, a byte-for-byte copy of wot/ot/optimal_transport.py made
                by oracle/build_ref.py; the NumPy port in oracle/ if that copy is missing) on a bounded sample.
  c3, c5        BASELINE.json configs[2] (50k x 50k, stored-K vs online-K) and a slice of configs[4] (sweep), N = 1.
  e2e_api       the same workload through the PUBLIC model API from expression matrices:
                OTModel(adata, growth_iters=3).compute_all_transport_maps -> local PCA on the GPU, cost, solves,
                .h5ad files (first --api-pairs day-pairs).
  row_sharded   N > 1: BASELINE.json configs[3], one 100k x 100k pair with its rows sharded over the N GPUs; the ranks
                exchange from inside the pass kernels over peer memory (headline) or with one NCCL all-reduce per Sinkhorn
                iteration (timed beside it); timings, parity checks against the one-GPU solve and float64 marginals.

`--impl reference` times the reference's own CPU implementation of the path on the host cores: cost
(sklearn pairwise_distances + np.median, ot_model.py:249-252) + compute_transport_matrix(optimal_transport_duality_gap,
growth_iters=3) on atlas pairs AT FULL SIZE (the smallest one, and at N = 1 also the median one); the remaining pairs
are extrapolated in proportion to I*J and the line says so.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULTS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7,
                batch_size=5, scaling_iter=3000, extra_iter=1000, inner_iter_max=50)
GROWTH_ITERS = 3
D = 30
WORKLOAD = "atlas-shaped compute_all_transport_maps: 39 day-pairs, 5-20k cells/day, d=30, growth_iters=3"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=39)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every day by this factor (debugging)")
    ap.add_argument("--kernel", default="online", choices=["stored", "online", "online_simt"])
    ap.add_argument("--cpu-cells", type=int, default=3200, help="cells/day of the bounded CPU sample (~10 s)")
    ap.add_argument("--api-pairs", type=int, default=6, help="day-pairs of the e2e_api leg (public OTModel API + .h5ad files)")
    ap.add_argument("--no-extras", action="store_true", help="skip the c3 / c5 / e2e_api / row_sharded blocks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=None,
                    help="day-pairs in flight per GPU, each on its own CUDA stream (wot_b200.pipeline)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU path
# ------------------------------------------------------------------------------------------------
def reference_module():
    """(module, kind): the unmodified reference solver module from oracle/_ref ('reference'), else the NumPy port of
    oracle/ ('port').  On the build container the copy is refreshed from /root/reference first."""
    from oracle import build_ref
    try:
        build_ref.build()
    except Exception:
        pass
    mod = build_ref.load()
    if mod is not None:
        return mod, "reference"
    from oracle import wot_oracle
    return wot_oracle, "port"


def host_threads():
    """Give BLAS every host core (torchrun exports OMP_NUM_THREADS=1, which would silently turn an N > 1 reference
    arm into a single-threaded one) and return the number in use."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return n


def reference_tmap(mod, kind, x0, x1, growth, growth_iters):
    """One transport map the way the reference computes it: ot_model.py:249-252 (cost) + :318 (growth loop)."""
    import sklearn.metrics
    t0 = time.perf_counter()
    cost = sklearn.metrics.pairwise.pairwise_distances(x0, x1, metric="sqeuclidean", n_jobs=-1)
    cost = cost / np.median(cost)
    params = dict(DEFAULTS, growth_iters=growth_iters, C=cost, G=growth.copy())
    if kind == "port":
        params["gap"] = "dense"
    tmap, _ = mod.compute_transport_matrix(mod.optimal_transport_duality_gap, **params)
    return time.perf_counter() - t0, float(tmap.sum())


def cpu_reference_sample(pair, cells, growth_iters=1):
    """Full solve of `pair` subsampled to `cells` cells/day.  Returns (seconds, I, J, kind)."""
    from wot_b200 import synthetic
    mod, kind = reference_module()
    n0, n1, seed = pair
    s = min(1.0, cells / max(n0, n1))
    m0, m1 = max(2, int(n0 * s)), max(2, int(n1 * s))
    x0, x1, growth = synthetic.day_pair_coords(m0, m1, d=D, seed=seed)
    sec, _ = reference_tmap(mod, kind, x0, x1, growth, growth_iters)
    return sec, m0, m1, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores.  Rank 0 alone works."""
    if rank != 0:
        return
    from wot_b200 import synthetic
    cores = host_threads()
    mod, kind = reference_module()
    pairs = synthetic.atlas_pairs(seed=1, scale=args.scale)
    order = sorted(range(len(pairs)), key=lambda k: pairs[k][0] * pairs[k][1])
    timed = [order[0]] + ([order[len(order) // 2]] if args.gpus == 1 and args.steps > 1 else [])
    secs, entries = [], []
    for k in timed:
        n0, n1, seed = pairs[k]
        x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=D, seed=seed)
        sec, _ = reference_tmap(mod, kind, x0, x1, growth, GROWTH_ITERS)
        secs.append(sec)
        entries.append(n0 * n1)
    # the 39-pair job: measured pairs as measured, every other pair at the measured seconds per matrix entry
    per_entry = sum(secs) / sum(entries)
    total = sum(secs[timed.index(k)] if k in timed else per_entry * pairs[k][0] * pairs[k][1] for k in range(len(pairs)))
    value = len(pairs) / total
    sample = ("%d atlas day-pair(s) at FULL size through the %s (%s) with growth_iters=3: %s; the other %d pairs of the "
              "39-pair job are extrapolated at the measured %.3g us per matrix entry (time is proportional to I*J, "
              "SURVEY.md section 6); BLAS threads %d, sklearn pairwise_distances n_jobs=-1"
              % (len(timed), "unmodified reference solver" if kind == "reference" else "NumPy port of the reference",
                 "oracle/_ref" if kind == "reference" else "oracle/",
                 ", ".join("%dx%d in %.1f s" % (pairs[k][0], pairs[k][1], t) for k, t in zip(timed, secs)),
                 len(pairs) - len(timed), 1e6 * per_entry, cores))
    line = {
        "impl": "reference", "metric": "day-pair transport maps per second", "value": value, "unit": "tmaps/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": 0, "ms_per_step": 1e3 * sum(secs) / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config("cpu", 1, args.scale),
        "cpu_baseline": {"value": value, "unit": "tmaps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "tmaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "measured_pairs": [{"shape": [pairs[k][0], pairs[k][1]], "seconds": t} for k, t in zip(timed, secs)],
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(kernel, streams, scale):
    """The `config` object: identical keys in both arms; the values that describe the GPU launch mode are null on the
    CPU arm."""
    gpu = kernel != "cpu"
    online = gpu and kernel != "stored"
    return {
        "workload": WORKLOAD, "solver": "duality_gap", "eps": 0.05, "lambda1": 1, "lambda2": 50, "growth_iters": GROWTH_ITERS,
        "scale": scale, "kernel": kernel if gpu else None, "streams": streams if gpu else None,
        "l2": (("every pass recomputes I*J = 25M-400M entries from L2-resident operands; nothing is cached between steps "
                "(each step is a different day-pair)") if online else
               "inputs larger than L2 (K and C are 0.1-1.6 GB per pair)") if gpu else None,
        "sharding": "one day-pair per GPU per step, no collective",
        "streams_note": ("day-pairs are independent: each GPU keeps `streams` of them in flight on separate CUDA streams "
                         "(wot_b200.pipeline) so one fills the other's kernel tails, checks and copies; a step is still "
                         "one day-pair") if gpu else None,
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class DevicePair:
    """One day-pair on the device, driven through the device-pointer C ABI."""

    def __init__(self, ctx, torch, max_i, max_j, kernel="stored"):
        from wot_b200 import _lib
        self.ctx, self.torch, self.lib, self._lib = ctx, torch, ctx.lib, _lib
        self.kernel = kernel
        dev = "cuda:%d" % ctx.device
        self.ld_max = (max_j + 31) // 32 * 32
        self.C = torch.empty(max_i * self.ld_max if kernel == "stored" else 32, dtype=torch.float32, device=dev)
        self.out = torch.empty(max_i * max_j, dtype=torch.float64, device=dev)
        self.G = torch.empty(max_i, dtype=torch.float64, device=dev)
        self.f = torch.empty(max_i, dtype=torch.float64, device=dev)
        self.g = torch.empty(max_j, dtype=torch.float64, device=dev)
        self.rows = torch.empty(max_i, dtype=torch.float64, device=dev)

    def attach(self, pair, x0_dev, x1_dev, growth_dev):
        """Point at one day-pair whose coordinates already sit in HBM."""
        self.I, self.J = pair[0], pair[1]
        self.x0, self.x1, self.G0 = x0_dev, x1_dev, growth_dev

    def run(self, prm):
        """cost + median + growth loop + coupling, inputs and outputs in HBM.  Returns per-solve infos."""
        lib, h, L = self.lib, self.ctx.handle, self._lib
        I, J = self.I, self.J
        ld = (J + 31) // 32 * 32
        med = C.c_double()
        P = lambda tns: C.c_void_p(tns.data_ptr())  # noqa: E731
        L.check(lib.wotb_cost_median_dev(h, P(self.x0), I, P(self.x1), J, D, None, C.byref(med)))
        online = self.kernel != "stored"
        if not online:
            L.check(lib.wotb_cost_matrix_dev(h, P(self.x0), I, P(self.x1), J, D, None, med.value, P(self.C), ld, L.F32))
        self.G[:I].copy_(self.G0)
        infos = []
        for it in range(GROWTH_ITERS):
            if it > 0:
                self.G[:I].copy_(self.rows[:I])
            info = L.Info()
            if online:
                L.check(lib.wotb_sinkhorn_online_dev(h, P(self.x0), I, P(self.x1), J, D, med.value, P(self.G),
                                                     C.byref(prm), P(self.f), P(self.g), P(self.rows), C.byref(info)))
            else:
                L.check(lib.wotb_sinkhorn_stored_dev(h, P(self.C), ld, I, J, P(self.G), C.byref(prm), P(self.f),
                                                     P(self.g), P(self.rows), C.byref(info)))
            infos.append(info.as_dict())
        last = infos[-1]
        if online:
            L.check(lib.wotb_coupling_online_dev(h, P(self.x0), I, P(self.x1), J, D, med.value, P(self.f), P(self.g),
                                                 last["eps_final"], last["out_scale"], P(self.out), J, L.F64, None))
        else:
            L.check(lib.wotb_coupling_dev(h, P(self.C), ld, I, J, P(self.f), P(self.g), last["eps_final"],
                                          last["out_scale"], P(self.out), J, L.F64, None))
        return infos


def matvec_bytes(info, I, J, fused):
    """Bytes of K the solve had to stream: one sweep per iteration when fused (two otherwise), plus one
    sweep per duality-gap check and one for the final row sums."""
    ld = (J + 31) // 32 * 32
    n = (1 if fused else 2) * info["iters"] + info["batches"][5] + 1
    return n * I * ld * 4, n


def isolated_matvec(ctx, torch, I, J, reps=20):
    """Average duration of one k_row and one k_col launch on an I x J kernel matrix (CUDA events on the
    library's stream inside wotb_bench_matvec_dev)."""
    ms_row, ms_col, ms_fused = C.c_double(), C.c_double(), C.c_double()
    from wot_b200 import _lib
    _lib.check(ctx.lib.wotb_bench_matvec_dev(ctx.handle, I, J, reps, C.byref(ms_row), C.byref(ms_col),
                                             C.byref(ms_fused)))
    return ms_row.value, ms_col.value, ms_fused.value


def online_pass_roofline(ctx, torch, I, J, reps=20):
    """One k_online_tc launch (a half-iteration: exp2 of all I*J entries recomputed from coordinates, row sums
    reduced) on a Sinkhorn-like state, timed with CUDA events on the library's stream inside
    wotb_online_rowsums_dev; the MUFU.EX2 peak is measured on the same GPU by wotb_bench_mufu_dev."""
    from tools.online_pass_check import make_inputs
    from wot_b200 import _lib
    x0, x1, scale, off_out, off_in = make_inputs(I, J, D, seed=6)
    dev = "cuda:%d" % ctx.device
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, x1, off_out, off_in)]
    sums = torch.empty(I, dtype=torch.float64, device=dev)
    ms, peak = C.c_double(), C.c_double()
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.wotb_online_rowsums_dev(ctx.handle, P(t[0]), I, P(t[1]), J, D, float(scale), P(t[2]), P(t[3]),
                                               2, reps, P(sums), C.byref(ms)))
    _lib.check(ctx.lib.wotb_bench_mufu_dev(ctx.handle, C.byref(peak)))
    achieved = I * J / (ms.value * 1e-3) / 1e12
    traffic, traffic_src = measured_traffic("k_online_tc", 4937728, "ncu --set full, profiles/r1t_k_online_tc_ncu_full.txt: dram "
                                            "bytes per launch at 12486x12405 (operands only; nothing of size I*J exists)")
    return {
        "traffic": traffic, "traffic_source": traffic_src,
        "bound": "mufu", "achieved": achieved, "peak": peak.value / 1e12, "unit": "Texp/s",
        "frac": achieved / (peak.value / 1e12),
        "peak_source": "MUFU.EX2 rate measured on this GPU (wotb_bench_mufu_dev: 16 independent ex2 chains per thread, "
                       "32 warps per SM); nominal 16/clk/SM x 148 x 1.965 GHz = 4.65 T/s",
        "kernel": "k_online_tc (tcgen05 cross term + offsets in TMEM, exp2 split between MUFU.EX2 and packed FMA-pipe "
                  "polynomial, one half-iteration per launch)",
        "shape": [I, J], "algorithmic_exp_per_launch": I * J, "pass_ms": ms.value,
        "algorithmic_operand_bytes_per_launch": (((I + 255) // 256 * 256) + ((J + 255) // 256 * 256)) * 192,
        "note": "achieved counts exponentials evaluated per second; entries moved to the FMA pipe make fractions above "
                "1.0 possible in principle",
    }


def measured_traffic(key, fallback_bytes, fallback_source):
    """DRAM bytes per launch of the dominant kernel from the round's `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/summarize_ncu.py traffic); a profiler cannot run inside the timed
    region, so the figure is per round, not per run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            rec = json.load(fh)[key]
        return rec["dram_bytes_per_launch"], rec["source"]
    except Exception:
        return fallback_bytes, fallback_source


def guarded(fn):
    """The extra blocks must never take the headline line down with them."""
    try:
        return fn()
    except Exception as exc:  # noqa: BLE001
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def c3_block(torch, local_rank, mufu_peak, hbm_peak, n=50000):
    """BASELINE.json configs[2]: one 50k x 50k pair (seed 2), stored-K vs online-K on one GPU, defaults."""
    from wot_b200 import _lib, synthetic
    x0, x1, growth = synthetic.day_pair_coords(n, n, d=D, seed=2)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream()
    ctx = _lib.Context(local_rank, stream.cuda_stream)
    lib, h = ctx.lib, ctx.handle
    out = {"shape": [n, n]}
    with torch.cuda.stream(stream):
        X0, X1, G = (torch.from_numpy(a).to(dev) for a in (x0, x1, growth))
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        med = C.c_double()
        t0 = time.perf_counter()
        _lib.check(lib.wotb_cost_median_dev(h, P(X0), n, P(X1), n, D, None, C.byref(med)))
        out["median_s"] = time.perf_counter() - t0
        res = {}
        for kernel in ("online", "stored"):
            prm = _lib.make_params(solver=_lib.SOLVER_DUALITY_GAP, kernel=_lib.KERNEL_STORED if kernel == "stored" else
                                   _lib.KERNEL_ONLINE, **DEFAULTS)
            f = torch.empty(n, dtype=torch.float64, device=dev)
            g = torch.empty(n, dtype=torch.float64, device=dev)
            rows = torch.empty(n, dtype=torch.float64, device=dev)
            info = _lib.Info()
            if kernel == "stored":
                ld = (n + 31) // 32 * 32
                Cm = torch.empty(n * ld, dtype=torch.float32, device=dev)
                _lib.check(lib.wotb_cost_matrix_dev(h, P(X0), n, P(X1), n, D, None, med.value, P(Cm), ld, _lib.F32))
                _lib.check(lib.wotb_sinkhorn_stored_dev(h, P(Cm), ld, n, n, P(G), C.byref(prm), P(f), P(g), P(rows),
                                                        C.byref(info)))
                del Cm
            else:
                _lib.check(lib.wotb_sinkhorn_online_dev(h, P(X0), n, P(X1), n, D, med.value, P(G), C.byref(prm), P(f),
                                                        P(g), P(rows), C.byref(info)))
            i = info.as_dict()
            sec = i["gpu_ms"] * 1e-3
            rec = {"iters": i["iters"], "batches": i["batches"], "solve_ms": i["gpu_ms"], "iters_per_s": i["iters"] / sec,
                   "workspace_gb": lib.wotb_workspace_bytes(h) / 1e9}
            if kernel == "stored":
                ld = (n + 31) // 32 * 32
                # K is read once per iteration: one CTA per row up to 23,040 columns (k_fused), a thread-block cluster
                # per row beyond (k_fused_cl, csrc/fused_cluster.cuh; round 1 / early round 2 needed two sweeps here)
                sweeps = 1
                gbs = sweeps * n * ld * 4 * i["iters"] / sec / 1e9
                rec.update(hbm_gbs_all_in=gbs, hbm_frac_all_in=gbs / hbm_peak, k_sweeps_per_iter=sweeps,
                           kernel="k_fused" if n <= 23040 else "k_fused_cl (cluster of 2-8 CTAs per row)")
            else:
                texp = 2.0 * n * n * i["iters"] / sec / 1e12
                rec.update(texp_per_s_all_in=texp, mufu_frac_all_in=texp / mufu_peak)
            out[kernel] = rec
            res[kernel] = (f.cpu().numpy(), g.cpu().numpy(), rows.cpu().numpy())
            lib.wotb_release_workspace(h)
            torch.cuda.empty_cache()
    a, b = res["stored"], res["online"]
    out["online_vs_stored"] = {"max_abs_df_over_eps": float(np.max(np.abs(a[0] - b[0])) / 0.05),
                               "max_abs_dg_over_eps": float(np.max(np.abs(a[1] - b[1])) / 0.05),
                               "max_rel_rowsum": float(np.max(np.abs(a[2] - b[2]) / np.abs(a[2]))),
                               "same_batches": out["stored"]["batches"] == out["online"]["batches"]}
    out["speedup_online_over_stored"] = out["stored"]["solve_ms"] / out["online"]["solve_ms"]
    ctx.close()
    return out


def c5_block(local_rank, n=10000):
    """A slice of BASELINE.json configs[4]: 8 of the 64 (eps, lambda1, lambda2) settings on the 10k x 10k pair
    (seed 4), two settings in flight (wot_b200.parallel.parameter_sweep)."""
    from wot_b200 import parallel, synthetic
    x0, x1, growth = synthetic.day_pair_coords(n, n, d=D, seed=4)
    grid = [dict(epsilon=e, lambda1=l1, lambda2=l2) for e in (0.01, 0.025, 0.05, 0.1) for l1, l2 in ((1.0, 50.0), (10.0, 10.0))]
    common = {k: v for k, v in DEFAULTS.items() if k not in ("epsilon", "lambda1", "lambda2")}
    t0 = time.perf_counter()
    res = parallel.parameter_sweep(x0, x1, growth, grid, kernel="auto", streams=2, **common)
    wall = time.perf_counter() - t0
    iters = sum(r["iters"] for r in res)
    return {"shape": [n, n], "settings": len(grid), "of": 64, "wall_s": wall, "settings_per_s": len(grid) / wall,
            "sinkhorn_iters": iters, "iters_per_s": iters / wall,
            "iters_by_setting": [[r["setting"]["epsilon"], r["setting"]["lambda1"], r["setting"]["lambda2"], r["iters"]] for r in res],
            "all_converged": all(r["status"] == 0 for r in res)}


def e2e_api_block(local_rank, n_pairs, scale):
    """The workload through the PUBLIC API from expression matrices (1,479 genes, as Notebook 2):
    OTModel(adata, growth_iters=3).compute_all_transport_maps(..., output_file_format='h5ad') = local PCA on the GPU
    + cost + 3 solves + float64 coupling to the host + the .h5ad file of every pair (written behind the solves)."""
    import shutil
    import tempfile

    import pandas as pd
    from wot_b200 import h5ad, ot, synthetic
    from wot_b200._anndata import AnnData
    sizes = [max(2, int(v * scale)) for v in synthetic.atlas_day_sizes(seed=1)[: n_pairs + 1]]
    X, day, growth = synthetic.expression_matrix(sizes, n_genes=1479, seed=1)
    obs = pd.DataFrame({"day": day * 0.5, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = ot.OTModel(adata, growth_iters=GROWTH_ITERS)
    tmp = tempfile.mkdtemp(prefix="wotb_bench_")
    try:
        walls = []
        for rep in range(2):                     # the first pass sizes workspaces and the page-locked output pool
            t0 = time.perf_counter()
            model.compute_all_transport_maps(tmap_out=os.path.join(tmp, "tmaps"), output_file_format="h5ad")
            walls.append(time.perf_counter() - t0)
        files = sorted(f for f in os.listdir(tmp) if f.endswith(".h5ad"))
        nbytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in files)
        back = h5ad.read_h5ad(os.path.join(tmp, files[0]), with_x=False)
        ok = len(files) == n_pairs and list(back["obs"]) == ["g%d" % k for k in range(GROWTH_ITERS + 1)]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"value": n_pairs / walls[1], "unit": "tmaps/s", "pairs": n_pairs, "cells": int(X.shape[0]), "genes": 1479,
            "wall_s": walls[1], "first_pass_wall_s": walls[0], "files": len(files), "h5ad_bytes": nbytes,
            "h2d_bytes": int(2 * X.nbytes - X[: sizes[0]].nbytes - X[-sizes[-1]:].nbytes),
            "files_ok": bool(ok), "includes": "local PCA on the GPU, cost + median, 3 solves per pair, float64 coupling to "
                                              "pinned host memory, .h5ad files written behind the solves"}


def row_sharded_block(torch, dist, rank, world, local_rank, mufu_peak, n=100000):
    """BASELINE.json configs[3]: ONE 100k x 100k pair (seed 3), online kernel, rows sharded over the ranks; per Sinkhorn
    iteration the ranks exchange over peer memory from inside the pass kernels (headline) or with one NCCL all-reduce
    (`nccl_exchange`, timed beside it).  Parity: the same solve on one GPU (every rank runs it redundantly), float64
    blockwise marginals of the returned potentials on sampled rows, the fixed-point equations."""
    from wot_b200 import _lib, parallel, synthetic
    x0, x1, growth = synthetic.day_pair_coords(n, n, d=D, seed=3)
    dev = torch.device("cuda", local_rank)
    def timed(exchange, median=None):
        kw = dict(DEFAULTS, exchange=exchange)
        if median is not None:
            kw["median"] = median
        parallel.sharded_online_solve(x0, x1, growth, **kw)               # warm-up: workspaces, NCCL channels, mappings
        best = None
        all_ms = []
        for _ in range(2):            # two timed solves, the faster one counts (single solves scatter by ~10 %, one 4-GPU
            tm = {}                   # run of the NCCL exchange took 2.6x its usual time); every time is listed
            r = parallel.sharded_online_solve(x0, x1, growth, timers=tm, **kw)
            t = torch.tensor([r["info"]["gpu_ms"], r["info"].get("graph_capture_ms", 0.0)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tm["graph_capture_ms"] = float(t[1].item())
            all_ms.append(float(t[0].item()))
            if best is None or all_ms[-1] < best[2]:
                best = (r, tm, all_ms[-1])
        best[1]["all_solve_ms"] = all_ms
        return best

    # the NCCL exchange (one all-reduce per iteration inside the graph), then the peer-memory exchange (no collective:
    # NVLink stores from the passes' finishing code); the block's headline is the one `exchange="auto"` selects
    res_n, timers_n, ms_nccl = timed("nccl")
    res, timers, solve_ms = timed("auto", median=res_n["median"])
    info = res["info"]
    info["median_ms"] = res_n["info"].get("median_ms")
    f, g, rowsum = res["f"], res["g"], res["rowsum"]
    # ---- the same pair on ONE GPU (redundantly on every rank): batch counts must be identical ----
    ctx = res["ctx"]
    X0, X1 = res["coords"]
    Gd = torch.from_numpy(growth).to(dev)
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    prm = _lib.make_params(solver=_lib.SOLVER_DUALITY_GAP, kernel=_lib.KERNEL_ONLINE, **DEFAULTS)
    f1 = torch.empty(n, dtype=torch.float64, device=dev)
    g1 = torch.empty(n, dtype=torch.float64, device=dev)
    r1 = torch.empty(n, dtype=torch.float64, device=dev)
    one = _lib.Info()
    with torch.cuda.stream(res["stream"]):
        _lib.check(ctx.lib.wotb_sinkhorn_online_dev(ctx.handle, P(X0), n, P(X1), n, D, res["median"], P(Gd), C.byref(prm),
                                                    P(f1), P(g1), P(r1), C.byref(one)))
    one = one.as_dict()
    torch.cuda.synchronize()
    # ---- float64 blockwise check on sampled rows (torch, independent of the library's kernels) ----
    rng = np.random.default_rng(5)
    rows = torch.from_numpy(np.sort(rng.choice(n, 512, replace=False))).to(dev)
    eps = info["eps_final"]
    cst = ((X0[rows][:, None, :] - X1[None, :, :]) ** 2).sum(-1) / res["median"] if n <= 20000 else None
    if cst is None:
        acc = torch.zeros(len(rows), dtype=torch.float64, device=dev)
        for c0 in range(0, n, 8192):
            blk = X1[c0:c0 + 8192]
            cc = ((X0[rows][:, None, :] - blk[None, :, :]) ** 2).sum(-1) / res["median"]
            acc += torch.exp((f[rows][:, None] + g[c0:c0 + 8192][None, :] - cc) / eps).sum(1)
        marg = acc * info["out_scale"]
    else:
        marg = torch.exp((f[rows][:, None] + g[None, :] - cst) / eps).sum(1) * info["out_scale"]
    rel_rowsum = float((torch.abs(rowsum[rows] - marg) / marg).max().item())
    # unbalanced fixed point (optimal_transport.py:133): row mass = p_i exp(-f_i / lambda1)
    Gs = Gd[rows]
    fixed = float((torch.abs(marg * n / n - Gs * torch.exp(-f[rows] / DEFAULTS["lambda1"])) / marg).max().item())
    iters = info["iters"]
    sec = solve_ms * 1e-3
    out = {"shape": [n, n], "n_gpus": world, "iters": iters, "batches": info["batches"], "solve_ms": solve_ms,
           "median_ms": info.get("median_ms"), "median_note": "exact np.median of the 1e10 distances, one pass split over the "
                                                              "ranks' row shards (counts all-reduced, window keys all-gathered)",
           "iters_per_s": iters / sec, "mufu_frac_aggregate": 2.0 * n * n * iters / sec / 1e12 / (world * mufu_peak),
           "exchange": timers.get("exchange"), "launch_mode": timers.get("mode"),
           "graph_capture_ms": timers.get("graph_capture_ms"), "all_solve_ms": timers.get("all_solve_ms"),
           "solve_ms_note": "the faster of two solves (all_solve_ms); CUDA events from the first launch to the last result on the "
                            "solve's stream, max over ranks; includes "
                            "graph_capture_ms of host-side stream capture + graph instantiation per solve (GPU idle)",
           "peer_bytes_out_per_iter": timers.get("peer_bytes_out_per_iter"),
           "nccl_exchange": {"solve_ms": ms_nccl, "iters": res_n["info"]["iters"], "batches": res_n["info"]["batches"],
                             "iters_per_s": res_n["info"]["iters"] / (ms_nccl * 1e-3),
                             "mufu_frac_aggregate": 2.0 * n * n * res_n["info"]["iters"] / (ms_nccl * 1e-3) / 1e12 / (world * mufu_peak),
                             "graph_capture_ms": timers_n.get("graph_capture_ms"), "all_solve_ms": timers_n.get("all_solve_ms"),
                             "allreduce_us_per_iter": timers_n.get("allreduce_us_per_iter"),
                             "allreduce_bytes": timers_n.get("allreduce_bytes"), "launch_mode": timers_n.get("mode"),
                             "max_abs_df_over_eps_vs_headline": float(torch.abs(res_n["f"] - f).max().item() / info["eps_final"])},
           "one_gpu": {"iters": one["iters"], "batches": one["batches"], "solve_ms": one["gpu_ms"]},
           "speedup_vs_one_gpu": one["gpu_ms"] / solve_ms, "efficiency_vs_one_gpu": one["gpu_ms"] / solve_ms / world,
           "parity": {"batches_equal_one_gpu": info["batches"] == one["batches"], "iters_equal_one_gpu": iters == one["iters"],
                      "max_abs_df_over_eps_vs_one_gpu": float(torch.abs(f - f1).max().item() / eps),
                      "max_abs_dg_over_eps_vs_one_gpu": float(torch.abs(g - g1).max().item() / eps),
                      "max_rel_rowsum_vs_float64_marginals_512_rows": rel_rowsum,
                      "max_rel_fixed_point_residual_512_rows": fixed,
                      "tolerance": 1e-4}}
    out["parity"]["ok"] = bool(out["parity"]["batches_equal_one_gpu"] and rel_rowsum <= 1e-4 and
                               out["parity"]["max_abs_df_over_eps_vs_one_gpu"] <= 1e-4)
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from wot_b200 import _lib, _pinned, synthetic
    from wot_b200.ot import optimal_transport as wot_ot

    torch.cuda.set_device(local_rank)
    dist = None
    bound_cpus = 0
    if world > 1:
        # one process per GPU: keep this rank's threads and page-locked buffers on the GPU's NUMA node (at N = 1 the
        # cpu_baseline leg wants every host core, so the process stays unbound)
        from wot_b200.parallel import bind_host_to_gpu
        bound_cpus = bind_host_to_gpu(local_rank)
        import torch.distributed as dist
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (set on the GPU boxes) prints a banner there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from wot_b200.pipeline import Pipeline
    # measured on B200 (profiles/r3e_streams.txt): online solves in flight 2 -> 8.19-8.33, 3 -> 8.53, 4 -> 8.57 tmaps/s;
    # the stored kernel's cooperative launches do not interleave (round 1: 4.82 vs 5.19), so it runs one at a time
    n_streams = max(1, args.streams) if args.streams else (3 if args.kernel != "stored" else 1)
    tstreams = [torch.cuda.Stream() for _ in range(n_streams)]
    pipe = Pipeline(local_rank, n_streams, make_stream=lambda k: tstreams[k].cuda_stream)
    ctx = pipe.contexts[0]
    _lib._contexts[local_rank] = ctx
    online = args.kernel != "stored"
    prm = _lib.make_params(solver=_lib.SOLVER_DUALITY_GAP, kernel=_lib.KERNEL_ONLINE if online else _lib.KERNEL_STORED,
                           online_simt=args.kernel == "online_simt", **DEFAULTS)

    pairs = synthetic.atlas_pairs(seed=1, scale=args.scale)
    n_pairs = len(pairs)
    max_i = max(p[0] for p in pairs)
    max_j = max(p[1] for p in pairs)
    total = args.warmup + args.steps
    mine = [pairs[(s * world + rank) % n_pairs] for s in range(total)]
    coords = {}
    for p in set(mine):
        coords[p] = synthetic.day_pair_coords(p[0], p[1], d=D, seed=p[2])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- leg 1: inputs resident in HBM, CUDA events on the library's streams ---------------------
    # Every step's coordinates are placed in HBM before the timed region.  Worker k (one library context and
    # CUDA stream each) runs steps k, k + streams, ...; the timed region spans from the earliest start event
    # to the latest end event over all streams.
    dev = "cuda:%d" % local_rank
    resident = {}
    for p in set(mine):
        x0, x1, g = coords[p]
        resident[p] = (torch.from_numpy(x0.ravel()).to(dev), torch.from_numpy(x1.ravel()).to(dev),
                       torch.from_numpy(g).to(dev))
    dps = [DevicePair(pipe.contexts[k], torch, max_i, max_j, args.kernel) for k in range(n_streams)]

    def worker_steps(k, first, last, timed):
        """Steps first + k, first + k + streams, ... < last on context k.  Returns per-solve infos."""
        dp, st = dps[k], tstreams[k]
        out = []
        with torch.cuda.stream(st):
            if timed:
                starts[k].record(st)
            for s in range(first + k, last, n_streams):
                dp.attach(mine[s], *resident[mine[s]])
                out.append((mine[s], dp.run(prm)))
            if timed:
                ends[k].record(st)
        return out

    def run_block(first, last, timed):
        threads, results = [], [None] * n_streams

        def body(k):
            results[k] = worker_steps(k, first, last, timed)
        for k in range(n_streams):
            t = threading.Thread(target=body, args=(k,))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        return [r for part in results for r in part]

    starts = [torch.cuda.Event(enable_timing=True) for _ in range(n_streams)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_streams)]
    run_block(0, args.warmup, False)
    big = max(set(mine), key=lambda p: p[0] * p[1])
    for k in range(n_streams):                 # size every context's grow-only workspaces once, outside the timing
        with torch.cuda.stream(tstreams[k]):
            dps[k].attach(big, *resident[big])
            dps[k].run(prm)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    done = run_block(args.warmup, total, True)
    torch.cuda.synchronize()
    gpu_ms = max(a.elapsed_time(b) for a in starts for b in ends)
    barrier()
    clocks = sampler.stop()
    iters = launches = 0
    mv_bytes = mv_launch = 0
    solve_ms = 0.0
    entry_iters = 0.0
    for (n0, n1, _), infos in done:
        for inf in infos:
            iters += inf["iters"]
            launches += inf["launches"]
            solve_ms += inf["gpu_ms"]
            entry_iters += float(inf["iters"]) * n0 * n1
            b_, n_ = matvec_bytes(inf, n0, n1, n1 <= 23040)
            mv_bytes += b_
            mv_launch += n_
        launches += 6 * 2 + 1 + 2 + 1       # median passes, cost (+pad), coupling
    t_dev = max_over_ranks(gpu_ms / 1e3)
    total_steps = args.steps * world
    value = total_steps / t_dev
    iters_all = sum_over_ranks(iters)

    # ---- leg 2: end to end through the host-buffer C ABI ---------------------------------------
    del dps, resident
    for c in pipe.contexts:
        c.lib.wotb_release_workspace(c.handle)
    torch.cuda.empty_cache()
    # n_streams solves on the SMs plus one more context whose coupling is crossing PCIe meanwhile
    # (Pipeline compute_slots; a step is still one day-pair through the host-buffer call)
    n_e2e = n_streams + 1 if n_streams > 1 else 1
    pipe2 = Pipeline(local_rank, n_e2e, compute_slots=n_streams if n_e2e > n_streams else 0)
    biggest = max(set(mine), key=lambda p: p[0] * p[1])
    pin_in = {}
    for p in set(mine[args.warmup:] + [biggest]):
        x0, x1, g = coords[p]
        bufs = []
        for arr in (x0, x1, g):
            pa = _pinned.empty(arr.shape, np.float64)
            pa[...] = arr
            bufs.append(pa)
        pin_in[p] = bufs
    outs = [_pinned.empty((max_i * max_j,), np.float64) for _ in range(n_e2e)]

    def e2e_step(k, p, growth_iters=GROWTH_ITERS):
        x0, x1, g = pin_in[p]
        view = outs[k][: p[0] * p[1]].reshape(p[0], p[1])
        wot_ot.solve_coords(x0, x1, g, _lib.SOLVER_DUALITY_GAP, growth_iters=growth_iters, kernel=args.kernel,
                            out=view, ctx=pipe2.contexts[k], **DEFAULTS)

    def e2e_block(steps_of):
        # worker k owns context k and output buffer k; it takes the next step from a shared counter
        threads = []
        for k in range(n_e2e):
            t = threading.Thread(target=steps_of, args=(k,))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()

    e2e_block(lambda k: e2e_step(k, biggest, 1))     # warm every context's workspaces on the largest pair
    barrier()
    next_step = [args.warmup]
    lock = threading.Lock()

    def timed_steps(k):
        while True:
            with lock:
                s = next_step[0]
                next_step[0] += 1
            if s >= total:
                return
            e2e_step(k, mine[s])

    t0 = time.perf_counter()
    e2e_block(timed_steps)
    h2d = d2h = 0
    for s in range(args.warmup, total):
        p = mine[s]
        h2d += (p[0] + p[1]) * D * 8 + p[0] * 8
        d2h += p[0] * p[1] * 8 + (GROWTH_ITERS + 1) * p[0] * 8 + (p[0] + p[1]) * 8
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_steps / t_e2e
    pipe2.close()
    del outs
    peak_c = C.c_double()
    _lib.check(ctx.lib.wotb_bench_mufu_dev(ctx.handle, C.byref(peak_c)))
    mufu_peak = peak_c.value / 1e12

    # ---- configs[3] on N > 1 GPUs: one 100k x 100k pair, rows sharded, peer-memory exchange (and NCCL beside it) ----
    row_sharded = None
    if world > 1 and not args.no_extras:
        for c in pipe.contexts:
            c.lib.wotb_release_workspace(c.handle)
        torch.cuda.empty_cache()
        n_big = max(512, int(100000 * args.scale))
        try:
            row_sharded = row_sharded_block(torch, dist, rank, world, local_rank, mufu_peak, n=n_big)
        except Exception as exc:  # noqa: BLE001 - every rank takes the same path (the solve is collective)
            row_sharded = {"error": "%s: %s" % (type(exc).__name__, exc)}
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernels, measured live ----------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ri, rj = int(np.mean([p[0] for p in pairs])), int(np.mean([p[1] for p in pairs]))
    ms_row, ms_col, ms_fused = isolated_matvec(ctx, torch, ri, rj)
    ld = (rj + 31) // 32 * 32
    alg = ri * ld * 4
    if ms_fused > 0:
        achieved = alg / (ms_fused * 1e-3) / 1e9
        kernel = "k_fused + k_col_finish (one Sinkhorn iteration, K streamed once through shared memory)"
    else:
        achieved = 2 * alg / ((ms_row + ms_col) * 1e-3) / 1e9
        kernel = "k_row + k_col (stored-K matvec pair = one Sinkhorn iteration)"
    roofline_stored = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "kernel": kernel,
        "shape": [ri, rj], "algorithmic_bytes_per_launch": alg,
        "fused_iter_ms": ms_fused, "row_ms": ms_row, "col_ms": ms_col,
        "row_gbs": alg / (ms_row * 1e-3) / 1e9, "col_gbs": alg / (ms_col * 1e-3) / 1e9,
    }
    roofline_stored["traffic"], roofline_stored["traffic_source"] = measured_traffic(
        "k_fused", 615000000, "ncu --set full, profiles/r1b_k_fused_ncu_full.txt: 3.075 GB read per 5-iteration launch "
                              "at this shape (round 1 capture)")
    roofline_online = online_pass_roofline(ctx, torch, ri, rj)      # launch mode of the timed region
    if n_streams > 1:
        # one solve at a time launches the passes with programmatic dependent launch (the pipeline turns it off)
        ctx.lib.wotb_set_pdl(1)
        roofline_online["frac_single_stream_pdl"] = online_pass_roofline(ctx, torch, ri, rj)["frac"]
        ctx.lib.wotb_set_pdl(0)
    if online:
        roofline = roofline_online
        # all-in: every Sinkhorn iteration evaluates 2*I*J exponentials; the denominator is the whole timed region
        # (median, operand packing, convergence checks, final row sums and coupling included)
        roofline["in_step_frac"] = 2.0 * entry_iters / (gpu_ms * 1e-3) / roofline["peak"] / 1e12
        other = ("roofline_stored", roofline_stored)
    else:
        roofline = roofline_stored
        roofline["in_step_gbs"] = mv_bytes / (gpu_ms * 1e-3) / 1e9
        roofline["in_step_note"] = ("bytes of K streamed by the Sinkhorn iterations of the timed steps / the whole timed "
                                    "region (cost, median, K builds, checks and coupling included)")
        other = ("roofline_online", roofline_online)

    cpu = None
    if not args.no_cpu_baseline:
        cores = host_threads()
        sec, m0, m1, kind = cpu_reference_sample(pairs[0], args.cpu_cells, growth_iters=1)
        mean_ij = float(np.mean([a * b for a, b, _ in pairs]))
        sec_per_tmap = sec / (m0 * m1) * mean_ij * GROWTH_ITERS
        cpu = {"value": 1.0 / sec_per_tmap, "unit": "tmaps/s", "cores": cores, "kind": kind,
               "sample": "one full solve (cost + duality-gap solver, growth_iters=1) of atlas pair 0 subsampled to %dx%d "
                         "through the %s in %.1f s; seconds per matrix entry scaled to the mean atlas pair and "
                         "growth_iters=3 (`--impl reference` times full-size pairs)"
                         % (m0, m1, "unmodified reference solver (oracle/_ref)" if kind == "reference" else
                            "NumPy port of the reference (oracle/)", sec)}

    # ---- the other BASELINE configs and the public-API leg (N = 1) ---------------------------------------
    extras = {}
    if world == 1 and not args.no_extras:
        _lib._contexts.pop(local_rank, None)      # `ctx` belongs to the pipeline and dies with it
        pipe.close()
        from wot_b200.pipeline import release_idle_contexts
        for c in list(_lib._contexts.values()):
            c.lib.wotb_release_workspace(c.handle)
        release_idle_contexts()
        torch.cuda.empty_cache()
        extras["e2e_api"] = guarded(lambda: e2e_api_block(local_rank, args.api_pairs, args.scale))
        release_idle_contexts()
        torch.cuda.empty_cache()
        extras["c5"] = guarded(lambda: c5_block(local_rank, n=max(256, int(10000 * args.scale))))
        release_idle_contexts()
        torch.cuda.empty_cache()
        extras["c3"] = guarded(lambda: c3_block(torch, local_rank, mufu_peak, peak, n=max(256, int(50000 * args.scale))))
    if row_sharded is not None:
        extras["row_sharded"] = row_sharded

    line = {
        "metric": "day-pair transport maps per second", "value": value, "unit": "tmaps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": ("f32 exponent (fp16x3 split on tcgen05) / f64 potentials" if online
                                                 else "f32 K / f64 potentials"),
        "data": "synthetic",
        "config": bench_config(args.kernel, n_streams, args.scale),
        "sinkhorn_iters_per_s": iters_all / t_dev,
        "sinkhorn_iters": int(iters_all),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "tmaps/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * t_e2e / args.steps,
                "contexts": n_e2e, "compute_slots": n_streams if n_e2e > n_streams else 0,
                "host_cpus_bound_to_gpu": bound_cpus},
        "gpu_launches": int(launches),
        "roofline": roofline,
        other[0]: other[1],
        "cpu_baseline": cpu,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if pipe.contexts:
        pipe.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import __graft_entry__
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "wot_b200", "csrc", "libwot_b200.so")):
        __graft_entry__.build()
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
