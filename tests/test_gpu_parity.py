"""GPU parity tests: the CUDA path (through the C ABI) against golden vectors of the unmodified reference
and against the float64 oracle.  Run on the B200 box with `-m gpu`.

Tolerance (BASELINE.json north_star): max relative error <= 1e-4 on coupling entries (entries >= 1e-12 * max),
row/column marginals, growth columns and dual potentials; final-stage batch count within +-1.
"""
import numpy as np
import pandas as pd
import pytest

from tests.helpers import DEFAULTS, RTOL, assert_coupling_close, max_rel_err, pair_cost

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ot():
    from wot_b200 import ot as _ot
    return _ot


def _solve(ot, name, C, G, **over):
    params = dict(DEFAULTS, **over)
    tmap = getattr(ot, name)(C=C, G=G, **params)
    return tmap, ot.last_solve_info()


def _check_potentials(info, f, g, eps):
    # potentials enter the coupling as exp(f/eps): 1e-4 relative on entries <=> 1e-4 * eps absolute on f, g
    assert np.max(np.abs(info["f"] - f)) <= RTOL * eps, np.max(np.abs(info["f"] - f))
    assert np.max(np.abs(info["g"] - g)) <= RTOL * eps, np.max(np.abs(info["g"] - g))


def test_reference_golden_case_identity(ot, golden):
    """/root/reference/tests/test_transport.py:20-32."""
    g = golden("ref_3x3")
    for name, tag in (("optimal_transport_duality_gap", "dg"), ("transport_stablev2", "fx")):
        tmap, _ = _solve(ot, name, g["C"], np.ones(3), epsilon=0.01)
        assert np.allclose(tmap, np.eye(3), atol=0.01, rtol=0)
        assert max_rel_err(tmap, g[tag + "_tmap"]) <= RTOL


def test_reference_golden_case_through_otmodel(ot, golden):
    """/root/reference/tests/test_transport.py:11-34 with the same API calls."""
    from wot_b200._anndata import AnnData
    rng = np.random.default_rng(18)
    adata = AnnData(rng.random((6, 1000)), pd.DataFrame({"day": [1, 1, 1, 2, 2, 2]}),
                    pd.DataFrame(index=np.arange(1000)))
    cost = np.array([[0, 100, 100], [100, 0, 100], [100, 100, 0]])
    model = ot.OTModel(adata, epsilon=0.01, lambda1=1, lambda2=50)
    tmap = model.compute_transport_map(1, 2, cost_matrix=cost)
    assert np.allclose(tmap.X, np.eye(3), atol=0.01, rtol=0)
    g = golden("otmodel_3x3")
    assert max_rel_err(tmap.X, g["tmap"]) <= RTOL
    np.testing.assert_allclose(tmap.obs["g1"].values, g["g1"], rtol=RTOL)


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_default_solver_vs_reference(ot, golden, tag):
    g = golden("dg_" + tag)
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    tmap, info = _solve(ot, "optimal_transport_duality_gap", C, G)
    assert_coupling_close(tmap, g["dg_tmap"])
    _check_potentials(info, g["dg_f"], g["dg_g"], 0.05)
    got = info["infos"][0]
    assert got["batches"][:5] == list(g["dg_batches"][:5])
    assert abs(got["batches"][5] - int(g["dg_batches"][5])) <= 1
    assert abs(got["iters"] - int(g["dg_iters"])) <= 5


VARIATIONS = {
    "eps01": dict(epsilon=0.01), "lam10_100": dict(lambda1=10, lambda2=100),
    "loose": dict(epsilon=0.1, lambda1=0.1, lambda2=1), "batch7": dict(batch_size=7),
    "tau1_2": dict(tau=1.2), "tau2_eps02": dict(tau=2.0, epsilon=0.02), "maxiter37": dict(max_iter=37),
    "eps0_2": dict(epsilon0=2.0), "tol1e-5": dict(tolerance=1e-5),
}


@pytest.mark.parametrize("tag", sorted(VARIATIONS))
@pytest.mark.parametrize("use_graph,fuse", [(True, True), (False, True), (True, False)])
def test_parameter_variations_vs_reference(ot, golden, tag, use_graph, fuse):
    g = golden("dg_variations")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    kw = VARIATIONS[tag]
    tmap, info = _solve(ot, "optimal_transport_duality_gap", C, G, use_graph=use_graph, fuse=fuse, **kw)
    assert_coupling_close(tmap, g[tag + "_tmap"])
    eps_final = float(g[tag + "_eps_final"])
    _check_potentials(info, g[tag + "_f"], g[tag + "_g"], eps_final)
    got = info["infos"][0]
    want_batches = [int(b) for b in g[tag + "_batches"]]
    assert got["batches"][:5] == want_batches[:5], (got["batches"], want_batches)
    assert abs(got["batches"][5] - want_batches[5]) <= 1
    assert abs(got["eps_final"] - eps_final) <= 1e-15
    if tag == "maxiter37":
        assert got["iters"] == 37 and got["status"] == 1
    if tag.startswith("tau"):
        assert got["tau_absorptions"] > 0


def test_fixed_iters_vs_reference(ot, golden):
    g = golden("fixed_iters")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    for tag, kw in (("default", {}), ("short", dict(scaling_iter=330, extra_iter=40, inner_iter_max=50)),
                    ("tau1_5", dict(scaling_iter=400, extra_iter=50, tau=1.5))):
        tmap, info = _solve(ot, "transport_stablev2", C, G, **kw)
        assert_coupling_close(tmap, g[tag + "_tmap"])
        _check_potentials(info, g[tag + "_f"], g[tag + "_g"], float(g[tag + "_eps_final"]))
        want_iters = kw.get("scaling_iter", 3000) + kw.get("extra_iter", 1000)
        assert info["infos"][0]["iters"] == want_iters


def test_growth_loop_vs_reference(ot, golden):
    g = golden("growth3")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    tmap, learned = ot.compute_transport_matrix(ot.optimal_transport_duality_gap,
                                                **dict(DEFAULTS, growth_iters=3, C=C, G=G.copy()))
    assert_coupling_close(tmap, g["tmap"])
    assert len(learned) == 3
    np.testing.assert_allclose(np.array(learned), g["learned"], rtol=RTOL)


def test_default_cost_vs_reference(ot, golden):
    from wot_b200 import synthetic
    g = golden("cost_default")
    n0, n1, seed = (int(v) for v in g["shape"])
    x0, x1, _ = synthetic.day_pair_coords(n0, n1, d=30, seed=seed)
    got = ot.OTModel.compute_default_cost_matrix(x0, x1, np.diag(g["sv"]))
    np.testing.assert_allclose(got, g["C"], rtol=1e-13, atol=0)
    got7 = ot.OTModel.compute_default_cost_matrix(x0[:, :7], x1[:, :7])
    np.testing.assert_allclose(got7, g["C_plain7"], rtol=1e-13, atol=0)


def test_median_is_exact(ot):
    """Radix select == np.median bit for bit, odd and even counts, ties, wide dimension."""
    from scipy.spatial.distance import cdist
    from wot_b200.ot import optimal_transport as impl
    rng = np.random.default_rng(3)
    for n0, n1, d in ((33, 35, 5), (64, 64, 30), (130, 77, 70), (257, 300, 30), (1, 1, 3), (2, 1, 4)):
        x0, x1 = rng.normal(size=(n0, d)), rng.normal(size=(n1, d))
        if n0 == 64:
            x1[:32] = x0[:32]          # many exact zeros and ties
        raw = cdist(x0, x1, metric="sqeuclidean")
        C = impl.default_cost_matrix(x0, x1)
        med = np.median(raw)
        np.testing.assert_allclose(C * med, raw, rtol=1e-13, atol=1e-300)
        assert np.median(C) == pytest.approx(1.0, abs=1e-15)


@pytest.mark.parametrize("shape,ties", [((8000, 7601), "none"), ((7999, 7601), "none"), ((9000, 8100), "some"),
                                        ((8200, 7700), "heavy")])
def test_median_sampled_window_is_exact(shape, ties):
    """From 6e7 distances on, the median is located by a random sample and found in ONE pass over the distances
    (count below the window, gather the window, select on the gathered values); it must still be np.median bit for
    bit: odd and even counts, tie groups inside the window, and data whose ties overflow the window (the device
    notices and the three-pass radix select takes over)."""
    import ctypes as C
    from scipy.spatial.distance import cdist
    from wot_b200 import _lib
    n0, n1 = shape
    rng = np.random.default_rng(n0 + n1)
    x0, x1 = rng.normal(size=(n0, 30)), rng.normal(size=(n1, 30))
    if ties == "some":
        x0[100:400] = x0[0]            # 300 identical cells: every column's distance to them repeats 300 times
    elif ties == "heavy":
        x0[:] = x0[rng.integers(0, 3, n0)]      # three distinct cells: each distance value repeats ~400 times
    want = np.median(cdist(x0, x1, metric="sqeuclidean"))
    ctx = _lib.context(0)
    out = np.empty((n0, n1))
    med = C.c_double()
    _lib.check(ctx.lib.wotb_default_cost_matrix_host(ctx.handle, _lib.ptr(np.ascontiguousarray(x0)), n0,
                                                     _lib.ptr(np.ascontiguousarray(x1)), n1, 30, None, _lib.ptr(out), C.byref(med)))
    assert med.value == want


def test_otmodel_default_path_vs_reference(ot, golden):
    """PCA -> cost -> solver -> growth columns (ot_model.py:255-326) against the unmodified reference."""
    from wot_b200 import synthetic
    from wot_b200._anndata import AnnData
    g = golden("otmodel_path")
    cells = [int(c) for c in g["cells"]]
    X, day, growth = synthetic.expression_matrix(cells, n_genes=int(g["n_genes"]), seed=int(g["seed"]))
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = ot.OTModel(adata, growth_iters=2)
    tm = model.compute_transport_map(0, 1)
    assert list(tm.obs.index) == list(g["obs_index"]) and list(tm.var.index) == list(g["var_index"])
    assert list(tm.obs.columns) == ["g0", "g1", "g2"]
    assert_coupling_close(np.asarray(tm.X), g["tmap"])
    for col in ("g0", "g1", "g2"):
        np.testing.assert_allclose(tm.obs[col].values, g[col], rtol=RTOL)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("shape", [(1500, 1637), (2000, 2000), (997, 4099), (4100, 513), (300, 19000), (130, 23040),
                                   (200, 24001), (150, 30011), (301, 47000), (70, 95000)])
def test_default_solver_vs_oracle_from_coords(ot, shape, fuse):
    """Config 1 scale: GPU default cost + STORED-kernel solver from coordinates against the float64 oracle.  fuse=True:
    K read once per iteration -- one CTA per row up to 23,040 columns (k_fused), a thread-block cluster per row beyond
    (k_fused_cl: 24,001 -> 2 CTAs, 47,000 -> 4, 95,000 -> 8); fuse=False: the two-sweep kernels."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    n0, n1 = shape
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=n0 + n1)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **DEFAULTS)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), fuse=fuse, kernel="stored", **DEFAULTS)
    assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    _check_potentials(got, info.f, info.g, 0.05)
    assert got["infos"][0]["batches"][:5] == info.batches[:5]
    assert abs(got["infos"][0]["batches"][5] - info.batches[5]) <= 1
    np.testing.assert_allclose(got["learned_growth"][-1], want.sum(axis=1), rtol=RTOL)


def test_out_dtype_float32_and_out_buffer(ot):
    C, G = pair_cost(200, 210, 5)
    ref, _ = _solve(ot, "optimal_transport_duality_gap", C, G)
    out = np.empty((200, 210), dtype=np.float32)
    got, _ = _solve(ot, "optimal_transport_duality_gap", C, G, out=out, out_dtype=np.float32)
    assert got is out
    np.testing.assert_allclose(got, ref, rtol=2e-7, atol=1e-37)


def test_fixed_point_property_large(ot):
    """Size-independent property at an atlas-scale pair: the returned potentials satisfy the unbalanced
    Sinkhorn fixed-point equations (optimal_transport.py:133-134) evaluated in float64 on the host."""
    from wot_b200 import synthetic
    n0, n1 = 6000, 7000
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=99)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), **DEFAULTS)
    info = ot.last_solve_info()
    eps, l1, l2 = 0.05, 1.0, 50.0
    f, g = info["f"], info["g"]
    rows = tmap.sum(axis=1) * n1        # r_i = a_i (K b)_i
    cols = tmap.sum(axis=0) * n1
    # at a fixed point: a = (p / (K b dy))^alpha1 e^{-u/(l1+eps)}  <=>  r_i / J = p_i exp(-f_i / l1)
    np.testing.assert_allclose(rows / n1, growth * np.exp(-f / l1), rtol=5e-4)
    np.testing.assert_allclose(cols / n0, growth.mean() * np.exp(-g / l2), rtol=5e-4)
    # growth rows come from the solver's own last pass (default kernel 'auto' -> online: exponent error ~1e-5)
    np.testing.assert_allclose(info["learned_growth"][-1], tmap.sum(axis=1), rtol=RTOL / 4)


# ---------------------------------------------------------------------------------------------------
# online kernel (K recomputed from coordinates, never stored)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["online", "online_simt"])
@pytest.mark.parametrize("shape,d", [((1500, 1637), 30), ((2000, 2000), 30), ((700, 3001), 30), ((3001, 700), 30),
                                     ((900, 1000), 5), ((600, 650), 50), ((130, 129), 33), ((257, 5000), 30)])
def test_online_kernel_vs_oracle(ot, shape, d, kernel):
    """exp((f_i + g_j - C_ij)/eps) recomputed tile by tile: same couplings, potentials and batch counts."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    n0, n1 = shape
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=n0 + n1 + d)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **DEFAULTS)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), kernel=kernel, **DEFAULTS)
    rep = assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    _check_potentials(got, info.f, info.g, 0.05)
    # batch counts within +-1 per stage (north_star).  The warm-stage criterion (:158-160) is a 1e-6 threshold on the
    # change of the iterates, so the ~1e-5 exponent error of the online kernels can move a crossing by one batch.
    batches = got["infos"][0]["batches"]
    assert all(abs(batches[k] - info.batches[k]) <= 1 for k in range(6)), (batches, info.batches, rep)
    np.testing.assert_allclose(got["learned_growth"][-1], want.sum(axis=1), rtol=RTOL)


@pytest.mark.parametrize("kernel", ["online", "online_simt"])
def test_online_kernel_growth_scale_and_fixed_iters(ot, kernel):
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(500, 560, d=30, seed=77)
    sv = np.linspace(3.0, 0.5, 30)
    cost = orc.compute_default_cost_matrix(x0, x1, np.diag(sv))
    params = dict(DEFAULTS, growth_iters=2)
    want, learned = orc.compute_transport_matrix(orc.optimal_transport_duality_gap, C=cost, G=growth.copy(), **params)
    tmap, got_learned = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, sv), C=None,
                                                    G=growth.copy(), kernel=kernel, **params)
    assert_coupling_close(tmap, want)
    np.testing.assert_allclose(np.array(got_learned), np.array(learned), rtol=RTOL)
    short = dict(DEFAULTS, scaling_iter=330, extra_iter=40, inner_iter_max=50)
    want = orc.transport_stablev2(C=cost, G=growth, **short)
    tmap, _ = ot.compute_transport_matrix(ot.transport_stablev2, coords=(x0, x1, sv), C=None, G=growth.copy(),
                                          kernel=kernel, **short)
    assert_coupling_close(tmap, want)


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("shape,d", [((300, 340), 30), ((3000, 3301), 30), ((700, 2500), 12), ((1111, 777), 40)])
def test_online_pass_kernels_vs_numpy(impl, shape, d):
    """One online-kernel pass, sums[i] = sum_j exp2(off_out_i + off_in_j + scale^2 <x_i, y_j>), against float64
    NumPy: the SIMT FP32 kernel (impl 0) and the tcgen05 kernel with 8 / 16 epilogue warps (impl 1 / 2).
    Exponent error budget ~1e-5 (3-term fp16 split / fp32 dot product) -> row sums within 5e-5."""
    import ctypes as C
    import torch
    from tools.online_pass_check import make_inputs, reference
    from wot_b200 import _lib
    n_out, n_in = shape
    x0, x1, scale, off_out, off_in = make_inputs(n_out, n_in, d, seed=n_out + d)
    want = reference(x0, x1, scale, off_out, off_in)
    ctx = _lib.context(0)
    dev = "cuda:%d" % ctx.device
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, x1, off_out, off_in)]
    sums = torch.empty(n_out, dtype=torch.float64, device=dev)
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.wotb_online_rowsums_dev(ctx.handle, P(t[0]), n_out, P(t[1]), n_in, d, float(scale), P(t[2]),
                                               P(t[3]), impl, 1, P(sums), None))
    got = sums.cpu().numpy()
    assert np.max(np.abs(got - want) / want) <= 5e-5


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] as parity cases
# ---------------------------------------------------------------------------------------------------
def _blockwise_marginals(x0, x1, median, f, g, eps, block=4096):
    """float64 restatement of the coupling's marginals (optimal_transport.py:153,164 with the SURVEY section 8
    a-note form tmap_ij = exp((f_i + g_j - C_ij)/eps)/J), evaluated blockwise with torch float64 on the GPU
    so that 50k x 50k and 100k x 100k fit.  Independent of the library's kernels.  Also counts the entries of
    the unnormalised cost below / at-or-below `median` (np.median property, ot_model.py:252)."""
    import torch
    dev = "cuda:0"
    X0, X1 = torch.from_numpy(x0).to(dev), torch.from_numpy(x1).to(dev)
    F, G = torch.from_numpy(f).to(dev), torch.from_numpy(g).to(dev)
    n1 = X1.shape[0]
    sq1 = (X1 * X1).sum(1)
    rows = torch.empty(X0.shape[0], dtype=torch.float64, device=dev)
    cols = torch.zeros(n1, dtype=torch.float64, device=dev)
    below = at_or_below = 0
    for r in range(0, X0.shape[0], block):
        xb = X0[r:r + block]
        dist = ((xb * xb).sum(1)[:, None] + sq1[None, :] - 2.0 * (xb @ X1.T)).clamp_(min=0.0)
        below += int((dist < median * (1 - 1e-9)).sum())
        at_or_below += int((dist <= median * (1 + 1e-9)).sum())
        dist.div_(-median * eps).add_(F[r:r + block, None] / eps).add_(G[None, :] / eps).exp_().div_(n1)
        rows[r:r + block] = dist.sum(1)
        cols += dist.sum(0)
        del dist
    return rows.cpu().numpy(), cols.cpu().numpy(), below, at_or_below


@pytest.mark.parametrize("n0,n1,seed,kernels", [(50000, 50000, 2, ("stored", "online")), (100000, 100000, 3, ("online",))])
def test_full_size_pairs_fixed_point_and_marginals(ot, n0, n1, seed, kernels):
    """configs[2] (50k x 50k, stored-K and online-K on one GPU) and configs[3] (100k x 100k, online-K; the
    row-sharded form of the same solve is in test_gpu_multi.py) at FULL size, where neither the reference nor
    the oracle can allocate.  Size-independent checks, all against an independent float64 blockwise evaluation:
    (1) the median is the exact np.median of the I*J distances; (2) row sums returned by the solver equal the
    float64 marginals of exp((f+g-C)/eps)/J; (3) f, g satisfy the unbalanced Sinkhorn fixed-point equations
    (optimal_transport.py:133-134); (4) stored-K and online-K agree within the north_star tolerance."""
    import ctypes as C
    import torch
    from wot_b200 import _lib, synthetic
    free, _ = torch.cuda.mem_get_info()
    if "stored" in kernels and free < 60e9:
        pytest.skip("needs 60 GB of HBM")
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=seed)
    eps, l1, l2 = DEFAULTS["epsilon"], float(DEFAULTS["lambda1"]), float(DEFAULTS["lambda2"])
    res = {}
    for kernel in kernels:
        _, learned = ot.optimal_transport.solve_coords(x0, x1, growth, _lib.SOLVER_DUALITY_GAP, kernel=kernel,
                                                       want_tmap=False,
                                                       **{k: v for k, v in DEFAULTS.items() if k != "growth_iters"})
        info = dict(ot.last_solve_info())
        res[kernel] = (np.array(info["f"]), np.array(info["g"]), learned[-1].copy(), info["median"], info["infos"][0])
        ctx = _lib.context(0)
        ctx.lib.wotb_release_workspace(ctx.handle)
        torch.cuda.empty_cache()
    total = n0 * n1
    for kernel, (f, g, rowsum, median, inf) in res.items():
        assert inf["status"] == 0 and abs(inf["gap"]) < DEFAULTS["tolerance"], inf
        rows, cols, below, at_or_below = _blockwise_marginals(x0, x1, median, f, g, eps)
        # (1) np.median of an even count is the mean of the two middle values; either way at most half of the
        # entries lie strictly below it and at least half at or below it
        assert below <= total // 2 <= at_or_below, (below, at_or_below, total)
        # (2) solver's own row sums vs float64 marginals
        np.testing.assert_allclose(rowsum, rows, rtol=RTOL)
        # (3) fixed point: r_i = p_i exp(-f_i/lambda1), c_j * J / I = q exp(-g_j/lambda2)
        np.testing.assert_allclose(rows, growth * np.exp(-f / l1), rtol=5e-4)
        np.testing.assert_allclose(cols * n1 / n0, growth.mean() * np.exp(-g / l2), rtol=5e-4)
    if len(res) == 2:
        a, b = res["stored"], res["online"]
        assert a[3] == b[3]
        assert np.max(np.abs(a[0] - b[0])) <= RTOL * eps and np.max(np.abs(a[1] - b[1])) <= RTOL * eps
        np.testing.assert_allclose(a[2], b[2], rtol=RTOL)
        assert all(abs(x - y) <= 1 for x, y in zip(a[4]["batches"], b[4]["batches"]))


SWEEP_CORNERS = [dict(epsilon=0.01, lambda1=0.1, lambda2=1), dict(epsilon=0.01, lambda1=50, lambda2=100),
                 dict(epsilon=0.025, lambda1=10, lambda2=10), dict(epsilon=0.05, lambda1=0.1, lambda2=100),
                 dict(epsilon=0.1, lambda1=50, lambda2=1), dict(epsilon=0.1, lambda1=1, lambda2=50),
                 dict(epsilon=0.005, lambda1=1, lambda2=50)]


@pytest.mark.parametrize("kernel", ["stored", "online"])
@pytest.mark.parametrize("setting", SWEEP_CORNERS, ids=lambda s: "eps%g_l%g_%g" % (s["epsilon"], s["lambda1"], s["lambda2"]))
def test_sweep_settings_vs_oracle(ot, setting, kernel):
    """configs[4]: corners and interior points of the 64-setting (epsilon, lambda1, lambda2) grid, plus epsilon = 0.005,
    on a pair the oracle finishes in seconds; same couplings, potentials and batch counts (70 ... 32,655 iterations).
    'online' is the product default: below a final epsilon of 0.02 the library runs the tcgen05 pass on its precise
    6-segment operands (exact accumulation of the large cancelling terms + compensation of the accumulator's
    round-toward-zero, csrc/online_pass.cuh)."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(420, 460, d=30, seed=4)
    params = dict(DEFAULTS, **setting)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **params)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), kernel=kernel, **params)
    assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    # potentials enter the coupling as exp(f / eps): 1e-4 relative on entries <=> 1e-4 * eps absolute on f and g
    _check_potentials(got, info.f, info.g, setting["epsilon"])
    # Batch counts: the final (duality-gap) stage within +-1 (north_star).  The warm stages end when the change of the
    # iterates falls below 1e-6 (:158-160): the stored kernel reproduces them exactly; the online kernel's per-entry
    # rounding noise (~1e-6, the ulp of an fp32 exponent) can move that crossing by 0.2 % of a slowly converging
    # stage (3 batches of 1981 at epsilon 0.01, lambda 50/100).
    batches = got["infos"][0]["batches"]
    assert abs(batches[5] - info.batches[5]) <= 1, (batches, info.batches)
    exact = kernel == "stored" and setting["epsilon"] >= 0.01     # stored at epsilon 0.005: one warm batch off (154 / 153)
    warm_tol = [0 if exact else max(1, int(np.ceil(0.002 * info.batches[k]))) for k in range(5)]
    assert all(abs(batches[k] - info.batches[k]) <= warm_tol[k] for k in range(5)), (batches, info.batches)
    assert ot.optimal_transport.resolve_kernel("auto", 420, 460, 30, setting["epsilon"]) == "online"


@pytest.mark.parametrize("kernel", ["online_fast", "online_precise"])
def test_online_operand_modes_at_defaults(ot, kernel):
    """Both operand forms of the tcgen05 pass at the default epsilon: couplings 1e-4, identical batch counts."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(900, 1000, d=30, seed=11)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **DEFAULTS)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), kernel=kernel, **DEFAULTS)
    assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    _check_potentials(got, info.f, info.g, 0.05)
    assert abs(got["infos"][0]["batches"][5] - info.batches[5]) <= 1


@pytest.mark.parametrize("kernel", ["online_fast", "online_precise"])
def test_persistent_batch_kernel_equals_pass_per_launch(ot, kernel):
    """k_online_batch (one cooperative launch per batch of iterations, opt-in) computes exactly what the
    pass-per-launch form computes: same slot-order reductions, same state machine -> bit-identical potentials."""
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(3300, 3200, d=30, seed=8)      # >= 148 work units in both passes
    res = {}
    for batch in (False, True):
        ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None, G=growth.copy(),
                                    kernel=kernel, online_batch=batch, want_tmap=False, **DEFAULTS)
        got = ot.last_solve_info()
        res[batch] = (np.array(got["f"]), np.array(got["g"]), got["infos"][0]["batches"], got["infos"][0]["launches"])
    assert res[True][2] == res[False][2]
    assert np.array_equal(res[True][0], res[False][0]) and np.array_equal(res[True][1], res[False][1])
    assert res[True][3] < res[False][3]          # far fewer launches: the batch kernel really ran


def test_atlas_size_pair_vs_oracle(ot):
    """configs[1] at atlas size against the oracle itself (not only size-independent properties): 5000 x 5200 cells,
    growth_iters = 3, default kernel policy; couplings, growth columns, potentials, batch counts."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(5000, 5200, d=30, seed=2026)
    params = dict(DEFAULTS, growth_iters=3)
    cost = orc.compute_default_cost_matrix(x0, x1)
    infos = []

    def solver(**kw):
        info = orc.SolveInfo()
        infos.append(info)
        return orc.optimal_transport_duality_gap(info=info, gap="marginal", **kw)

    want, learned = orc.compute_transport_matrix(solver, C=cost, G=growth.copy(), **params)
    tmap, got_learned = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                                    G=growth.copy(), **params)
    assert_coupling_close(tmap, want)
    np.testing.assert_allclose(np.array(got_learned), np.array(learned), rtol=RTOL)
    got = ot.last_solve_info()
    _check_potentials(got, infos[-1].f, infos[-1].g, 0.05)
    for k in range(3):
        b, w = got["infos"][k]["batches"], infos[k].batches
        assert abs(b[5] - w[5]) <= 1 and all(abs(x - y) <= 1 for x, y in zip(b[:5], w[:5])), (k, b, w)
    np.testing.assert_allclose(got["learned_growth"][-1], want.sum(axis=1), rtol=RTOL)


def test_otmodel_covariate_path(ot, tmp_path):
    """ot_model.py:159-163, :280-282: with_covariates=True computes one map per (day pair, covariate pair) on the
    cells that carry those covariate values and names the files '{prefix}_{t0}_{t1}_cv{a}_cv{b}'."""
    from wot_b200 import h5ad, synthetic
    from wot_b200._anndata import AnnData
    X, day, growth = synthetic.expression_matrix([260, 300], n_genes=120, seed=5)
    rng = np.random.default_rng(3)
    cov = rng.integers(0, 2, len(day))
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth, "covariate": cov}, index=["c%d" % i for i in range(len(day))])
    var = pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])])
    model = ot.OTModel(AnnData(X, obs, var), growth_iters=1, local_pca=10)
    out = str(tmp_path / "maps" / "tm")
    model.compute_all_transport_maps(tmap_out=out, output_file_format="h5ad", with_covariates=True)
    import os
    names = sorted(os.listdir(tmp_path / "maps"))
    assert names == sorted("tm_0.0_1.0_cv%d_cv%d.h5ad" % (a, b) for a in (0, 1) for b in (0, 1))
    for a in (0, 1):
        for b in (0, 1):
            got = h5ad.read_h5ad(str(tmp_path / "maps" / ("tm_0.0_1.0_cv%d_cv%d.h5ad" % (a, b))))
            rows = obs.index[(day == 0) & (cov == a)]
            cols = obs.index[(day == 1) & (cov == b)]
            assert list(got["obs_index"]) == list(rows) and list(got["var_index"]) == list(cols)
            # the same map from a model that only holds those cells
            keep = ((day == 0) & (cov == a)) | ((day == 1) & (cov == b))
            sub = ot.OTModel(AnnData(X[keep], obs[keep], var), growth_iters=1, local_pca=10)
            want = sub.compute_transport_map(0, 1)
            np.testing.assert_allclose(got["X"], np.asarray(want.X), rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(got["obs"]["g1"], want.obs["g1"].values, rtol=1e-12)


def test_parameter_sweep_driver_single_gpu(ot):
    """parallel.parameter_sweep on one GPU (serial queue): per-setting row sums equal per-setting direct solves."""
    from wot_b200 import parallel, synthetic
    x0, x1, growth = synthetic.day_pair_coords(900, 1000, d=30, seed=4)
    grid = parallel.sweep_grid(epsilons=(0.05, 0.1), lambda1s=(1,), lambda2s=(10, 50))
    common = {k: v for k, v in DEFAULTS.items() if k not in ("epsilon", "lambda1", "lambda2", "growth_iters")}
    res = parallel.parameter_sweep(x0, x1, growth, grid, kernel="online", **common)
    assert [r["setting"] for r in res] == grid
    res2 = parallel.parameter_sweep(x0, x1, growth, grid, kernel="online", streams=2, **common)   # two in flight
    for a, b in zip(res, res2):
        assert a["setting"] == b["setting"] and a["iters"] == b["iters"] and a["batches"] == b["batches"]
        np.testing.assert_array_equal(a["rowsum"], b["rowsum"])
    for r in res:
        tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                              G=growth.copy(), kernel="online", **dict(DEFAULTS, **r["setting"]))
        np.testing.assert_allclose(r["rowsum"], tmap.sum(axis=1), rtol=RTOL)
        assert r["status"] == 0 and r["iters"] > 0


def test_pipeline_streams_match_serial(ot):
    """wot_b200.pipeline: day-pairs in flight on separate CUDA streams give bit-identical couplings, potentials and
    batch counts to one-at-a-time solves (every kernel reduces in a fixed order, so timing cannot leak in)."""
    from wot_b200 import _lib, synthetic
    from wot_b200.pipeline import Pipeline
    shapes = [(900, 1000, 1), (1500, 1100, 2), (700, 2100, 3), (1300, 1300, 4), (600, 500, 5)]
    prm = {k: v for k, v in DEFAULTS.items() if k != "growth_iters"}

    def solve(ctx, shape, kernel):
        x0, x1, growth = synthetic.day_pair_coords(shape[0], shape[1], d=30, seed=shape[2])
        tmap, learned = ot.optimal_transport.solve_coords(x0, x1, growth, _lib.SOLVER_DUALITY_GAP, growth_iters=2,
                                                          kernel=kernel, pinned=False, ctx=ctx, **prm)
        info = ot.last_solve_info()
        return np.array(tmap), np.array(learned), np.array(info["f"]), [i["batches"] for i in info["infos"]]

    for kernel in ("online", "stored"):
        serial = [solve(None, s, kernel) for s in shapes]
        with Pipeline(0, streams=3) as pipe:
            piped = pipe.map(lambda ctx, s: solve(ctx, s, kernel), shapes)
        for a, b in zip(serial, piped):
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
            np.testing.assert_array_equal(a[2], b[2])
            assert a[3] == b[3]


def test_compute_all_transport_maps_pipelined_equals_serial(ot, tmp_path):
    """OTModel.compute_all_transport_maps (ot_model.py:124-201) with day-pairs in flight on several streams writes
    the same files, bit for bit, as the serial loop (streams=1), including the row order of '{prefix}_g.txt'."""
    import os
    from wot_b200 import synthetic
    from wot_b200._anndata import AnnData
    X, day, growth = synthetic.expression_matrix([310, 280, 350, 300, 330], n_genes=80, seed=21)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    var = pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])])
    outs = {}
    for streams in (1, 2):
        model = ot.OTModel(AnnData(X.copy(), obs.copy(), var.copy()), growth_iters=2, local_pca=10, streams=streams)
        d = tmp_path / ("s%d" % streams)
        d.mkdir()
        model.compute_all_transport_maps(tmap_out=str(d / "tm"), output_file_format="npz")
        outs[streams] = d
    names = sorted(os.listdir(outs[1]))
    assert names == sorted(os.listdir(outs[2])) and len(names) == 5 and "tm_g.txt" in names
    for n in names:
        if n.endswith(".npz"):
            a, b = np.load(outs[1] / n, allow_pickle=True), np.load(outs[2] / n, allow_pickle=True)
            np.testing.assert_array_equal(a["X"], b["X"])
            np.testing.assert_array_equal(a["obs_values"], b["obs_values"])
        else:
            assert (outs[1] / n).read_text() == (outs[2] / n).read_text()


@pytest.mark.parametrize("cells,genes,k", [([600, 700], 300, 30), ([900, 800], 1479, 30), ([300, 280], 700, 10),
                                           ([3000, 3200], 1000, 30), ([5000, 7000], 1479, 30)])
def test_gpu_pca_vs_sklearn(cells, genes, k):
    """local PCA on the GPU (csrc/pca.cu) against the reference's own call on scikit-learn
    (wot/ot/util.py:240-255): same signs, components and singular values to roundoff, the default cost built from
    them (ot_model.py:242-253) to 1e-9."""
    from oracle import wot_oracle
    from wot_b200 import synthetic
    from wot_b200.ot import util
    X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=3)
    m1, m2 = X[day == 0], X[day == 1]
    p1, p2, pca, mu = util.compute_pca_sklearn(m1, m2, k)
    q1, q2, gpca, mu2 = util.compute_pca(m1, m2, k)
    assert isinstance(gpca, util.LocalPCA)            # 'auto' took the GPU path
    np.testing.assert_allclose(mu2, mu, rtol=0, atol=1e-13)
    np.testing.assert_allclose(gpca.singular_values_, pca.singular_values_, rtol=1e-10)
    np.testing.assert_allclose(np.vstack([q1, q2]), np.vstack([p1, p2]), rtol=0, atol=1e-8)
    if sum(cells) <= 7000:
        want = wot_oracle.compute_default_cost_matrix(p1, p2, np.diag(pca.singular_values_))
        got = wot_oracle.compute_default_cost_matrix(q1, q2, np.diag(gpca.singular_values_))
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)


def test_config0_otmodel_2k_pair_vs_oracle(ot):
    """BASELINE.json configs[0] at full size: 2 days x 2,000 cells x 1,000 genes through OTModel defaults
    (local_pca=30, eps=0.05, lambda1=1, lambda2=50, growth_iters=1): GPU PCA -> GPU cost -> GPU solver, against the
    float64 oracle fed with the reference's own scikit-learn PCA (util.py:240-255)."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    from wot_b200._anndata import AnnData
    from wot_b200.ot import util
    X, day, growth = synthetic.expression_matrix([2000, 2000], n_genes=1000, seed=0)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    model = ot.OTModel(AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])])))
    tm = model.compute_transport_map(0, 1)
    got_info = ot.last_solve_info()["infos"][0]
    p0, p1, pca, _ = util.compute_pca_sklearn(X[day == 0], X[day == 1], 30)
    cost = orc.compute_default_cost_matrix(p0, p1, np.diag(pca.singular_values_))
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=cost, G=np.power(growth[day == 0], 1.0), info=info, gap="marginal", **DEFAULTS)
    assert_coupling_close(np.asarray(tm.X), want)
    assert abs(got_info["batches"][5] - info.batches[5]) <= 1
    np.testing.assert_allclose(tm.obs["g1"].values, want.sum(axis=1), rtol=RTOL)


def test_implicit_transport_map_push_forward_pull_back(ot):
    """populations pushed forward / pulled back through a coupling that is never materialised
    (wot_b200.tmap.ImplicitTransportMap) equal the reference's dense products p @ tmap.X and tmap.X @ p.T
    (transport_map_model.py:290, :356) on the coupling compute_transport_map returns for the same pair."""
    from wot_b200 import synthetic
    from wot_b200._anndata import AnnData
    X, day, growth = synthetic.expression_matrix([700, 820, 760], n_genes=200, seed=9)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    model = ot.OTModel(AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])])), growth_iters=2,
                       local_pca=20)
    dense = model.compute_transport_map(0, 1)
    imp = model.compute_implicit_transport_map(0, 1)
    T = np.asarray(dense.X)
    assert imp.shape == T.shape and list(imp.obs.columns) == ["g0", "g1", "g2"]
    np.testing.assert_allclose(imp.obs["g2"].values, dense.obs["g2"].values, rtol=RTOL)
    rng = np.random.default_rng(0)
    p_rows = np.vstack([np.ones(700) / 700, (rng.random(700) < 0.1).astype(float), rng.random(700)])
    p_cols = np.vstack([np.ones(820) / 820, (rng.random(820) < 0.05).astype(float)])
    np.testing.assert_allclose(imp.push_forward(p_rows), p_rows @ T, rtol=RTOL)
    np.testing.assert_allclose(imp.pull_back(p_cols), (T @ p_cols.T).T, rtol=RTOL)
    np.testing.assert_allclose(imp.row_sums(), T.sum(axis=1), rtol=RTOL)
    np.testing.assert_allclose(imp.col_sums(), T.sum(axis=0), rtol=RTOL)
    # normalised push-forward then pull-back of a cell set, as TransportMapModel.ancestors / descendants chain them
    q = imp.push_forward(p_rows[1], normalize=True)
    np.testing.assert_allclose(q, (p_rows[1] @ T) / (p_rows[1] @ T).sum(), rtol=RTOL)
    np.testing.assert_allclose(imp.push_forward(-p_rows[0]), -(p_rows[0] @ T), rtol=RTOL)   # any sign (float64 apply)
    with pytest.raises(ValueError):
        imp.push_forward(np.ones(701))


@pytest.mark.parametrize("kernel", ["stored", "online", "online_simt"])
@pytest.mark.parametrize("shape,d", [((1, 1), 3), ((1, 7), 2), ((6, 1), 4), ((2, 3), 1), ((5, 300), 30), ((257, 3), 30),
                                     ((33, 65), 46), ((40, 50), 47)])
def test_tiny_and_ragged_shapes_vs_oracle(ot, shape, d, kernel):
    """Edge shapes: single cells on either side, fewer cells than one tile, one coordinate, the largest d the
    tcgen05 pass takes (46) and the first one it hands to the SIMT pass (47)."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    n0, n1 = shape
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=17 + n0 + n1)
    if n0 * n1 == 1:
        x1 = x1 + 1.0       # a single pair at distance 0 has median 0 (the reference divides by it, too)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **DEFAULTS)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), kernel=kernel, **DEFAULTS)
    assert tmap.shape == want.shape
    assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    assert abs(got["infos"][0]["batches"][5] - info.batches[5]) <= 1


@pytest.mark.parametrize("eps", [0.05, 0.01])
@pytest.mark.parametrize("shape,d", [((1, 1), 3), ((2, 300), 1), ((300, 2), 5), ((257, 129), 14), ((130, 700), 30),
                                     ((700, 513), 31), ((300, 340), 46)])
def test_precise_operands_edge_shapes_vs_oracle(ot, shape, d, eps):
    """The 6-segment operand form over every K-segment width (kseg 16 / 32 / 48: the last one leaves shared memory for
    ONE B stage), ragged and tiny shapes, at the default epsilon (forced) and at 0.01 (the library's own choice)."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    n0, n1 = shape
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=29 + n0 + n1 + d)
    if n0 * n1 == 1:
        x1 = x1 + 1.0
    params = dict(DEFAULTS, epsilon=eps)
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **params)
    tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                          G=growth.copy(), kernel="online_precise" if eps >= 0.02 else "online", **params)
    assert_coupling_close(tmap, want)
    got = ot.last_solve_info()
    # The coupling sees f_i + g_j: that sum is held to 1e-4 * eps on every pair.  With a handful of cells a constant may
    # move between f and g (nearly balanced transport determines it only through the lambda terms): each vector alone
    # gets twice the allowance here (measured: 1.04e-4 * eps on the 2 x 300, d = 1 case, equal and opposite in f and g).
    df, dg = got["f"] - info.f, got["g"] - info.g
    assert np.max(np.abs(df[:, None] + dg[None, :])) <= RTOL * eps
    assert np.max(np.abs(df)) <= 2 * RTOL * eps and np.max(np.abs(dg)) <= 2 * RTOL * eps
    assert abs(got["infos"][0]["batches"][5] - info.batches[5]) <= 1


def test_fixed_iters_small_epsilon_online_vs_oracle(ot):
    """transport_stablev2 (optimal_transport.py:167-236) with a final epsilon of 0.01 on the online kernel: precise
    operands are chosen from `epsilon` itself for this solver (its schedule ends there, :184-185)."""
    from oracle import wot_oracle as orc
    from wot_b200 import synthetic
    x0, x1, growth = synthetic.day_pair_coords(350, 410, d=30, seed=91)
    cost = orc.compute_default_cost_matrix(x0, x1)
    short = dict(DEFAULTS, epsilon=0.01, scaling_iter=400, extra_iter=60, inner_iter_max=50)
    want = orc.transport_stablev2(C=cost, G=growth, **short)
    tmap, _ = ot.compute_transport_matrix(ot.transport_stablev2, coords=(x0, x1, None), C=None, G=growth.copy(),
                                          kernel="online", **short)
    assert_coupling_close(tmap, want)


def test_nan_gap_raises_like_the_reference(ot):
    """optimal_transport.py:162-163: a NaN duality gap raises RuntimeError("Overflow encountered in duality gap
    computation, ...").  A caller-supplied cost 1e4 times the normalised one underflows every K entry in the reference
    (float64) and here (fp32 K) alike."""
    import warnings
    from oracle import wot_oracle as orc
    C, G = pair_cost(60, 70, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(RuntimeError, match="Overflow encountered in duality gap computation"):
            orc.optimal_transport_duality_gap(C=C * 1e4, G=G, info=orc.SolveInfo(), **DEFAULTS)
    with pytest.raises(RuntimeError, match="Overflow encountered in duality gap computation"):
        _solve(ot, "optimal_transport_duality_gap", C * 1e4, G)
    # the context stays usable after the error
    tmap, _ = _solve(ot, "optimal_transport_duality_gap", C, G)
    assert np.isfinite(tmap).all()


@pytest.mark.parametrize("tag", ["tall", "wide"])
def test_gpu_pca_vs_reference_golden(golden, tag):
    """csrc/pca.cu against tests/golden/pca_randomized.npz (the unmodified reference's wot.ot.compute_pca,
    util.py:240-255, randomized-solver shapes in both orientations): same signs, loadings, singular values."""
    from wot_b200 import synthetic
    from wot_b200.ot import util
    g = golden("pca_randomized")
    cells, genes, k = [int(c) for c in g[tag + "_cells"]], int(g[tag + "_genes"]), int(g[tag + "_k"])
    X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=int(g["seed"]))
    q0, q1, pca, mean = util.compute_pca(X[day == 0], X[day == 1], k)
    assert isinstance(pca, util.LocalPCA)
    np.testing.assert_allclose(pca.singular_values_, g[tag + "_sv"], rtol=1e-10)
    np.testing.assert_allclose(mean, g[tag + "_mean"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(np.vstack([q0, q1]), np.vstack([g[tag + "_pca0"], g[tag + "_pca1"]]), rtol=0, atol=1e-8)


def test_cli_optimal_transport_end_to_end(ot, tmp_path):
    """`wot optimal_transport` (wot/commands/optimal_transport.py:12-30) through `python -m wot_b200 optimal_transport`:
    matrix + cell_days + growth-rate files in, '{out}_{t0}_{t1}.npz' per day-pair and '{out}_g.txt' out, equal to
    the model API on the same inputs."""
    import os
    import subprocess
    import sys
    from wot_b200 import synthetic
    from wot_b200._anndata import AnnData
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    X, day, growth = synthetic.expression_matrix([260, 300, 280], n_genes=60, seed=31)
    ids = ["c%d" % i for i in range(len(day))]
    genes = ["g%d" % i for i in range(X.shape[1])]
    pd.DataFrame(X, index=ids, columns=genes).to_csv(tmp_path / "matrix.txt", sep="\t", index_label="id")
    pd.DataFrame({"day": day}, index=ids).to_csv(tmp_path / "days.txt", sep="\t", index_label="id")
    pd.DataFrame({"cell_growth_rate": growth}, index=ids).to_csv(tmp_path / "growth.txt", sep="\t", index_label="id")
    cmd = [sys.executable, "-m", "wot_b200", "optimal_transport", "--matrix", str(tmp_path / "matrix.txt"),
           "--cell_days", str(tmp_path / "days.txt"), "--cell_growth_rates", str(tmp_path / "growth.txt"),
           "--growth_iters", "2", "--local_pca", "12", "--epsilon", "0.06", "--format", "npz", "--out", str(tmp_path / "tm")]
    run = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    names = sorted(n for n in os.listdir(tmp_path) if n.startswith("tm_"))
    assert names == ["tm_0.0_1.0.npz", "tm_1.0_2.0.npz", "tm_g.txt"]
    # the text round trip keeps float64 to repr precision, so the API on the parsed matrix gives the same maps
    parsed = pd.read_csv(tmp_path / "matrix.txt", sep="\t", index_col=0)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=ids)
    model = ot.OTModel(AnnData(parsed.values, obs, pd.DataFrame(index=genes)), growth_iters=2, local_pca=12, epsilon=0.06)
    for t0, t1 in ((0.0, 1.0), (1.0, 2.0)):
        want = model.compute_transport_map(t0, t1)
        got = np.load(tmp_path / ("tm_%s_%s.npz" % (t0, t1)), allow_pickle=True)
        np.testing.assert_allclose(got["X"], np.asarray(want.X), rtol=1e-9, atol=0)
        assert list(got["obs_columns"]) == ["g0", "g1", "g2"]
        np.testing.assert_allclose(got["obs_values"], want.obs.values, rtol=1e-9)
    g = pd.read_csv(tmp_path / "tm_g.txt", sep="\t", index_col="id")
    assert list(g.columns) == ["g0", "g1", "g2"] and len(g) == 260 + 300


def test_gpu_pca_rank_deficient_input_takes_the_exact_gpu_solver(caplog):
    """An expression matrix of rank 20 cannot feed 40 independent test vectors: the Cholesky-QR of the range finder
    reports it, backend='gpu' raises, backend='auto' logs a warning and returns the exact GPU solver's result: the
    components of the non-zero singular values equal scikit-learn's, nothing is computed on the CPU."""
    import logging
    from wot_b200.ot import util
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1300, 20)) @ rng.standard_normal((20, 300))
    m1, m2 = X[:600], X[600:]
    with pytest.raises(ValueError, match="not positive definite"):
        util.compute_pca(m1, m2, 30, backend="gpu")
    with caplog.at_level(logging.WARNING, logger="wot"):
        p1, p2, pca, _ = util.compute_pca(m1, m2, 30)
    assert "rank-deficient" in caplog.text and isinstance(pca, util.LocalPCA)
    assert p1.shape == (600, 30) and p2.shape == (700, 30)
    s1, s2, ref, _ = util.compute_pca_sklearn(m1, m2, 30)
    np.testing.assert_allclose(pca.singular_values_[:18], ref.singular_values_[:18], rtol=1e-9)
    np.testing.assert_allclose(np.vstack([p1, p2])[:, :18], np.vstack([s1, s2])[:, :18], rtol=0, atol=1e-8)
    assert np.all(np.isfinite(p1)) and np.all(np.isfinite(p2))


@pytest.mark.parametrize("cells,genes,k,solver", [([60, 70], 2000, 30, "covariance_eigh"), ([100, 120], 400, 30, "full"),
                                                  ([15, 18], 200, 30, "full"), ([15, 18], 600, 30, "covariance_eigh"), ([300, 350], 36, 30, "full"),
                                                  ([400, 500], 9500, 30, "covariance_eigh")])
def test_gpu_pca_exact_solver_shapes_vs_sklearn(cells, genes, k, solver):
    """The shapes where scikit-learn's svd_solver='auto' leaves its randomized solver (few cells: covariance_eigh; tiny
    or k close to the rank: full LAPACK SVD): compute_pca stays on the GPU (exact path) and equals scikit-learn."""
    from oracle import wot_oracle
    from wot_b200 import synthetic
    from wot_b200.ot import util
    X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=11)
    m1, m2 = X[day == 0], X[day == 1]
    kk = min(k, sum(cells))
    assert util.sklearn_solver_choice(genes, sum(cells), kk) == solver
    p1, p2, pca, mu = util.compute_pca_sklearn(m1, m2, k)
    q1, q2, gpca, mu2 = util.compute_pca(m1, m2, k)
    assert isinstance(gpca, util.LocalPCA)
    np.testing.assert_allclose(mu2, mu, rtol=0, atol=1e-13)
    np.testing.assert_allclose(gpca.singular_values_, pca.singular_values_, rtol=1e-9)
    np.testing.assert_allclose(gpca.mean_, pca.mean_, rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.vstack([q1, q2]), np.vstack([p1, p2]), rtol=0, atol=1e-7)
    want = wot_oracle.compute_default_cost_matrix(p1, p2, np.diag(pca.singular_values_))
    got = wot_oracle.compute_default_cost_matrix(q1, q2, np.diag(gpca.singular_values_))
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)
