"""CPU tests of the host side: C-ABI surface, configuration plumbing, CLI parsing, error behaviour."""
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function include/wot_b200.h declares is exported by the built library and bound in _lib."""
    from wot_b200 import _build, _lib
    _build.build_library()
    header = open(os.path.join(ROOT, "include", "wot_b200.h")).read()
    declared = set(re.findall(r"\b(wotb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, name
    assert set(_lib.SIGNATURES) == declared
    assert b"sm_100a" in lib.wotb_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wot_b200 import _lib, ot
    with pytest.raises(_lib.WotB200Error, match="no CPU fallback"):
        ot.optimal_transport_duality_gap(C=np.ones((3, 3)), G=np.ones(3), lambda1=1, lambda2=50, epsilon=0.05,
                                         batch_size=5, tolerance=1e-8, tau=1e4, epsilon0=1, max_iter=1e7)


def test_params_struct_layout_matches_header():
    import ctypes
    from wot_b200 import _lib
    assert ctypes.sizeof(_lib.Params) == 7 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.Info) == 8 + 6 * 4 + 2 * 4 + 6 * 8 + 2 * 8
    p = _lib.make_params(tau=None, fuse=False, unknown_key=3)
    assert np.isnan(p.tau) and p.reserved == 1


def _adata(n_days=3, cells=5, genes=8, seed=0):
    from wot_b200._anndata import AnnData
    rng = np.random.default_rng(seed)
    day = np.repeat(np.arange(n_days, dtype=float), cells)
    obs = pd.DataFrame({"day": day}, index=["c%d" % i for i in range(len(day))])
    return AnnData(rng.random((len(day), genes)), obs, pd.DataFrame(index=["g%d" % i for i in range(genes)]))


def test_otmodel_defaults_and_errors():
    from wot_b200 import ot
    m = ot.OTModel(_adata(), epsilon=0.1)
    assert m.ot_config["epsilon"] == 0.1 and m.ot_config["lambda2"] == 50 and m.ot_config["tau"] == 10000
    assert m.ot_config["local_pca"] == 0          # 30 > 8 genes -> PCA disabled, like ot_model.py:105-109
    assert m.timepoints == [0.0, 1.0, 2.0]
    assert m.solver is ot.optimal_transport_duality_gap
    assert ot.OTModel(_adata(), solver="fixed_iters").solver is ot.transport_stablev2
    with pytest.raises(ValueError, match="Unknown solver"):
        ot.OTModel(_adata(), solver="nope")
    bad = _adata()
    bad.obs = bad.obs.rename(columns={"day": "when"})
    with pytest.raises(ValueError, match="Days information not available"):
        ot.OTModel(bad)
    m = ot.OTModel(_adata(), config=pd.DataFrame({"t0": [0.0], "t1": [1.0], "epsilon": [0.02]}))
    assert m.day_pairs == {(0.0, 1.0): {"epsilon": 0.02}}
    with pytest.raises(ValueError, match="not present in day_pairs"):
        m.compute_transport_map(1.0, 2.0)
    assert ot.OTModel(_adata()).compute_transport_map(0.0, 7.0) is None     # no cells at t1 (ot_model.py:287-292)


def test_configuration_parsers(tmp_path):
    from wot_b200 import ot
    per_t = pd.DataFrame({"t": [0, 1, 2], "epsilon": [0.1, 0.2, 0.4]})
    assert ot.parse_configuration(per_t) == {(0.0, 1.0): {"epsilon": pytest.approx(0.15)},
                                             (1.0, 2.0): {"epsilon": pytest.approx(0.3)}}
    assert ot.parse_configuration("t0,t1,lambda1;0,1,3;") == {(0, 1): {"lambda1": 3}}
    with pytest.raises(ValueError):
        ot.parse_configuration(pd.DataFrame({"x": [1]}))
    pf = tmp_path / "params.txt"
    pf.write_text("epsilon\t0.07\nlambda1\t2\n")
    assert ot.parse_parameter_file(str(pf)) == {"epsilon": 0.07, "lambda1": 2}
    m = ot.OTModel(_adata(), parameters=str(pf), epsilon=0.5)
    assert m.ot_config["epsilon"] == 0.07         # the parameter file wins over kwargs (ot_model.py:97-103)


def test_cli_parser_matches_reference_flags():
    from wot_b200.commands import create_parser
    args = create_parser().parse_args(["--matrix", "m.txt", "--cell_days", "d.txt"])
    assert (args.epsilon, args.lambda1, args.lambda2, args.growth_iters, args.local_pca) == (0.05, 1, 50, 1, 30)
    assert (args.tau, args.epsilon0, args.batch_size, args.scaling_iter, args.inner_iter_max) == (10000, 1, 5, 3000, 50)
    assert args.solver == "duality_gap" and args.format == "h5ad" and args.out == "./tmaps" and not args.no_overwrite
    assert args.day_field == "day" and args.growth_rate_field == "cell_growth_rate"
    args = create_parser().parse_args(["--matrix", "m", "--cell_days", "d", "--tolerance", "1e-6", "--no_overwrite",
                                       "--solver", "fixed_iters", "--kernel", "online"])
    assert args.tolerance == 1e-6 and args.no_overwrite and args.solver == "fixed_iters" and args.kernel == "online"


def test_initialize_ot_model_from_files(tmp_path):
    from wot_b200 import ot
    ad = _adata()
    mat = tmp_path / "matrix.txt"
    pd.DataFrame(ad.X, index=ad.obs.index, columns=ad.var.index).to_csv(mat, sep="\t", index_label="id")
    days = tmp_path / "days.txt"
    ad.obs[["day"]].to_csv(days, sep="\t", index_label="id")
    m = ot.initialize_ot_model(str(mat), cell_days=str(days), epsilon=0.03)
    assert m.timepoints == [0.0, 1.0, 2.0] and m.ot_config["epsilon"] == 0.03
    assert (m.matrix.obs["cell_growth_rate"] == 1.0).all()     # io.py:538 default
    with pytest.raises(ValueError, match="not found"):
        ot.initialize_ot_model(str(mat), cell_days=str(tmp_path / "missing.txt"))


def test_synthetic_generators_are_deterministic():
    from wot_b200 import synthetic
    a = synthetic.day_pair_coords(50, 60, seed=3)
    b = synthetic.day_pair_coords(50, 60, seed=3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    pairs = synthetic.atlas_pairs()
    assert len(pairs) == 39 and all(5000 <= p[0] <= 20000 and 5000 <= p[1] <= 20000 for p in pairs)
    assert pairs[0][1] == pairs[1][0]            # consecutive days share a population


def test_sklearn_solver_policy_matches_installed_sklearn():
    """ot.util.sklearn_solver_choice decides whether compute_pca may take the GPU path (only where scikit-learn's
    svd_solver='auto' itself runs the randomized solver).  Checked against what the installed PCA actually picks."""
    import sklearn.decomposition
    from wot_b200.ot import util
    rng = np.random.default_rng(0)
    for n_samples, n_features, k in [(300, 620, 30), (60, 400, 10), (700, 580, 10), (520, 40, 30), (2000, 90, 30),
                                     (501, 300, 30), (40, 600, 38)]:
        pca = sklearn.decomposition.PCA(n_components=min(k, n_samples, n_features), random_state=58951)
        pca.fit(rng.standard_normal((n_samples, n_features)))
        assert util.sklearn_solver_choice(n_samples, n_features, min(k, n_samples, n_features)) == pca._fit_svd_solver, \
            (n_samples, n_features, k, pca._fit_svd_solver)


def test_kernel_policy_and_day_slices():
    from wot_b200.ot import optimal_transport as wot_ot
    from wot_b200.ot.ot_model import _rows_key
    assert wot_ot.resolve_kernel("auto", 1000, 1000, 30, 0.05) == "online"
    assert wot_ot.resolve_kernel("auto", 1000, 1000, 30, 0.01) == "online"      # precise operands below eps 0.02
    assert wot_ot.resolve_kernel("auto", 1000, 1000, 47, 0.05) == "stored"      # beyond the tcgen05 K budget
    assert wot_ot.resolve_kernel("stored", 1000, 1000, 30, 0.05) == "stored"
    assert wot_ot._kernel_id("online_simt") == (wot_ot._lib.KERNEL_ONLINE, True, None)
    assert wot_ot._kernel_id("online_precise") == (wot_ot._lib.KERNEL_ONLINE, False, True)
    assert wot_ot._lib.make_params(online_precise=True).reserved & 4
    assert wot_ot._lib.make_params(online_precise=False).reserved & 8
    with pytest.raises(ValueError):
        wot_ot._kernel_id("fast")
    mask = np.array([False, True, True, True, False])
    assert _rows_key(mask) == slice(1, 4)                                        # one run of rows: a view
    scattered = np.array([True, False, True, False, False])
    assert _rows_key(scattered) is scattered                                     # otherwise the mask itself
    assert _rows_key(np.zeros(3, dtype=bool)) is not None


def test_cli_additive_flags():
    from wot_b200.commands import optimal_transport as cli
    args = cli.create_parser().parse_args(["--matrix", "m.txt", "--cell_days", "d.txt", "--streams", "3", "--kernel", "auto"])
    assert args.streams == 3 and args.kernel == "auto" and args.format == "h5ad" and args.out == "./tmaps"


def test_pinned_block_lives_as_long_as_any_view():
    """A pooled page-locked block goes back to the pool only when the LAST array viewing it is gone, including the
    plain-ndarray views NumPy derives (np.asarray, slices, DataFrame): a second solve must not overwrite a map the
    caller still holds."""
    import ctypes
    import gc

    import pandas as pd
    from wot_b200 import _pinned

    class Block:
        def __init__(self, n):
            self.nbytes, self.buf = n, (ctypes.c_char * n)()

    made = []

    def alloc(n):
        made.append(Block(n))
        return made[-1]

    old_alloc, old_free = _pinned._alloc, list(_pinned._free)
    _pinned._alloc, _pinned._free[:] = alloc, []
    try:
        a = _pinned.empty((4, 5), np.float64)
        a[...] = 1.5
        plain, part, frame = np.asarray(a), a[1:3], pd.DataFrame(a)
        del a
        gc.collect()
        assert len(_pinned._free) == 0                     # still viewed
        b = _pinned.empty((4, 5), np.float64)              # the "second solve": must get a different block
        b[...] = 7.0
        assert len(made) == 2 and plain[0, 0] == 1.5 and part[0, 0] == 1.5 and frame.iloc[3, 4] == 1.5
        del plain, part, frame
        gc.collect()
        assert len(_pinned._free) == 1                     # now recycled
        c = _pinned.empty((2, 2), np.float64)
        assert len(made) == 2                              # served from the pool
        del b, c
    finally:
        _pinned._alloc, _pinned._free[:] = old_alloc, old_free
