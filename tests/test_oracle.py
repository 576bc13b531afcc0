"""The oracle (oracle/wot_oracle.py) against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from oracle import wot_oracle as orc
from tests.helpers import DEFAULTS, pair_cost, max_rel_err


def _run(name, C, G, gap="dense", **over):
    info = orc.SolveInfo()
    params = dict(DEFAULTS, **over)
    tmap = getattr(orc, name)(C=C, G=G, info=info, gap=gap, **params)
    return tmap, info


def test_reference_golden_case_identity(golden):
    """/root/reference/tests/test_transport.py:20-32 -- 3x3 cost of 0/100, eps=0.01 -> identity, atol 0.01."""
    g = golden("ref_3x3")
    for solver, tag in (("optimal_transport_duality_gap", "dg"), ("transport_stablev2", "fx")):
        tmap, _ = _run(solver, g["C"], np.ones(3), epsilon=0.01)
        assert np.allclose(tmap, np.eye(3), atol=0.01, rtol=0)
        np.testing.assert_allclose(tmap, g[tag + "_tmap"], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("tag", ["small", "mid"])
@pytest.mark.parametrize("gap", ["dense", "marginal"])
def test_default_solver_matches_reference(golden, tag, gap):
    g = golden("dg_" + tag)
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    np.testing.assert_allclose([C.sum(), C[0, 0], C[-1, -1]], g["C_checksum"], rtol=1e-12)
    tmap, info = _run("optimal_transport_duality_gap", C, G, gap=gap)
    assert info.iters == int(g["dg_iters"])
    assert list(info.batches) == list(g["dg_batches"])
    tol = 1e-12 if gap == "dense" else 1e-9
    assert max_rel_err(tmap, g["dg_tmap"]) < tol
    np.testing.assert_allclose(info.f, g["dg_f"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(info.g, g["dg_g"], rtol=0, atol=1e-11)


def test_parameter_variations_match_reference(golden):
    g = golden("dg_variations")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    over = {
        "eps01": dict(epsilon=0.01), "lam10_100": dict(lambda1=10, lambda2=100),
        "loose": dict(epsilon=0.1, lambda1=0.1, lambda2=1), "batch7": dict(batch_size=7),
        "tau1_2": dict(tau=1.2), "tau2_eps02": dict(tau=2.0, epsilon=0.02), "maxiter37": dict(max_iter=37),
        "eps0_2": dict(epsilon0=2.0), "tol1e-5": dict(tolerance=1e-5),
    }
    assert sorted(over) == list(g["names"])
    for tag, kw in over.items():
        for gap in ("dense", "marginal"):
            tmap, info = _run("optimal_transport_duality_gap", C, G, gap=gap, **kw)
            assert info.iters == int(g[tag + "_iters"]), tag
            assert list(info.batches) == list(g[tag + "_batches"]), tag
            assert max_rel_err(tmap, g[tag + "_tmap"]) < 1e-9, tag
            np.testing.assert_allclose(info.f, g[tag + "_f"], rtol=0, atol=1e-10)
    assert _run("optimal_transport_duality_gap", C, G, tau=1.2)[1].tau_absorptions > 0
    assert _run("optimal_transport_duality_gap", C, G, max_iter=37)[1].hit_max_iter


def test_fixed_iters_matches_reference(golden):
    g = golden("fixed_iters")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    for tag, kw in (("default", {}), ("short", dict(scaling_iter=330, extra_iter=40, inner_iter_max=50)),
                    ("tau1_5", dict(scaling_iter=400, extra_iter=50, tau=1.5))):
        tmap, info = _run("transport_stablev2", C, G, **kw)
        assert max_rel_err(tmap, g[tag + "_tmap"]) < 1e-11, tag
        np.testing.assert_allclose(info.f, g[tag + "_f"], rtol=0, atol=1e-11)


def test_growth_loop_matches_reference(golden):
    g = golden("growth3")
    n0, n1, seed = (int(v) for v in g["shape"])
    C, G = pair_cost(n0, n1, seed)
    tmap, learned = orc.compute_transport_matrix(orc.optimal_transport_duality_gap,
                                                 **dict(DEFAULTS, growth_iters=3, C=C, G=G.copy()))
    assert max_rel_err(tmap, g["tmap"]) < 1e-11
    np.testing.assert_allclose(np.array(learned), g["learned"], rtol=1e-11)


def test_default_cost_matches_reference(golden):
    from wot_b200 import synthetic
    g = golden("cost_default")
    n0, n1, seed = (int(v) for v in g["shape"])
    x0, x1, _ = synthetic.day_pair_coords(n0, n1, d=30, seed=seed)
    np.testing.assert_allclose(orc.compute_default_cost_matrix(x0, x1, np.diag(g["sv"])), g["C"], rtol=1e-12)
    np.testing.assert_allclose(orc.compute_default_cost_matrix(x0[:, :7], x1[:, :7]), g["C_plain7"], rtol=1e-12)


def test_marginal_gap_identity():
    """SURVEY 8 a-note: primal/dual from marginals == the reference's dense arithmetic."""
    rng = np.random.default_rng(5)
    C, G = pair_cost(40, 50, 21)
    eps, l1, l2 = 0.05, 1.0, 50.0
    u, v = rng.normal(0, .1, 40), rng.normal(0, .1, 50)
    a, b = np.exp(rng.normal(0, .2, 40)), np.exp(rng.normal(0, .2, 50))
    K = np.exp((u[:, None] - C + v[None, :]) / eps)
    K0 = np.exp(-C / eps)
    R = (K.T * a).T * b
    dx, dy = np.full(40, 1 / 40), np.full(50, 1 / 50)
    q = np.full(50, G.mean())
    pri = orc.primal_dense(C, K0, R, dx, dy, G, q, eps, l1, l2)
    dua = orc.dual_dense(K0, R, dx, dy, G, q, a * np.exp(u / eps), b * np.exp(v / eps), eps, l1, l2)
    pri2, dua2 = orc.gap_from_marginals(R.sum(1), R.sum(0), u + eps * np.log(a), v + eps * np.log(b),
                                        K0.sum(), G, q, eps, l1, l2)
    assert abs(pri - pri2) < 1e-13 * abs(pri) and abs(dua - dua2) < 1e-13 * abs(dua)


@pytest.mark.parametrize("cells,genes,k", [([600, 700], 300, 30), ([450, 420], 1479, 30), ([300, 280], 700, 10)])
def test_pca_restatement_matches_sklearn(cells, genes, k):
    """oracle/pca_oracle.py (randomized range finder with Cholesky-QR normalisation, the form the CUDA path runs)
    against the reference's own call, sklearn PCA(k, random_state=58951).fit(x.T) (wot/ot/util.py:240-255), on the
    installed scikit-learn: singular values and the default cost matrix built from the coordinates
    (ot_model.py:242-253; sign-invariant) agree to roundoff.  Covers both orientations of transpose='auto'."""
    from oracle import pca_oracle, wot_oracle
    from wot_b200 import synthetic
    from wot_b200.ot import util
    X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=3)
    m1, m2 = X[day == 0], X[day == 1]
    assert pca_oracle.solver_choice(genes, sum(cells), k) == util.sklearn_solver_choice(genes, sum(cells), k) == "randomized"
    p1, p2, pca, mu = util.compute_pca_sklearn(m1, m2, k)
    q1, q2, sv, mu2 = pca_oracle.compute_pca(m1, m2, k)
    np.testing.assert_allclose(mu2, mu, rtol=0, atol=1e-14)
    np.testing.assert_allclose(sv, pca.singular_values_, rtol=1e-12)
    want = wot_oracle.compute_default_cost_matrix(p1, p2, np.diag(pca.singular_values_))
    got = wot_oracle.compute_default_cost_matrix(q1, q2, np.diag(sv))
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("tag", ["tall", "wide"])
def test_pca_restatement_matches_reference_golden(golden, tag):
    """oracle/pca_oracle.py against tests/golden/pca_randomized.npz, produced by the UNMODIFIED reference's
    wot.ot.compute_pca (make_golden.py pca_cases): loadings up to sign, singular values, gene means."""
    from oracle import pca_oracle
    from wot_b200 import synthetic
    g = golden("pca_randomized")
    cells, genes, k = [int(c) for c in g[tag + "_cells"]], int(g[tag + "_genes"]), int(g[tag + "_k"])
    X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=int(g["seed"]))
    q0, q1, sv, mean = pca_oracle.compute_pca(X[day == 0], X[day == 1], k)
    np.testing.assert_allclose(sv, g[tag + "_sv"], rtol=1e-11)
    np.testing.assert_allclose(mean, g[tag + "_mean"], rtol=0, atol=1e-13)
    want = np.vstack([g[tag + "_pca0"], g[tag + "_pca1"]])
    got = np.vstack([q0, q1])
    sign = np.sign((got * want).sum(axis=0))
    np.testing.assert_allclose(got * sign, want, rtol=0, atol=1e-9)
