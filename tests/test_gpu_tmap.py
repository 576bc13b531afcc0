"""Consumers of transport maps on implicit couplings: push-forward / pull-back of several
populations in one sweep, chaining over day-pairs (trajectories, fates, transition tables), glue, and the draw of
interpolate_with_ot -- all against DENSE products of the float64 ORACLE's couplings, with the reference's semantics
(/root/reference/wot/tmap/transport_map_model.py:40-143, :235-365; wot/tmap/util.py:74-94; wot/ot/util.py:109-147)."""
import numpy as np
import pandas as pd
import pytest

from tests.helpers import DEFAULTS, RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def chain():
    """Three consecutive day-pairs (4 days): implicit maps from the GPU solve, dense couplings from the oracle."""
    from oracle import wot_oracle as orc
    from wot_b200 import _lib, synthetic
    from wot_b200.ot import optimal_transport as wot_ot
    from wot_b200.tmap import ImplicitTransportMap, ImplicitTransportMapModel
    sizes = [230, 301, 187, 260]
    rng = np.random.default_rng(17)
    days = [synthetic.day_pair_coords(n, 2, d=12, seed=40 + k)[0] + 0.15 * k for k, n in enumerate(sizes)]
    ids = [["d%d_c%d" % (k, i) for i in range(n)] for k, n in enumerate(sizes)]
    maps, dense = {}, {}
    for k in range(3):
        x0, x1 = days[k], days[k + 1]
        growth = np.exp(rng.normal(0, 0.2, len(x0)))
        want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, **DEFAULTS)
        _, _ = wot_ot.solve_coords(x0, x1, growth, _lib.SOLVER_DUALITY_GAP, want_tmap=False,
                                   **{k2: v for k2, v in DEFAULTS.items() if k2 != "growth_iters"})
        last = wot_ot.last_solve_info()
        info = last["infos"][-1]
        maps[(float(k), float(k + 1))] = ImplicitTransportMap(
            x0, x1, last["f"], last["g"], last["median"], info["eps_final"], info["out_scale"],
            obs=pd.DataFrame(index=ids[k]), var=pd.DataFrame(index=ids[k + 1]), t0=float(k), t1=float(k + 1))
        dense[(float(k), float(k + 1))] = want
    return ImplicitTransportMapModel(maps), dense, sizes, ids


def test_multi_population_apply_vs_oracle_products(chain):
    """All populations in ONE call; signed weights allowed; shapes that are not multiples of the 64-wide tile."""
    model, dense, sizes, _ = chain
    rng = np.random.default_rng(1)
    m, T = model.tmaps[(1.0, 2.0)], dense[(1.0, 2.0)]
    for n_pop in (1, 3, 8, 11):
        p = rng.random((n_pop, sizes[1]))
        p[0, ::3] = 0.0
        np.testing.assert_allclose(m.push_forward(p), p @ T, rtol=RTOL)
        q = rng.random((n_pop, sizes[2]))
        np.testing.assert_allclose(m.pull_back(q), (T @ q.T).T, rtol=RTOL)
    signed = rng.standard_normal((2, sizes[1]))
    np.testing.assert_allclose(m.push_forward(signed), signed @ T, rtol=1e-3, atol=1e-4 * np.abs(signed @ T).max())
    np.testing.assert_allclose(m.row_sums(), T.sum(axis=1), rtol=RTOL)
    np.testing.assert_allclose(m.col_sums(), T.sum(axis=0), rtol=RTOL)
    np.testing.assert_allclose(m.to_dense(), T, rtol=RTOL, atol=1e-12 * T.max())


def _ref_push(dense, times, p, i, j, normalize):
    while i < j:
        p = p @ dense[(times[i], times[i + 1])]
        if normalize:
            p = (p.T / np.sum(p, axis=1)).T
        i += 1
    return p


def _ref_pull(dense, times, p, i, j, normalize):
    while i > j:
        p = (dense[(times[i - 1], times[i])] @ p.T).T
        if normalize:
            p = (p.T / np.sum(p, axis=1)).T
        i -= 1
    return p


def test_chained_push_forward_pull_back(chain):
    from wot_b200.tmap import Population
    model, dense, sizes, _ = chain
    times = model.timepoints
    rng = np.random.default_rng(2)
    pops = [Population(1.0, rng.random(sizes[1]), "a"), Population(1.0, rng.random(sizes[1]), "b")]
    stack = np.vstack([p.p for p in pops])
    for normalize in (True, False):
        got = model.push_forward(*pops, to_time=3.0, normalize=normalize)
        np.testing.assert_allclose(np.vstack([g.p for g in got]), _ref_push(dense, times, stack, 1, 3, normalize), rtol=3 * RTOL)
        got = model.pull_back(*pops, to_time=0.0, normalize=normalize)
        np.testing.assert_allclose(np.vstack([g.p for g in got]), _ref_pull(dense, times, stack, 1, 0, normalize), rtol=3 * RTOL)
    assert model.push_forward(pops[0]).time == 2.0 and model.pull_back(pops[0]).time == 0.0
    with pytest.raises(ValueError):
        model.push_forward(Population(3.0, np.ones(sizes[3])))
    with pytest.raises(ValueError):
        model.pull_back(Population(0.0, np.ones(sizes[0])))
    with pytest.raises(ValueError):
        model.push_forward(pops[0], Population(2.0, np.ones(sizes[2])))


def test_trajectories_fates_transition_table(chain):
    """transport_map_model.py:105-143 (trajectories), :40-69 (fates), :71-103 (transition_table) on dense oracle maps."""
    from wot_b200.tmap import Population
    model, dense, sizes, ids = chain
    times = model.timepoints
    sets = [ids[2][:60], ids[2][60:130]]
    pops = model.population_from_ids(*sets, at_time=2.0, names=["A", "B"])
    # ---- trajectories: normalised populations, pulled back to day 0 and pushed forward to day 3 ----
    start = np.vstack([p.p / p.p.sum() for p in pops])
    blocks, cur = [start.T], start
    for i in (2, 1):
        cur = _ref_pull(dense, times, cur, i, i - 1, True)
        blocks.insert(0, cur.T)
    blocks.append(_ref_push(dense, times, start, 2, 3, True).T)
    traj = model.trajectories(pops)
    assert list(traj.columns) == ["A", "B"] and list(traj.index) == sum(ids, [])
    np.testing.assert_allclose(traj.values, np.concatenate(blocks), rtol=5 * RTOL)
    # ---- fates: unnormalised pull-backs of A, B and the completing 'Other' population, row-normalised ----
    other = 1.0 - np.clip(pops[0].p + pops[1].p, 0, 1)
    stack = np.vstack([pops[0].p, pops[1].p, other])
    blocks, cur = [stack.T], stack
    for i in (2, 1):
        cur = _ref_pull(dense, times, cur, i, i - 1, False)
        blocks.insert(0, cur.T)
    want = np.concatenate(blocks)
    want = want / want.sum(axis=1, keepdims=True)
    fates = model.fates(pops)
    assert list(fates.columns) == ["A", "B", "Other"] and len(fates) == sum(sizes[:3])
    np.testing.assert_allclose(fates.values, want, rtol=5 * RTOL, atol=1e-12)
    # ---- transition table from two day-0 populations to the day-2 populations ----
    starts = model.population_from_ids(ids[0][:100], ids[0][100:], at_time=0.0, names=["s0", "s1"])
    ends = [Population(2.0, stack[k], n) for k, n in enumerate(["A", "B", "Other"])]
    table = model.transition_table(starts, ends)
    end_p = _ref_pull(dense, times, stack, 2, 0, False)
    want = np.vstack([p.p for p in starts]) @ end_p.T
    np.testing.assert_allclose(table.values, want / want.sum(), rtol=5 * RTOL)


def test_glue_and_interpolate(chain):
    """glue_transport_maps (wot/tmap/util.py:74-94) as lazy composition and as a dense product; interpolate_with_ot
    (wot/ot/util.py:109-147) draws the same cell pairs as np.random.choice on the dense oracle coupling."""
    from wot_b200._anndata import AnnData
    from wot_b200.ot.util import interpolate_with_ot
    from wot_b200.tmap import glue_transport_maps
    model, dense, sizes, ids = chain
    m01, m12 = model.tmaps[(0.0, 1.0)], model.tmaps[(1.0, 2.0)]
    T01, T12 = dense[(0.0, 1.0)], dense[(1.0, 2.0)]
    glued = glue_transport_maps(m01, m12)
    assert glued.shape == (sizes[0], sizes[2])
    rng = np.random.default_rng(3)
    p = rng.random((3, sizes[0]))
    np.testing.assert_allclose(glued.push_forward(p), p @ (T01 @ T12), rtol=3 * RTOL)
    q = rng.random((2, sizes[2]))
    np.testing.assert_allclose(glued.pull_back(q), ((T01 @ T12) @ q.T).T, rtol=3 * RTOL)
    # dense AnnData maps: re-indexing of the intermediate day (:90-91) and a float64 product
    perm = rng.permutation(sizes[1])
    a = AnnData(T01, pd.DataFrame(index=ids[0]), pd.DataFrame(index=ids[1]))
    b = AnnData(T12[perm], pd.DataFrame(index=[ids[1][k] for k in perm]), pd.DataFrame(index=ids[2]))
    g = glue_transport_maps(a, b)
    np.testing.assert_allclose(np.asarray(g.X), T01 @ T12, rtol=1e-12)
    assert list(g.obs.index) == ids[0] and list(g.var.index) == ids[2]
    # ---- interpolate_with_ot ----
    genes = 7
    e0, e1 = rng.random((sizes[0], genes)), rng.random((sizes[1], genes))
    for frac in (0.5, 0.25):
        prob = T01 / np.power(T01.sum(axis=0), 1.0 - frac)
        prob = prob.flatten(order="C")
        prob = prob / prob.sum()
        np.random.seed(1234)
        choices = np.random.choice(sizes[0] * sizes[1], p=prob, size=4000)
        want = np.asarray([e0[c // sizes[1]] * (1 - frac) + e1[c % sizes[1]] * frac for c in choices])
        np.random.seed(1234)
        got = interpolate_with_ot(e0, e1, m01, frac, 4000)
        # the GPU's coupling differs from the oracle's by <= 1e-4 relative, so a draw that lands within that distance
        # of a cell boundary of the cumulative sum may pick the neighbouring pair: allow a handful of 4000
        same = np.all(np.isclose(got, want, rtol=0, atol=1e-12), axis=1)
        assert same.mean() >= 0.995, same.mean()
    with pytest.raises(TypeError):
        interpolate_with_ot(e0, e1, T01, 0.5, 10)


def test_model_from_directory_equals_implicit_model(tmp_path):
    """Files written by OTModel.compute_all_transport_maps (default '.h5ad') read back through
    ImplicitTransportMapModel.from_directory (the consumer side of the layout, transport_map_model.py:652-732): same
    meta table, and trajectories / fates equal those of the model built on the implicit couplings of the same solves."""
    from wot_b200 import ot, synthetic
    from wot_b200._anndata import AnnData
    from wot_b200.tmap import ImplicitTransportMapModel, StoredTransportMap
    X, day, growth = synthetic.expression_matrix([210, 260, 190], n_genes=120, seed=21)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    model = ot.OTModel(AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])])), growth_iters=2,
                       local_pca=15)
    model.compute_all_transport_maps(tmap_out=str(tmp_path / "tm"))
    stored = ImplicitTransportMapModel.from_directory(str(tmp_path / "tm"))
    implicit = ImplicitTransportMapModel.from_ot_model(model)
    assert stored.timepoints == implicit.timepoints == [0.0, 1.0, 2.0]
    assert list(stored.meta.index) == list(implicit.meta.index) and list(stored.meta["day"]) == list(implicit.meta["day"])
    assert all(isinstance(m, StoredTransportMap) for m in stored.tmaps.values())
    assert list(stored.tmaps[(0.0, 1.0)].obs.columns) == ["g0", "g1", "g2"]
    day1 = list(obs.index[day == 1.0])
    pops_s = stored.population_from_ids(day1[:50], day1[50:140], at_time=1.0, names=["A", "B"])
    pops_i = implicit.population_from_ids(day1[:50], day1[50:140], at_time=1.0, names=["A", "B"])
    np.testing.assert_allclose(stored.trajectories(pops_s).values, implicit.trajectories(pops_i).values, rtol=5 * RTOL)
    np.testing.assert_allclose(stored.fates(pops_s).values, implicit.fates(pops_i).values, rtol=5 * RTOL, atol=1e-12)
    StoredTransportMap.resident = 1                      # the second map evicts the first: results must not change
    try:
        np.testing.assert_allclose(stored.trajectories(pops_s).values, implicit.trajectories(pops_i).values, rtol=5 * RTOL)
    finally:
        StoredTransportMap.resident = 2
        StoredTransportMap._cache.clear()
