"""Host-side multi-process logic on CPU (gloo, world_size 2): unit sharding and the distributed form of
compute_all_transport_maps.  The solver here is the float64 oracle plugged into OTModel.solver (the reference's
de-facto plugin hook, ot_model.py:88-94) with caller-supplied cost matrices, so no GPU is touched."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_units_balances_and_keeps_order():
    from wot_b200.parallel import shard_units
    costs = [9, 1, 8, 2, 7, 3, 6, 4, 5]
    parts = shard_units(costs, 3)
    assert sorted(sum(parts, [])) == list(range(9))
    loads = [sum(costs[k] for k in p) for p in parts]
    assert max(loads) - min(loads) <= 2
    assert all(p == sorted(p) for p in parts)
    assert shard_units(costs, 1) == [list(range(9))]
    assert shard_units([], 4) == [[], [], [], []]


def _make_model():
    from oracle import wot_oracle as orc
    from wot_b200 import ot, synthetic
    from wot_b200._anndata import AnnData
    X, day, growth = synthetic.expression_matrix([30, 36, 33, 31], n_genes=40, seed=5)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = ot.OTModel(adata, growth_iters=2, local_pca=0)
    model.solver = lambda **kw: orc.optimal_transport_duality_gap(**kw)   # CPU plug-in solver
    costs = []
    for t0, t1 in ((0.0, 1.0), (1.0, 2.0), (2.0, 3.0)):
        a, b = X[day == t0], X[day == t1]
        costs.append(orc.compute_default_cost_matrix(a, b))
    return model, costs


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wot_b200 import parallel
        model, costs = _make_model()
        done = parallel.compute_all_transport_maps(model, tmap_out=os.path.join(out_dir, "tm"),
                                                   output_file_format="npz", cost_matrices=costs)
        assert len(done) >= 1
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_distributed_compute_all_transport_maps_matches_serial(tmp_path):
    import torch.multiprocessing as mp
    serial, dist_dir = tmp_path / "serial", tmp_path / "dist"
    serial.mkdir()
    dist_dir.mkdir()
    model, costs = _make_model()
    model.compute_all_transport_maps(tmap_out=str(serial / "tm"), output_file_format="npz", cost_matrices=costs)
    mp.spawn(_worker, args=(2, _free_port(), str(dist_dir)), nprocs=2, join=True)
    names = sorted(os.listdir(serial))
    assert names == sorted(os.listdir(dist_dir))
    assert "tm_g.txt" in names and "tm_0.0_1.0.npz" in names and len(names) == 4
    for n in names:
        if n.endswith(".npz"):
            a, b = np.load(serial / n, allow_pickle=True), np.load(dist_dir / n, allow_pickle=True)
            np.testing.assert_array_equal(a["X"], b["X"])
            np.testing.assert_array_equal(a["obs_values"], b["obs_values"])
        else:
            assert (serial / n).read_text() == (dist_dir / n).read_text()


def test_no_overwrite_skips_existing(tmp_path):
    model, costs = _make_model()
    out = str(tmp_path / "tm")
    model.compute_all_transport_maps(tmap_out=out, output_file_format="npz", cost_matrices=costs)
    stamp = os.path.getmtime(out + "_0.0_1.0.npz")
    model.solver = None   # would raise if any pair were recomputed
    model.compute_all_transport_maps(tmap_out=out, output_file_format="npz", cost_matrices=costs, overwrite=False)
    assert os.path.getmtime(out + "_0.0_1.0.npz") == stamp


# ---- parameter sweep (BASELINE.json configs[4]): dynamic queue over ranks ---------------------------
def _sweep_inputs():
    from wot_b200 import parallel, synthetic
    x0, x1, growth = synthetic.day_pair_coords(60, 70, d=5, seed=11)
    grid = parallel.sweep_grid(epsilons=(0.05, 0.1), lambda1s=(1, 10), lambda2s=(10, 50))
    return x0, x1, growth, grid


def _oracle_solve(x0, x1, G, growth_iters=1, kernel=None, **params):
    import time

    from oracle import wot_oracle as orc
    time.sleep(0.05)     # units are milliseconds long here: give the other rank time to reach the queue
    info = orc.SolveInfo()
    prm = dict(epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7, batch_size=5)
    prm.update(params)
    tmap = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=G, info=info, **prm)
    return {"iters": info.iters, "batches": list(info.batches), "rowsum": tmap.sum(axis=1)}


def _sweep_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import pickle

    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wot_b200 import parallel
        x0, x1, growth, grid = _sweep_inputs()
        res = parallel.parameter_sweep(x0, x1, growth, grid, solve=_oracle_solve)
        with open(os.path.join(out_dir, "sweep_%d.pkl" % rank), "wb") as fh:
            pickle.dump(res, fh)
    finally:
        dist.destroy_process_group()


def test_sweep_grid_is_the_64_setting_grid():
    from wot_b200 import parallel
    grid = parallel.sweep_grid()
    assert len(grid) == 64 and len({tuple(sorted(s.items())) for s in grid}) == 64
    assert grid[0] == dict(epsilon=0.01, lambda1=0.1, lambda2=1.0)
    q = parallel.WorkQueue([3, 1, 2])
    assert list(q) == [3, 1, 2]


@pytest.mark.timeout(300)
def test_parameter_sweep_two_ranks_matches_serial(tmp_path):
    import pickle

    import torch.multiprocessing as mp

    from wot_b200 import parallel
    x0, x1, growth, grid = _sweep_inputs()
    serial = parallel.parameter_sweep(x0, x1, growth, grid, solve=_oracle_solve)
    mp.spawn(_sweep_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    per_rank = [pickle.load(open(tmp_path / ("sweep_%d.pkl" % r), "rb")) for r in range(2)]
    ran = [0, 0]
    for k, want in enumerate(serial):
        assert want["setting"] == grid[k]
        for got in (per_rank[0][k], per_rank[1][k]):     # every rank holds every result, in grid order
            assert got["setting"] == grid[k] and got["iters"] == want["iters"] and got["batches"] == want["batches"]
            np.testing.assert_array_equal(got["rowsum"], want["rowsum"])
        ran[per_rank[0][k]["rank"]] += 1
    assert sum(ran) == len(grid) and min(ran) >= 1          # both ranks drew from the queue


# ---------------------------------------------------------------------------------------------------
# host-side plumbing of the row-sharded solve (wot_b200.parallel.DistComm / ThreadComm)
# ---------------------------------------------------------------------------------------------------
def test_thread_comm_gathers_in_rank_order_and_synchronises():
    """ThreadComm: the ranks are threads of one process; all_gather_object returns every rank's object in rank
    order on every rank, twice in a row (the slots are reusable), barrier() is a rendezvous."""
    import threading
    from wot_b200.parallel import ThreadComm
    world = 3
    comms = ThreadComm.make(world)
    out, order = [None] * world, []

    def run(r):
        c = comms[r]
        assert (c.rank, c.world, c.in_process) == (r, world, True)
        first = c.all_gather_object(("ptr", r * 100))
        second = c.all_gather_object(None if r != 1 else "err")
        order.append(r)
        c.barrier()
        assert len(order) == world                      # nobody passes the barrier before everybody arrived
        out[r] = (first, second)
        with pytest.raises(RuntimeError):
            c.all_reduce(None)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    for r in range(world):
        assert out[r] == ([("ptr", 0), ("ptr", 100), ("ptr", 200)], [None, "err", None])


def test_row_slices_tile_the_rows_in_256_row_blocks():
    """hi_rows(n, rank, world): the rows of the slice of 256-row blocks a rank owns (the split wotb_online_open makes);
    the slices tile [0, n) for every world size, empty slices included."""
    from wot_b200.parallel import hi_rows
    for n in (1, 255, 256, 257, 3000, 100000):
        for world in (1, 2, 3, 8):
            sizes = [hi_rows(n, r, world) for r in range(world)]
            assert sum(sizes) == n and all(s >= 0 for s in sizes)
            assert all(s % 256 == 0 for s in sizes[:-1] if s) or sum(1 for s in sizes if s % 256) <= 1


def _comm_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import pickle

    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wot_b200.parallel import DistComm
        c = DistComm()
        got = c.all_gather_object((None, bytes([rank]) * 64))     # what _PeerBuffers exchanges: (error, IPC handle)
        t = torch.full((5,), float(rank + 1), dtype=torch.float64)
        c.all_reduce(t)
        c.barrier()
        pickle.dump((c.rank, c.world, c.in_process, got, t.tolist()), open(os.path.join(out_dir, "comm_%d.pkl" % rank), "wb"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dist_comm_two_ranks_gloo(tmp_path):
    """DistComm over a gloo process group of two ranks: rank / world, the object gather of the 64-byte handles in
    rank order, the SUM all-reduce of the nccl exchange."""
    import pickle

    import torch.multiprocessing as mp
    mp.spawn(_comm_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        rank, world, in_process, got, summed = pickle.load(open(tmp_path / ("comm_%d.pkl" % r), "rb"))
        assert (rank, world, in_process) == (r, 2, False)
        assert got == [(None, b"\x00" * 64), (None, b"\x01" * 64)]
        assert summed == [3.0] * 5


def test_bind_host_to_gpu_is_harmless_without_nvml_device(monkeypatch):
    """bind_host_to_gpu never raises: without a usable NVML device (or with WOTB_NO_NUMA_BIND=1) it reports 0 CPUs
    bound and leaves the affinity alone."""
    from wot_b200 import parallel
    before = os.sched_getaffinity(0)
    monkeypatch.setenv("WOTB_NO_NUMA_BIND", "1")
    assert parallel.bind_host_to_gpu(0) == 0
    monkeypatch.delenv("WOTB_NO_NUMA_BIND")
    n = parallel.bind_host_to_gpu(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
