"""The built-in .h5ad writer / HDF5 reader (wot_b200/h5ad.py): the output layout of the transport-map path,
/root/reference/wot/ot/ot_model.py:195 -> wot/io/io.py:447, read back the way
/root/reference/wot/tmap/transport_map_model.py:668-673 (file-name pattern) and :709-721 (h5py access) do."""
import os
import re
import struct

import numpy as np
import pandas as pd
import pytest

from wot_b200 import h5ad


def _real_hdf5_file():
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return path if os.path.exists(path) else None


def test_reader_on_a_file_written_by_libhdf5():
    """Pins the independent reader on a file produced by the real HDF5 library (MATLAB 7.3 = HDF5 1.6 behind a
    512-byte user block; shipped with scipy's test data): superblock, symbol-table group, B-tree, local heap,
    version-1 object header, contiguous float64 dataset, fixed-length string attribute."""
    path = _real_hdf5_file()
    if path is None:
        pytest.skip("scipy test data not installed")
    with h5ad.H5Reader(path) as f:
        assert f.base == 512 and f.root.keys() == ["testdouble"]
        ds = f["testdouble"]
        assert ds.attrs == {"MATLAB_class": "double"}
        np.testing.assert_allclose(ds.read().ravel(), np.arange(9) * np.pi / 4, rtol=1e-15)


def _maps(n_i, n_j, n_cols, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    X = rng.random((n_i, n_j)).astype(dtype)
    obs = ["cell_%d" % i for i in range(n_i)]
    var = ["t1_%d" % j for j in range(n_j)]
    cols = [("g%d" % k, rng.random(n_i)) for k in range(n_cols)]
    return X, obs, cols, var


@pytest.mark.parametrize("shape,n_cols,dtype", [((37, 53), 4, np.float64), ((1, 1), 1, np.float64), ((200, 3), 0, np.float32),
                                                ((9000, 2), 2, np.float64), ((5, 70000), 31, np.float32)])
def test_round_trip_is_bit_exact(tmp_path, shape, n_cols, dtype):
    """X, ids and growth columns come back bit for bit; index sets larger than one global heap collection (8192
    strings) and more columns than one default symbol-table node holds are covered."""
    X, obs, cols, var = _maps(*shape, n_cols, dtype=dtype)
    path = str(tmp_path / "tmaps_7.0_7.5.h5ad")
    eof = h5ad.write_h5ad(path, X, obs, cols, var)
    assert os.path.getsize(path) == eof
    got = h5ad.read_h5ad(path)
    assert got["X"].dtype == X.dtype and np.array_equal(got["X"], X)
    assert list(got["obs_index"]) == obs and list(got["var_index"]) == var
    assert list(got["obs"]) == [c[0] for c in cols]
    for name, values in cols:
        assert np.array_equal(got["obs"][name], values)
    assert got["attrs"]["encoding-type"] == "anndata"


def test_layout_is_what_the_reference_reader_expects(tmp_path):
    """transport_map_model.py:709-721 with h5py calls replaced by the reader: `_index` attribute, id datasets under
    /obs and /var, dense /X; :668-673: the file-name pattern."""
    X, obs, cols, var = _maps(20, 30, 3)
    obs[3] = "AAACCTGAGCTAGTTC-1_déjà"          # non-ASCII ids survive (UTF-8 variable-length strings)
    path = str(tmp_path / "tmaps_10.0_10.5.h5ad")
    h5ad.write_h5ad(path, X, obs, cols, var)
    day_regex = r"([0-9]*\.?[0-9]+)"
    m = re.compile("tmaps" + r"_{}_{}[\.h5ad|\.loom]".format(day_regex, day_regex)).match(os.path.basename(path))
    assert m is not None and float(m.group(1)) == 10.0 and float(m.group(2)) == 10.5
    with h5ad.H5Reader(path) as f:
        assert sorted(f.root.keys()) == ["X", "obs", "var"]
        o, v = f["/obs"], f["/var"]
        obs_key, var_key = o.attrs.get("_index", "index"), v.attrs.get("_index", "index")
        rids = o[obs_key].read().astype(str)
        cids = v[var_key].read().astype(str)
        assert list(rids) == obs and list(cids) == var
        assert list(o.attrs["column-order"]) == ["g0", "g1", "g2"]
        assert o.attrs["encoding-type"] == "dataframe" and f["X"].attrs["encoding-type"] == "array"
        assert f["X"].shape == (20, 30)
        np.testing.assert_array_equal(f["obs/g2"].read(), cols[2][1])


def test_file_structure_invariants(tmp_path):
    """Things libhdf5 checks when it opens a file: signature, end-of-file address, 8-byte aligned structures whose
    signatures sit where the pointers say, the matrix data contiguous at the recorded address."""
    X, obs, cols, var = _maps(11, 13, 2, seed=3)
    path = str(tmp_path / "t.h5ad")
    h5ad.write_h5ad(path, X, obs, cols, var)
    raw = open(path, "rb").read()
    assert raw[:8] == h5ad.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == h5ad.UNDEF and drv == h5ad.UNDEF and eof == len(raw)
    _, root_hdr, cache, _, tree, heap = struct.unpack_from("<QQIIQQ", raw, 56)
    assert cache == 1 and raw[tree:tree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP" and raw[root_hdr] == 1
    assert root_hdr % 8 == 0 and tree % 8 == 0 and heap % 8 == 0
    for m in re.finditer(b"GCOL", raw):                   # every global heap collection: version 1, size >= 4096
        at = m.start()
        if at % 8 == 0 and raw[at + 4] == 1:
            size, = struct.unpack_from("<Q", raw, at + 8)
            assert size >= 4096 and at + size <= len(raw)
    tail = np.frombuffer(raw[-X.nbytes:], dtype=np.float64).reshape(X.shape)
    assert np.array_equal(tail, X)


def test_anndata_write_and_async_writer(tmp_path):
    from wot_b200 import io as wio
    from wot_b200._anndata import AnnData, HAVE_ANNDATA
    if HAVE_ANNDATA:
        pytest.skip("anndata installed: AnnData.write is anndata's own")
    X, obs, cols, var = _maps(40, 25, 3, seed=9)
    ad = AnnData(X, pd.DataFrame({n: v for n, v in cols}, index=obs), pd.DataFrame(index=var))
    paths = [str(tmp_path / ("tm_%d.0_%d.5" % (k, k))) for k in range(4)]
    with h5ad.AsyncWriter(depth=2) as w:
        for p in paths:
            w.submit(lambda p=p: wio.write_dataset(ad, p, output_format="h5ad"))
    for p in paths:
        got = h5ad.read_h5ad(p + ".h5ad")
        assert np.array_equal(got["X"], X) and list(got["obs_index"]) == obs and list(got["obs"]) == ["g0", "g1", "g2"]
    assert ad.T.shape == (25, 40) and list(ad.T.obs.index) == var
    with pytest.raises(OSError):
        with h5ad.AsyncWriter() as w:
            w.submit(lambda: wio.write_dataset(ad, str(tmp_path / "no_such_dir" / "x"), output_format="h5ad"))


def test_model_from_directory_of_written_maps(tmp_path):
    """The consumer side of the output layout (transport_map_model.py:652-732): a directory of '{prefix}_{t0}_{t1}.h5ad'
    files written here becomes a model with the reference's meta table (every cell id with its day), the maps opened
    lazily with ids only; unrelated files and other prefixes are ignored; an empty match raises like the reference."""
    import numpy as np
    import pandas as pd

    from wot_b200 import _lib
    from wot_b200 import h5ad as h5
    from wot_b200.tmap import ImplicitTransportMapModel, StoredTransportMap
    rng = np.random.default_rng(3)
    sizes = [5, 7, 4]
    ids = [["d%d_%d" % (k, i) for i in range(n)] for k, n in enumerate(sizes)]
    mats = {}
    for k, (t0, t1) in enumerate([(0.0, 1.5), (1.5, 2.0)]):
        X = rng.random((sizes[k], sizes[k + 1]))
        mats[(t0, t1)] = X
        h5.write_h5ad(str(tmp_path / ("tm_%s_%s.h5ad" % (t0, t1))), X, ids[k], [("g0", np.ones(sizes[k]))], ids[k + 1])
    (tmp_path / "tm_g.txt").write_text("id\tg0\n")
    h5.write_h5ad(str(tmp_path / "other_0.0_1.5.h5ad"), mats[(0.0, 1.5)], ids[0], [], ids[1])
    model = ImplicitTransportMapModel.from_directory(str(tmp_path / "tm"))
    assert model.timepoints == [0.0, 1.5, 2.0] and sorted(model.tmaps) == [(0.0, 1.5), (1.5, 2.0)]
    assert list(model.meta.index) == sum(ids, []) and list(model.meta["day"]) == [0.0] * 5 + [1.5] * 7 + [2.0] * 4
    m = model.tmaps[(0.0, 1.5)]
    assert isinstance(m, StoredTransportMap) and m.shape == (5, 7) and list(m.obs.columns) == ["g0"]
    pops = model.population_from_ids(ids[1][:3], at_time=1.5, names=["A"])
    assert pops[0].p.sum() == 3
    import torch
    if torch.cuda.is_available():
        got = model.pull_back(*pops, normalize=False)
        np.testing.assert_allclose(got.p, mats[(0.0, 1.5)] @ pops[0].p, rtol=1e-13)
    else:
        with pytest.raises(_lib.WotB200Error, match="no CPU fallback"):
            model.pull_back(*pops)
    with pytest.raises(ValueError, match="No transport maps found"):
        ImplicitTransportMapModel.from_directory(str(tmp_path / "nothing"))


def test_model_from_directory_reads_npz_maps_too(tmp_path):
    """The 'npz' output format (usable without anndata / h5py) is read back by from_directory like the .h5ad files."""
    import numpy as np
    import pandas as pd

    from wot_b200 import io as wio
    from wot_b200._anndata import AnnData
    from wot_b200.tmap import ImplicitTransportMapModel
    rng = np.random.default_rng(5)
    ids = [["a%d" % i for i in range(4)], ["b%d" % i for i in range(6)]]
    X = rng.random((4, 6))
    obs = pd.DataFrame({"g0": np.ones(4), "g1": np.arange(4.0)}, index=ids[0])
    wio.write_dataset(AnnData(X, obs, pd.DataFrame(index=ids[1])), str(tmp_path / "maps_3_4.5"), output_format="npz")
    model = ImplicitTransportMapModel.from_directory(str(tmp_path / "maps"))
    assert model.timepoints == [3.0, 4.5]
    m = model.tmaps[(3.0, 4.5)]
    assert m.shape == (4, 6) and list(m.obs.index) == ids[0] and list(m.var.index) == ids[1]
    assert list(m.obs.columns) == ["g0", "g1"]
    assert list(model.meta["day"]) == [3.0] * 4 + [4.5] * 6
