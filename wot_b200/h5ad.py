"""A self-contained .h5ad writer and reader for transport maps (no h5py / anndata needed).

The reference writes a transport map with `AnnData.write` (wot/ot/ot_model.py:195 -> wot/io/io.py:447) and reads
it back with h5py (wot/tmap/transport_map_model.py:709-721): `/X` dense [I, J], `/obs` a group whose `_index`
attribute names the dataset of row ids (plus the growth columns g0..gN), `/var` likewise for the column ids.
`anndata` and `h5py` are not installed in this image, so the file is produced directly: HDF5 in its classic,
most widely readable form (superblock version 0, version-1 object headers, symbol-table groups with one B-tree
node + local heap each, contiguous datasets, variable-length UTF-8 strings in global heap collections -- the same
structures libhdf5 1.8+ writes by default), with the attributes anndata >= 0.7 puts on an AnnData file
(`encoding-type`, `encoding-version`, `_index`, `column-order`).  The dense matrix is the last thing in the file and is streamed from the
caller's buffer (a pinned block the GPU wrote) with plain write() calls, which release the GIL:
`AsyncWriter` moves that to a background thread so the next solve overlaps the disk.

`read_h5ad` is an independent reader of that subset of HDF5 (it is pinned on a real libhdf5-written file in
tests/test_h5ad.py).  Format reference: "HDF5 File Format Specification Version 3.0", sections III (disk format
level 1: B-trees, heaps, symbol table nodes) and IV (object headers and messages).
"""
from __future__ import annotations

import os
import queue
import struct
import threading

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
_GROUP_INTERNAL_K = 16
_X_ALIGN = 4096


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


# ---------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------
def _dt_float(itemsize):
    """IV.A.2.d datatype message, class 1 (floating point), little endian IEEE."""
    if itemsize == 8:
        return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if itemsize == 4:
        return struct.pack("<B3BI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    raise ValueError("float32 or float64 only")


def _dt_fixed_string(n):
    # class 3, null-padded (1), UTF-8 (1 << 4)
    return struct.pack("<B3BI", 0x13, 0x11, 0x00, 0x00, n)


def _dt_vlen_string():
    # class 9: type = string (1), padding = null terminated (0), character set UTF-8 (1); base type: 1-byte string
    base = struct.pack("<B3BI", 0x13, 0x10, 0x00, 0x00, 1)
    return struct.pack("<B3BI", 0x19, 0x01, 0x01, 0x00, 16) + base


def _dataspace(shape):
    """IV.A.2.b version 1; rank 0 is a scalar."""
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attribute(name, dtype_msg, shape, raw):
    """IV.A.2.m attribute message, version 1 (name, datatype and dataspace each padded to 8 bytes)."""
    nm = name.encode("utf-8") + b"\x00"
    ds = _dataspace(shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dtype_msg), len(ds)) + _pad8(nm) + _pad8(dtype_msg) + _pad8(ds) + raw
    return _message(0x000C, body)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


class _File:
    """Metadata blob under construction: everything except the big matrix, addressed from 0."""

    def __init__(self):
        self.buf = bytearray()
        self.heap_strings = []      # pending (collection, index) assignments happen in flush_global_heap

    def tell(self):
        return len(self.buf)

    def put(self, data, align=8):
        pad = -len(self.buf) % align
        self.buf += b"\x00" * pad
        at = len(self.buf)
        self.buf += data
        return at

    def patch(self, at, data):
        self.buf[at:at + len(data)] = data

    # ---- global heap: variable-length strings ------------------------------------------------------------
    def vlen_strings(self, strings, per_collection=8192):
        """Writes the strings into global heap collections (III.E) and returns the 16-byte descriptors
        (length, collection address, object index) that make up the dataset / attribute data."""
        out = bytearray()
        enc = [s.encode("utf-8") for s in strings]
        for start in range(0, len(enc), per_collection):
            chunk = enc[start:start + per_collection]
            objs = bytearray()
            for k, e in enumerate(chunk):
                objs += struct.pack("<HH4xQ", k + 1, 1, len(e)) + _pad8(e)
            size = 16 + len(objs)
            total = max(4096, size)
            if 0 < total - size < 16:
                total = size + 16
            free = total - size
            if free:
                objs += struct.pack("<HH4xQ", 0, 0, free) + b"\x00" * (free - 16)
            addr = self.put(b"GCOL" + struct.pack("<B3xQ", 1, total) + bytes(objs))
            for k, e in enumerate(chunk):
                out += struct.pack("<IQI", len(e), addr, k + 1)
        return bytes(out)

    def str_attr(self, name, value):
        return _attribute(name, _dt_vlen_string(), (), self.vlen_strings([value]))

    def str_array_attr(self, name, values):
        if len(values) == 0:      # what h5py stores for an empty list: a zero-length float64 array
            return _attribute(name, _dt_float(8), (0,), b"")
        return _attribute(name, _dt_vlen_string(), (len(values),), self.vlen_strings(list(values)))

    # ---- datasets -----------------------------------------------------------------------------------------
    def dataset(self, dtype_msg, shape, raw=None, data_addr=None, nbytes=None, attrs=()):
        """Contiguous dataset; `raw` is stored right behind the header, or (`data_addr`, `nbytes`) point at
        storage written elsewhere.  Returns the object header address."""
        if raw is not None:
            nbytes = len(raw)
        # fill value message v2 as libhdf5 writes it for a default dataset: late allocation, fill "if set", default value
        fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)
        layout_at_msgs = [
            _message(0x0001, _dataspace(shape)),
            _message(0x0003, dtype_msg, flags=1),          # constant message
            _message(0x0005, fill),
        ]
        tail = list(attrs)
        # the layout message holds the data address: compute it from the header size when the data follows
        layout_len = len(_message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0)))
        header_len = 16 + sum(len(m) for m in layout_at_msgs) + layout_len + sum(len(m) for m in tail)
        pad = -len(self.buf) % 8
        header_addr = len(self.buf) + pad
        if raw is not None:
            data_addr = header_addr + header_len
            data_addr += -data_addr % 8
        layout = _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr if nbytes else UNDEF, nbytes))
        at = self.put(_object_header(layout_at_msgs + [layout] + tail))
        assert at == header_addr
        if raw is not None and nbytes:
            got = self.put(raw)
            assert got == data_addr
        return at

    # ---- groups -------------------------------------------------------------------------------------------
    def group(self, entries, leaf_k, attrs=()):
        """Old-style group (symbol table message): one local heap, one B-tree node, one symbol table node.
        entries: list of (name, object header address, (btree, heap) or None).  Returns (header, btree, heap)."""
        entries = sorted(entries, key=lambda e: e[0].encode("utf-8"))
        assert len(entries) <= 2 * leaf_k
        # local heap data segment: the empty string at offset 0, then the names
        seg = bytearray(b"\x00" * 8)
        offs = []
        for name, _, _ in entries:
            offs.append(len(seg))
            seg += _pad8(name.encode("utf-8") + b"\x00")
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 16)                     # one free block closing the segment (next = 1: none)
        snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(entries)))
        for (name, addr, cache), off in zip(entries, offs):
            if cache is None:
                snod += struct.pack("<QQII16x", off, addr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", off, addr, 1, 0, cache[0], cache[1])
        snod += b"\x00" * (40 * (2 * leaf_k - len(entries)))
        snod_at = self.put(bytes(snod))
        heap_hdr_at = self.put(b"\x00" * 32)
        seg_at = self.put(bytes(seg))
        self.patch(heap_hdr_at, b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_at))
        k = _GROUP_INTERNAL_K
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF))
        tree += struct.pack("<QQQ", 0, snod_at, offs[-1] if offs else 0)
        tree += b"\x00" * (24 + (2 * k + 1) * 8 + 2 * k * 8 - len(tree))
        tree_at = self.put(bytes(tree))
        hdr_at = self.put(_object_header([_message(0x0011, struct.pack("<QQ", tree_at, heap_hdr_at))] + list(attrs)))
        return hdr_at, tree_at, heap_hdr_at


def _float_bytes(a):
    return np.ascontiguousarray(a, dtype="<f8").tobytes()


def build_metadata(shape, x_itemsize, obs_index, obs_columns, var_index, obs_index_name="_index",
                   var_index_name="_index"):
    """The whole file except the matrix.  Returns (bytes, address of the matrix data, end-of-file address)."""
    n_i, n_j = int(shape[0]), int(shape[1])
    obs_index = [str(s) for s in obs_index]
    var_index = [str(s) for s in var_index]
    if len(obs_index) != n_i or len(var_index) != n_j:
        raise ValueError("index lengths do not match the matrix shape")
    f = _File()
    f.put(b"\x00" * 96)                                       # superblock, written last
    leaf_k = max(4, -(-(len(obs_columns) + 1) // 2))

    def enc(kind, version):
        return [f.str_attr("encoding-type", kind), f.str_attr("encoding-version", version)]

    def frame(index_name, index, columns):
        entries = [(index_name, f.dataset(_dt_vlen_string(), (len(index),), raw=f.vlen_strings(index),
                                          attrs=enc("string-array", "0.2.0")), None)]
        for name, values in columns:
            values = np.asarray(values, dtype=np.float64)
            if values.shape != (len(index),):
                raise ValueError("column %r has the wrong length" % (name,))
            entries.append((name, f.dataset(_dt_float(8), values.shape, raw=_float_bytes(values),
                                            attrs=enc("array", "0.2.0")), None))
        attrs = enc("dataframe", "0.2.0") + [f.str_attr("_index", index_name),
                                             f.str_array_attr("column-order", [c[0] for c in columns])]
        return f.group(entries, leaf_k, attrs)

    obs = frame(obs_index_name, obs_index, list(obs_columns))
    var = frame(var_index_name, var_index, [])
    x_nbytes = n_i * n_j * x_itemsize
    # the matrix header needs the address of data that lies behind ALL metadata: reserve, finish, patch
    x_attrs = enc("array", "0.2.0")
    x_hdr = f.dataset(_dt_float(x_itemsize), (n_i, n_j), data_addr=0, nbytes=x_nbytes, attrs=x_attrs)
    root = f.group([("X", x_hdr, None), ("obs", obs[0], obs[1:]), ("var", var[0], var[1:])], leaf_k,
                   enc("anndata", "0.1.0"))
    x_addr = f.tell() + (-f.tell() % _X_ALIGN) if x_nbytes else UNDEF
    # patch the layout message of X (it is the 4th message: find the placeholder by its unique encoding)
    placeholder = struct.pack("<BBQQ", 3, 1, 0 if x_nbytes else UNDEF, x_nbytes)
    at = f.buf.find(placeholder, x_hdr)
    assert 0 < at < x_hdr + 256
    f.patch(at, struct.pack("<BBQQ", 3, 1, x_addr, x_nbytes))
    eof = (x_addr + x_nbytes) if x_nbytes else f.tell()
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", leaf_k, _GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQIIQQ", 0, root[0], 1, 0, root[1], root[2])
    assert len(sb) == 96
    f.patch(0, sb)
    meta = bytes(f.buf)
    if x_nbytes:
        meta += b"\x00" * (x_addr - len(meta))
    return meta, x_addr, eof


def write_h5ad(path, X, obs_index, obs_columns, var_index, obs_index_name="_index", var_index_name="_index",
               chunk_bytes=64 << 20):
    """X [I, J] float32/float64 (C-contiguous is streamed without a copy), obs_columns: [(name, float64 [I])]."""
    X = np.asarray(X)
    if X.ndim != 2 or X.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("X must be a 2-D float32 or float64 array")
    if not X.flags.c_contiguous:
        X = np.ascontiguousarray(X)
    meta, x_addr, eof = build_metadata(X.shape, X.dtype.itemsize, obs_index, obs_columns, var_index, obs_index_name,
                                       var_index_name)
    tmp = str(path) + ".part"
    with open(tmp, "wb", buffering=0) as fh:
        fh.write(meta)
        flat = memoryview(X).cast("B")
        for at in range(0, len(flat), chunk_bytes):
            fh.write(flat[at:at + chunk_bytes])
    os.replace(tmp, str(path))
    return eof


def write_anndata(path, adata, **kw):
    """`AnnData.write(path)` for the result of OTModel.compute_transport_map (ot_model.py:326): obs index + float
    columns, var index, dense X."""
    obs = adata.obs
    cols = [(str(c), np.asarray(obs[c], dtype=np.float64)) for c in obs.columns]
    return write_h5ad(path, np.asarray(adata.X), list(obs.index.astype(str)), cols, list(adata.var.index.astype(str)),
                      **kw)


class AsyncWriter:
    """Jobs (file writes) on `threads` background threads, at most `depth` waiting, so that the disk overlaps the next
    solve.  One thread is the default and the fastest: measured on the B200 box (profiles/r2n_api_breakdown.txt) six
    1.2 GB maps go to the page cache at 4.7 GB/s from one thread, 3.0 GB/s from three, 2.4 GB/s from six.
    Errors surface at the next submit() or at close()."""

    def __init__(self, depth=2, threads=1):
        self._jobs = queue.Queue(maxsize=max(1, depth))
        self._err = None
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(1, threads))]
        for t in self._threads:
            t.start()

    def _run(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            try:
                if self._err is None:
                    job()
            except BaseException as exc:  # noqa: BLE001 - re-raised on the submitting thread
                self._err = exc

    def _check(self):
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def submit(self, fn):
        self._check()
        self._jobs.put(fn)

    def close(self):
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._check()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---------------------------------------------------------------------------------------------------------------
# reader (independent of the writer's code paths: it parses what is on disk)
# ---------------------------------------------------------------------------------------------------------------
class H5Object:
    def __init__(self, reader, addr):
        self._r, self.addr = reader, addr
        self.messages = reader._messages(addr)
        self.attrs = {}
        for mtype, data in self.messages:
            if mtype == 0x000C:
                name, value = reader._parse_attribute(data)
                self.attrs[name] = value

    @property
    def is_group(self):
        return any(t == 0x0011 for t, _ in self.messages)

    def keys(self):
        return list(self._links())

    def _links(self):
        for t, data in self.messages:
            if t == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data)
                return self._r._group_entries(btree, heap)
        raise TypeError("not a group")

    def __contains__(self, name):
        return name in self._links()

    def __getitem__(self, name):
        node = self
        for part in [p for p in str(name).split("/") if p]:
            node = H5Object(node._r, node._links()[part])
        return node

    # ---- dataset ------------------------------------------------------------------------------------------
    def _describe(self):
        shape = dtype = layout = None
        for t, data in self.messages:
            if t == 0x0001:
                shape = self._r._parse_dataspace(data)
            elif t == 0x0003:
                dtype = self._r._parse_datatype(data)
            elif t == 0x0008:
                layout = data
        if shape is None or dtype is None or layout is None:
            raise TypeError("not a dataset")
        return shape, dtype, layout

    @property
    def shape(self):
        return self._describe()[0]

    def read(self):
        shape, dtype, layout = self._describe()
        version, cls = layout[0], layout[1]
        count = int(np.prod(shape)) if shape else 1
        if version in (1, 2):                           # HDF5 1.6 files: dimensionality, class, address, sizes
            rank, cls = layout[1], layout[2]
            if cls != 1:
                raise NotImplementedError("version-%d layout of class %d" % (version, cls))
            addr, = struct.unpack_from("<Q", layout, 8)
            size = count * dtype[1]
            raw = self._r._read(addr, size) if size and addr != UNDEF else b""
        elif version != 3:
            raise NotImplementedError("data layout message version %d" % version)
        elif cls == 1:                                  # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            raw = self._r._read(addr, size) if size and addr != UNDEF else b""
        elif cls == 0:                                  # compact
            size, = struct.unpack_from("<H", layout, 2)
            raw = layout[4:4 + size]
        else:
            raise NotImplementedError("chunked datasets are not supported by this reader")
        return self._r._decode(dtype, raw, shape, count)


class H5Reader:
    """Minimal HDF5 reader: superblock 0/1, version-1 object headers (with continuation blocks), symbol-table groups,
    contiguous and compact datasets of fixed-point / floating-point / fixed and variable-length string types,
    version 1-3 attribute messages."""

    def __init__(self, path):
        self.fh = open(path, "rb")
        self.size = os.fstat(self.fh.fileno()).st_size
        at = 0
        while True:                                      # a user block pushes the superblock to 512, 1024, ...
            self.fh.seek(at)
            if self.fh.read(8) == SIGNATURE:
                break
            at = 512 if at == 0 else at * 2
            if at >= self.size:
                raise ValueError("not an HDF5 file")
        sb = self._raw(at, 128)
        version = sb[8]
        if version not in (0, 1):
            raise NotImplementedError("superblock version %d" % version)
        if sb[13] != 8 or sb[14] != 8:
            raise NotImplementedError("only 8-byte offsets and lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", sb, 16)
        p = 24 if version == 0 else 28
        # every address in the file is relative to the base address (the superblock's offset when a user block exists)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", sb, p)
        _, root_hdr, cache, _, tree, heap = struct.unpack_from("<QQIIQQ", sb, p + 32)
        self.root = H5Object(self, root_hdr)

    def close(self):
        self.fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __getitem__(self, name):
        return self.root[name]

    def _raw(self, at, n):
        self.fh.seek(at)
        return self.fh.read(n)

    def _read(self, addr, n):
        return self._raw(self.base + addr, n)

    # ---- object headers -------------------------------------------------------------------------------------
    def _messages(self, addr):
        head = self._read(addr, 16)
        version, n_msgs, _, size = struct.unpack_from("<BxHII", head)
        if version != 1:
            raise NotImplementedError("object header version %d (only the classic version 1 is read)" % version)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            at, size = blocks.pop(0)
            data = self._read(at, size)
            p = 0
            while p + 8 <= len(data) and len(out) < n_msgs:
                mtype, msize, _ = struct.unpack_from("<HHB", data, p)
                body = data[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                      # continuation
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, body))
        return out

    # ---- groups ------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, offset):
        hdr = self._read(heap_addr, 32)
        if hdr[:4] != b"HEAP":
            raise ValueError("bad local heap signature")
        seg_size, _, seg_addr = struct.unpack_from("<QQQ", hdr, 8)
        seg = self._read(seg_addr, seg_size)
        end = seg.index(b"\x00", offset)
        return seg[offset:end].decode("utf-8")

    def _group_entries(self, btree, heap):
        out = {}

        def walk(addr):
            node = self._read(addr, 24 + (4 * self.internal_k + 1) * 8)
            if node[:4] != b"TREE":
                raise ValueError("bad B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", node, 4)
            if ntype != 0:
                raise ValueError("not a group B-tree")
            for k in range(used):
                child, = struct.unpack_from("<Q", node, 24 + 8 + 16 * k)
                if level > 0:
                    walk(child)
                    continue
                snod = self._read(child, 8 + 40 * 2 * self.leaf_k)
                if snod[:4] != b"SNOD":
                    raise ValueError("bad symbol table node signature")
                n, = struct.unpack_from("<H", snod, 6)
                for e in range(n):
                    name_off, hdr = struct.unpack_from("<QQ", snod, 8 + 40 * e)
                    out[self._heap_name(heap, name_off)] = hdr

        walk(btree)
        return out

    # ---- messages ------------------------------------------------------------------------------------------------
    @staticmethod
    def _parse_dataspace(data):
        version, rank, flags = data[0], data[1], data[2]
        p = 8 if version == 1 else 4
        return tuple(struct.unpack_from("<Q", data, p + 8 * k)[0] for k in range(rank))

    @staticmethod
    def _parse_datatype(data):
        cls, version = data[0] & 0x0F, data[0] >> 4
        bits = data[1] | (data[2] << 8) | (data[3] << 16)
        size, = struct.unpack_from("<I", data, 4)
        if cls == 0:
            return ("int", size, bool(bits & 0x08), "<" if not bits & 1 else ">")
        if cls == 1:
            return ("float", size, "<" if not bits & 1 else ">")
        if cls == 3:
            return ("string", size, bits & 0x0F)
        if cls == 9:
            if bits & 0x0F != 1:
                raise NotImplementedError("variable-length sequences")
            return ("vlen_string",)
        raise NotImplementedError("datatype class %d" % cls)

    def _global_heap_object(self, addr, index):
        hdr = self._read(addr, 16)
        if hdr[:4] != b"GCOL":
            raise ValueError("bad global heap signature")
        total, = struct.unpack_from("<Q", hdr, 8)
        data = self._read(addr, total)
        p = 16
        while p + 16 <= total:
            idx, _, size = struct.unpack_from("<HH4xQ", data, p)
            if idx == 0:
                break
            if idx == index:
                return data[p + 16:p + 16 + size]
            p += 16 + size + (-size % 8)
        raise KeyError("global heap object %d not found" % index)

    def _decode(self, dtype, raw, shape, count):
        kind = dtype[0]
        if kind == "float":
            return np.frombuffer(raw, dtype="%sf%d" % (dtype[2], dtype[1]), count=count).reshape(shape).copy()
        if kind == "int":
            code = ("i" if dtype[2] else "u") + str(dtype[1])
            return np.frombuffer(raw, dtype=dtype[3] + code, count=count).reshape(shape).copy()
        if kind == "string":
            n = dtype[1]
            vals = [raw[k * n:(k + 1) * n].split(b"\x00")[0].decode("utf-8") for k in range(count)]
        else:
            vals = []
            cache = {}
            for k in range(count):
                length, addr, idx = struct.unpack_from("<IQI", raw, 16 * k)
                if length == 0 and addr == 0:
                    vals.append("")
                    continue
                if addr not in cache:
                    cache[addr] = self._collection(addr)
                vals.append(cache[addr][idx][:length].decode("utf-8"))
        if not shape:
            return vals[0]
        return np.array(vals, dtype=object).reshape(shape)

    def _collection(self, addr):
        hdr = self._read(addr, 16)
        if hdr[:4] != b"GCOL":
            raise ValueError("bad global heap signature")
        total, = struct.unpack_from("<Q", hdr, 8)
        data = self._read(addr, total)
        objs, p = {}, 16
        while p + 16 <= total:
            idx, _, size = struct.unpack_from("<HH4xQ", data, p)
            if idx == 0:
                break
            objs[idx] = data[p + 16:p + 16 + size]
            p += 16 + size + (-size % 8)
        return objs

    def _parse_attribute(self, data):
        version = data[0]
        if version == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8
            name = data[p:p + nsz].split(b"\x00")[0].decode("utf-8")
            p += nsz + (-nsz % 8)
            dt = data[p:p + dsz]
            p += dsz + (-dsz % 8)
            sp = data[p:p + ssz]
            p += ssz + (-ssz % 8)
        elif version in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8 if version == 2 else 9
            name = data[p:p + nsz].split(b"\x00")[0].decode("utf-8")
            p += nsz
            dt = data[p:p + dsz]
            p += dsz
            sp = data[p:p + ssz]
            p += ssz
        else:
            raise NotImplementedError("attribute message version %d" % version)
        shape = self._parse_dataspace(sp)
        dtype = self._parse_datatype(dt)
        count = int(np.prod(shape)) if shape else 1
        return name, self._decode(dtype, data[p:], shape, count)


def read_h5ad(path, with_x=True):
    """dict(X, obs_index, obs (name -> column), var_index) of a transport-map .h5ad, read the way
    transport_map_model.py:709-721 does: the `_index` attribute of /obs and /var names the id dataset."""
    with H5Reader(path) as f:
        obs, var = f["obs"], f["var"]
        obs_key = obs.attrs.get("_index", "index")
        var_key = var.attrs.get("_index", "index")
        out = {"obs_index": obs[obs_key].read().astype(str), "var_index": var[var_key].read().astype(str)}
        order = obs.attrs.get("column-order")
        names = [str(c) for c in (order if order is not None and np.ndim(order) else [])]
        if not names:
            names = [k for k in obs.keys() if k != obs_key]
        out["obs"] = {name: obs[name].read() for name in names}
        if with_x:
            out["X"] = f["X"].read()
        out["attrs"] = dict(f.root.attrs)
    return out
