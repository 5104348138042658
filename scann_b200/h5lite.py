"""Minimal, dependency-free HDF5 reader / writer for Keras-2.10 weight files (h5py / libhdf5 are not installed).

The reference saves and restores models with ``model.save("...h5")`` / ``ModelCheckpoint`` / ``load_model``
(scann/models/scann_model.py:79-96,223-230), i.e. Keras' legacy HDF5 layout:

    /                       attrs: keras_version, backend, (model_config, training_config)
    /model_weights          attrs: layer_names (fixed-length byte strings)      [or the root itself for save_weights]
    /model_weights/<layer>  attrs: weight_names
    /model_weights/<layer>/<weight path>   float32 dataset, e.g. "local_attention/query/kernel:0"

Subset of the HDF5 file format implemented (what h5py writes with its default ``libver='earliest'``):
superblock version 0, version-1 object headers (with continuation blocks), symbol-table groups (version-1 B-tree +
local heap + SNOD nodes), contiguous and compact dataset layouts, fixed-point / IEEE-float / fixed-length-string
/ variable-length-string datatypes (variable-length data through global heap collections), attribute messages
version 1-3.  Not implemented (raises): chunked or filtered datasets, version-2 object headers / link messages
("latest" format), compound types.  The writer emits the same classic structures, so files it writes are
readable by libhdf5 / h5py.

Pinned against a file written by libhdf5 itself: scipy ships a MATLAB v7.3 (= HDF5 with a 512-byte user block)
test file, parsed in tests/test_h5lite.py.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# =============================================================================================== reader
class _Dataset:
    def __init__(self, name, shape, dtype, data, attrs):
        self.name, self.shape, self.dtype, self._data, self.attrs = name, shape, dtype, data, attrs

    def __getitem__(self, key):
        return self._data[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)


class _Group:
    def __init__(self, name, attrs, children):
        self.name, self.attrs, self._children = name, attrs, children

    def keys(self):
        return list(self._children.keys())

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, _Group) or part not in node._children:
                raise KeyError(path)
            node = node._children[part]
        return node

    def items(self):
        return self._children.items()

    def visit_datasets(self, prefix=""):
        for k, v in self._children.items():
            p = f"{prefix}/{k}" if prefix else k
            if isinstance(v, _Group):
                yield from v.visit_datasets(p)
            else:
                yield p, v


class File(_Group):
    """Read-only view of an HDF5 file: ``File(path)["model_weights"]["dense_embed"].attrs["weight_names"]``."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self._b = f.read()
        b = self._b
        base = -1
        off = 0
        while off < len(b):                      # the superblock may follow a user block of 512, 1024, ... bytes
            if b[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise H5Error("not an HDF5 file (no signature)")
        ver = b[base + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} is not supported (only the classic format, versions 0/1)")
        self._O, self._L = b[base + 13], b[base + 14]
        if self._O != 8 or self._L != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = base + 24 + (4 if ver == 1 else 0)
        self._base = struct.unpack_from("<Q", b, p)[0]          # base address: all file addresses are relative to it
        if self._base == 0 and base != 0:
            self._base = base
        p += 32                                                   # base, free-space, end-of-file, driver info
        _, ohdr, cache, _ = struct.unpack_from("<QQII", b, p)     # root group symbol table entry
        root = self._read_object("/", ohdr)
        super().__init__("/", root.attrs, root._children)

    # ---- low level ------------------------------------------------------------------------------
    def _at(self, addr: int) -> int:
        return self._base + addr

    def _messages(self, addr: int) -> List[Tuple[int, int, bytes]]:
        b = self._b
        p = self._at(addr)
        ver = b[p]
        if ver != 1:
            raise H5Error(f"object header version {ver} is not supported (file written with libver='latest'?)")
        nmsg, _refs, hsize = struct.unpack_from("<HII", b, p + 2)
        blocks = [(p + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, q)
                data = b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x10:                                  # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._at(caddr), clen))
                out.append((mtype, flags, data))
        return out

    def _parse_datatype(self, d: bytes, p: int = 0):
        """-> (kind, size, np dtype or info, bytes consumed)"""
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", d, p)
        cls, q = cv & 0x0F, p + 8
        if cls == 0:                                               # fixed point
            order = ">" if b0 & 1 else "<"
            signed = bool(b0 & 8)
            return "num", size, np.dtype(f"{order}{'i' if signed else 'u'}{size}"), q + 4 - p
        if cls == 1:                                               # IEEE float
            order = ">" if b0 & 1 else "<"
            return "num", size, np.dtype(f"{order}f{size}"), q + 12 - p
        if cls == 3:                                               # fixed-length string
            return "str", size, np.dtype(f"S{size}"), q - p
        if cls == 9:                                               # variable length
            is_str = (b0 & 0x0F) == 1
            bk, bsize, bdt, used = self._parse_datatype(d, q)
            return ("vstr" if is_str else "vlen"), size, (bk, bsize, bdt), q + used - p
        if cls == 7:                                               # reference
            return "num", size, np.dtype(f"V{size}"), q - p
        raise H5Error(f"datatype class {cls} is not supported")

    @staticmethod
    def _parse_dataspace(d: bytes) -> Tuple[int, ...]:
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:                                          # null dataspace
                return (0,)
            p = 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", d, p + 8 * i)[0] for i in range(rank))

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        b = self._b
        p = self._at(addr)
        if b[p:p + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        csize = struct.unpack_from("<Q", b, p + 8)[0]
        q, end = p + 16, p + csize
        while q + 16 <= end:
            idx, _ref, _res, size = struct.unpack_from("<HHIQ", b, q)
            if idx == 0:
                break
            if idx == index:
                return b[q + 16:q + 16 + size]
            q += 16 + (size + 7) // 8 * 8
        raise H5Error("global heap object not found")

    def _decode(self, kind, size, info, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if kind in ("num", "str"):
            arr = np.frombuffer(raw[:n * size], dtype=info).reshape(shape)
            if kind == "num" and arr.dtype.byteorder == ">":
                arr = arr.astype(arr.dtype.newbyteorder("<"))
            return arr if shape else arr.reshape(())[()]
        if kind == "vstr":
            vals = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(gaddr, gidx)[:ln].decode("utf-8", "replace") if ln else "")
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        raise H5Error(f"cannot decode {kind}")

    def _parse_attribute(self, d: bytes):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        name = d[p:p + nsz].split(b"\x00")[0].decode()
        p += pad(nsz)
        kind, size, info, _ = self._parse_datatype(d, p)
        p += pad(tsz)
        shape = self._parse_dataspace(d[p:p + ssz]) if ssz else ()
        p += pad(ssz)
        return name, self._decode(kind, size, info, shape, d[p:])

    def _group_children(self, btree: int, heap: int) -> Dict[str, int]:
        b = self._b
        hp = self._at(heap)
        if b[hp:hp + 4] != b"HEAP":
            raise H5Error("bad local heap")
        hdata = self._at(struct.unpack_from("<Q", b, hp + 24)[0])
        out: Dict[str, int] = {}

        def walk(addr):
            p = self._at(addr)
            sig = b[p:p + 4]
            if sig == b"TREE":
                _type, _level, used = struct.unpack_from("<BBH", b, p + 4)
                q = p + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", b, q + 8 + 16 * i)[0]
                    walk(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", b, p + 6)[0]
                for i in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, p + 8 + 40 * i)
                    s = hdata + noff
                    out[b[s:b.index(b"\x00", s)].decode()] = ohdr
            else:
                raise H5Error("bad group B-tree node")

        walk(btree)
        return out

    def _read_object(self, name: str, addr: int):
        msgs = self._messages(addr)
        attrs, shape, dt, layout, stab = {}, None, None, None, None
        for mtype, _flags, d in msgs:
            if mtype == 0x0C:
                k, v = self._parse_attribute(d)
                attrs[k] = v
            elif mtype == 0x01:
                shape = self._parse_dataspace(d)
            elif mtype == 0x03:
                dt = self._parse_datatype(d)
            elif mtype == 0x08:
                layout = d
            elif mtype == 0x11:
                stab = struct.unpack_from("<QQ", d, 0)
            elif mtype in (0x02, 0x06):
                raise H5Error("link-info / link messages (libver='latest' groups) are not supported")
        if stab is not None:
            kids = {k: self._read_object(f"{name.rstrip('/')}/{k}", a) for k, a in self._group_children(*stab).items()}
            return _Group(name, attrs, kids)
        if dt is None or shape is None or layout is None:
            raise H5Error(f"{name}: not a group and not a dataset")
        kind, size, info, _ = dt
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 1:
                daddr, dsize = struct.unpack_from("<QQ", layout, 2)
                raw = b"" if daddr == UNDEF else self._b[self._at(daddr):self._at(daddr) + dsize]
                if daddr == UNDEF:
                    raw = bytes(n * size)
            elif cls == 0:
                csz = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + csz]
            else:
                raise H5Error(f"{name}: chunked / filtered datasets are not supported")
        elif ver in (1, 2):
            rank, cls = layout[1], layout[2]
            if cls == 1:
                daddr = struct.unpack_from("<Q", layout, 8)[0]
                raw = self._b[self._at(daddr):self._at(daddr) + n * size]
            elif cls == 0:
                p = 8 + 4 * rank
                csz = struct.unpack_from("<I", layout, p)[0]
                raw = layout[p + 4:p + 4 + csz]
            else:
                raise H5Error(f"{name}: chunked datasets are not supported")
        else:
            raise H5Error(f"{name}: layout version {ver}")
        data = self._decode(kind, size, info, shape, raw)
        return _Dataset(name, shape, getattr(data, "dtype", None), data, attrs)


# =============================================================================================== writer
def _pad8(x: bytes) -> bytes:
    return x + bytes(-len(x) % 8)


def _dt_bytes(dtype: np.dtype) -> bytes:
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        size = dtype.itemsize
        es, ms, bias = {4: (8, 23, 127), 8: (11, 52, 1023), 2: (5, 10, 15)}[size]
        return (struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) +
                struct.pack("<HHBBBBI", 0, size * 8, ms, es, 0, ms, bias))
    if dtype.kind in "iu":
        size = dtype.itemsize
        return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, max(dtype.itemsize, 1))
    raise H5Error(f"dtype {dtype} is not supported by the writer")


def _ds_bytes(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_message(name: str, value) -> bytes:
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        arr, shape = np.array(value, dtype=f"S{max(len(value), 1)}"), ()
    else:
        arr = np.ascontiguousarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind == "O":
            arr = np.array([x if isinstance(x, bytes) else str(x).encode() for x in arr.ravel()]).reshape(arr.shape)
        shape = arr.shape
    nm = name.encode() + b"\x00"
    dt, ds = _dt_bytes(arr.dtype), _ds_bytes(shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    if len(body) > 64000:
        raise H5Error(f"attribute {name!r} exceeds the 64 KB object-header message limit")
    return _message(0x0C, body)


class _WNode:
    def __init__(self):
        self.attrs: Dict[str, object] = {}
        self.children: Dict[str, "_WNode"] = {}
        self.data: Optional[np.ndarray] = None


class Writer:
    """``w = Writer(); w.attr("/", "backend", "tensorflow"); w.dataset("/g/x", arr); w.save(path)``.
    Groups are created implicitly.  Classic format: superblock 0, symbol-table groups, contiguous datasets."""

    LEAF_K, INT_K = 64, 16          # a group holds at most 2 * LEAF_K = 128 links (one symbol-table node)

    def __init__(self):
        self.root = _WNode()

    def _node(self, path: str, create: bool = True) -> _WNode:
        node = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[part] = _WNode()
            node = node.children[part]
        return node

    def group(self, path: str) -> None:
        self._node(path)

    def attr(self, path: str, name: str, value) -> None:
        self._node(path).attrs[name] = value

    def dataset(self, path: str, array) -> None:
        arr = np.ascontiguousarray(array)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        self._node(path).data = arr

    # ---- serialisation --------------------------------------------------------------------------
    def save(self, path: str) -> None:
        O = 8
        snod_size = 8 + 2 * self.LEAF_K * 40
        tree_size = 24 + (2 * self.INT_K + 1) * 8 + 2 * self.INT_K * O
        chunks: List[Tuple[int, bytes]] = []
        cursor = [96]                               # superblock 0 with 8-byte addresses: 56 + 40 bytes

        def alloc(n: int) -> int:
            a = cursor[0]
            cursor[0] += (n + 7) // 8 * 8
            return a

        def emit(node: _WNode) -> int:
            """writes the object (children first) and returns its object-header address"""
            msgs = [_attr_message(k, v) for k, v in node.attrs.items()]
            if node.data is not None:
                arr = node.data
                daddr = alloc(max(arr.nbytes, 1))
                chunks.append((daddr, arr.tobytes()))
                head = [_message(0x01, _ds_bytes(arr.shape)), _message(0x03, _dt_bytes(arr.dtype), flags=1),
                        _message(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),
                        _message(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]
                msgs = head + msgs
            else:
                names = sorted(node.children)                 # symbol-table entries are sorted by name
                if len(names) > 2 * self.LEAF_K:
                    raise H5Error("too many links in one group for this writer")
                child_addr = {n: emit(node.children[n]) for n in names}
                heap = bytearray(8)                           # offset 0: the empty string
                noff = {}
                for n in names:
                    noff[n] = len(heap)
                    heap += _pad8(n.encode() + b"\x00")
                free_off = len(heap)
                heap += struct.pack("<QQ", 1, 16)             # one free block (next = 1 means last, size 16)
                hdata = alloc(len(heap))
                chunks.append((hdata, bytes(heap)))
                haddr = alloc(32)
                chunks.append((haddr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, hdata)))
                snod = alloc(snod_size)
                ents = b"".join(struct.pack("<QQII16x", noff[n], child_addr[n], 0, 0) for n in names)
                chunks.append((snod, (b"SNOD" + struct.pack("<BBH", 1, 0, len(names)) + ents).ljust(snod_size, b"\x00")))
                tree = alloc(tree_size)
                if names:
                    body = struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, noff[names[-1]])
                else:
                    body = struct.pack("<BBHQQ", 0, 0, 0, UNDEF, UNDEF)
                chunks.append((tree, (b"TREE" + body).ljust(tree_size, b"\x00")))
                node._stab = (tree, haddr)
                msgs = [_message(0x11, struct.pack("<QQ", tree, haddr))] + msgs
            body = b"".join(msgs)
            ohdr = alloc(16 + len(body))
            chunks.append((ohdr, struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body))
            return ohdr

        root_ohdr = emit(self.root)
        eof = cursor[0]
        sb = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", self.LEAF_K, self.INT_K, 0) +
              struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) +
              struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", *self.root._stab))
        assert len(sb) == 96
        out = bytearray(eof)
        out[:96] = sb
        for a, data in chunks:
            out[a:a + len(data)] = data
        with open(path, "wb") as f:
            f.write(bytes(out))


# =============================================================================================== Keras layout
def _as_names(v) -> List[str]:
    if isinstance(v, (bytes, str)):
        v = [v]
    return [x.decode() if isinstance(x, bytes) else str(x) for x in np.asarray(v).ravel().tolist()]


def load_keras_weights(path: str) -> List[Tuple[str, List[Tuple[str, np.ndarray]]]]:
    """[(layer name, [(weight name, array), ...]), ...] in the file's ``layer_names`` / ``weight_names`` order --
    the order ``keras.Model.load_weights`` consumes (hdf5_format.load_weights_from_hdf5_group).  Accepts both the
    full-model layout (``/model_weights``) and the ``save_weights`` layout (root)."""
    f = File(path)
    g = f["model_weights"] if ("layer_names" not in f.attrs and "model_weights" in f) else f
    if "layer_names" not in g.attrs:
        raise H5Error("not a Keras weight file: no layer_names attribute")
    out = []
    for lname in _as_names(g.attrs["layer_names"]):
        lg = g[lname]
        ws = []
        for wname in _as_names(lg.attrs.get("weight_names", [])):
            ws.append((wname, np.asarray(lg[wname])))
        out.append((lname, ws))
    return out


def save_keras_weights(path: str, layers: List[Tuple[str, List[Tuple[str, np.ndarray]]]], full_model: bool = True,
                       model_config: Optional[str] = None, keras_version: str = "2.10.0") -> None:
    """Writes the Keras-2.10 legacy HDF5 layout (full-model ``/model_weights`` or ``save_weights`` style)."""
    w = Writer()
    root = "/model_weights" if full_model else "/"
    for k, v in (("keras_version", keras_version), ("backend", "tensorflow")):
        w.attr("/", k, v)
        if full_model:
            w.attr(root, k, v)
    if full_model and model_config is not None:
        w.attr("/", "model_config", model_config)
    w.group(root)
    w.attr(root, "layer_names", np.array([n.encode() for n, _ in layers]))
    for lname, ws in layers:
        lp = f"{root.rstrip('/')}/{lname}"
        w.group(lp)
        w.attr(lp, "weight_names", np.array([n.encode() for n, _ in ws]) if ws else np.zeros((0,), "S1"))
        for wname, arr in ws:
            w.dataset(f"{lp}/{wname}", np.asarray(arr, np.float32))
    w.save(path)
