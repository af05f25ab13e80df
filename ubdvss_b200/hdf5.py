"""Minimal HDF5 reader / writer for Keras weight files (net.py:418-427, 443-494) - no h5py.

The reference saves ``model.h5`` / ``inference_model.h5`` / ``model_weights.h5`` with Keras 2.x on top of
h5py; neither is installed here, so this module reads and writes the subset of the HDF5 file format
(HDF5 File Format Specification, version 1.1 / 2.0 objects) those files use:

* superblock version 0/1 (what libhdf5 writes with ``libver='earliest'``, h5py's default) and 2/3,
  optionally behind a user block (the signature is searched at 0, 512, 1024, ...);
* groups as symbol tables (version-1 B-tree + local heap + ``SNOD`` nodes) or as compact link messages;
* object headers version 1 and 2 with continuation blocks;
* datasets: contiguous, compact and chunked (version-1 chunk B-tree, optional deflate / shuffle filters),
  little- or big-endian integers and IEEE floats;
* attributes (message versions 1-3): numeric arrays, fixed-length strings, variable-length strings
  (global heap).

``write_keras_weights`` produces a superblock-0 file with symbol-table groups, contiguous float32 datasets
and fixed-length string attributes in the layout ``keras.engine.saving.save_weights_to_hdf5_group`` uses
(``layer_names`` / ``backend`` / ``keras_version`` on the root, ``weight_names`` on every layer group,
datasets at ``<layer>/<weight name>``).  There is no libhdf5 in this environment to cross-check the written
files; the reader is exercised on a file written by the real library (SciPy's MATLAB-7.3 test file)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(IOError):
    pass


# ======================================================================================= reader

class _Dtype:
    def __init__(self, cls, size, np_dtype=None, vlen_string=False, base=None):
        self.cls, self.size, self.np_dtype, self.vlen_string, self.base = cls, size, np_dtype, vlen_string, base


class Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters, attrs):
        self._f, self.name, self.shape, self._dt, self._layout, self._filters, self.attrs = f, name, shape, dtype, layout, filters, attrs

    @property
    def dtype(self):
        return self._dt.np_dtype

    def __getitem__(self, key):
        return self.read()[key]

    def read(self) -> np.ndarray:
        f, dt = self._f, self._dt
        if dt.np_dtype is None:
            raise HDF5Error(f"{self.name}: unsupported dataset datatype class {dt.cls}")
        n = int(np.prod(self.shape)) if self.shape else 1
        kind = self._layout[0]
        if kind == "compact":
            raw = self._layout[1]
        elif kind == "contiguous":
            addr, size = self._layout[1], self._layout[2]
            raw = b"\0" * (n * dt.size) if addr == UNDEF else f._read(addr, n * dt.size)
        elif kind == "chunked":
            return self._read_chunked()
        else:
            raise HDF5Error(f"{self.name}: unsupported layout {kind}")
        return np.frombuffer(raw[:n * dt.size], dtype=dt.np_dtype).reshape(self.shape).copy()

    def _read_chunked(self):
        f, dt = self._f, self._dt
        _, btree, chunk = self._layout
        rank = len(self.shape)
        out = np.zeros(self.shape, dtype=dt.np_dtype)
        if btree == UNDEF:
            return out
        for offs, size, mask, addr in f._chunk_btree(btree, rank):
            raw = f._read(addr, size)
            for i, (fid, _) in reversed(list(enumerate(self._filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                                  # shuffle
                    a = np.frombuffer(raw, np.uint8).reshape(dt.size, -1)
                    raw = a.T.tobytes()
                else:
                    raise HDF5Error(f"{self.name}: unsupported filter {fid}")
            block = np.frombuffer(raw, dtype=dt.np_dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = block[sl_in]
        return out


class Group:
    def __init__(self, f, name, links, attrs):
        self._f, self.name, self._links, self.attrs = f, name, links, attrs

    def keys(self):
        return list(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(Group):
    """``File(path)[name]`` / ``.attrs`` / ``.keys()`` in the spirit of ``h5py.File`` (read-only)."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self._buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self._buf = fh.read()
        self._cache = {}
        base = 0
        while True:
            if self._buf[base:base + 8] == SIGNATURE:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self._buf):
                raise HDF5Error("not an HDF5 file (signature not found)")
        b = self._buf
        ver = b[base + 8]
        if ver in (0, 1):
            self.O, self.L = b[base + 13], b[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.O)
            p += 4 * self.O                                     # base, free-space, end of file, driver info
            root_hdr = self._u(p + self.O, self.O)              # symbol table entry: name offset, header address
        elif ver in (2, 3):
            self.O, self.L = b[base + 9], b[base + 10]
            p = base + 12
            self.base = self._u(p, self.O)
            root_hdr = self._u(p + 3 * self.O, self.O)
        else:
            raise HDF5Error(f"unsupported superblock version {ver}")
        if self.base == 0 and base:
            self.base = base                                    # user block: addresses are relative to the superblock
        root = self._object(root_hdr, "/")
        super().__init__(self, "/", root._links, root.attrs)

    # ---- raw access
    def _read(self, addr, n):
        a = self.base + addr
        if a < 0 or a + n > len(self._buf):
            raise HDF5Error("address outside the file")
        return self._buf[a:a + n]

    def _u(self, pos, n):
        return int.from_bytes(self._buf[pos:pos + n], "little")

    # ---- object headers
    def _messages(self, addr):
        b, O, L = self._buf, self.O, self.L
        p = self.base + addr
        out = []
        if b[p:p + 4] == b"OHDR":                               # version 2
            flags = b[p + 5]
            q = p + 6
            if flags & 0x20:
                q += 16
            if flags & 0x10:
                q += 4
            szb = 1 << (flags & 3)
            size0 = self._u(q, szb)
            q += szb
            blocks = [(q, q + size0)]
            while blocks:
                q, end = blocks.pop(0)
                while q + 4 <= end:
                    mtype = b[q]
                    msize = self._u(q + 1, 2)
                    q += 4 + (2 if flags & 0x04 else 0)
                    body = b[q:q + msize]
                    if mtype == 0x10:
                        ca, cl = int.from_bytes(body[:O], "little"), int.from_bytes(body[O:O + L], "little")
                        blocks.append((self.base + ca + 4, self.base + ca + cl - 4))      # skip "OCHK", stop before checksum
                    elif mtype != 0:
                        out.append((mtype, body))
                    q += msize
            return out
        if b[p] != 1:
            raise HDF5Error(f"unsupported object header version {b[p]} at {addr}")
        nmsg = self._u(p + 2, 2)
        size0 = self._u(p + 8, 4)
        blocks = [(p + 16, p + 16 + size0)]
        while blocks and len(out) < nmsg + 64:
            q, end = blocks.pop(0)
            while q + 8 <= end:
                mtype = self._u(q, 2)
                msize = self._u(q + 2, 2)
                body = b[q + 8:q + 8 + msize]
                if mtype == 0x10:
                    ca, cl = int.from_bytes(body[:O], "little"), int.from_bytes(body[O:O + L], "little")
                    blocks.append((self.base + ca, self.base + ca + cl))
                elif mtype != 0:
                    out.append((mtype, body))
                q += 8 + msize
        return out

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs, links = {}, {}
        shape = dtype = layout = None
        filters = []
        is_group = False
        for mtype, body in msgs:
            if mtype == 0x0C:
                k, v = self._attribute(body)
                attrs[k] = v
            elif mtype == 0x11:                                 # symbol table: B-tree + local heap
                is_group = True
                O = self.O
                self._symtab(int.from_bytes(body[:O], "little"), int.from_bytes(body[O:2 * O], "little"), links)
            elif mtype == 0x06:                                 # link message (compact new-style group)
                is_group = True
                k, a = self._link(body)
                if a is not None:
                    links[k] = a
            elif mtype == 0x02:
                is_group = True
            elif mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)[0]
            elif mtype == 0x08:
                layout = self._layout(body)
            elif mtype == 0x0B:
                filters = self._filters(body)
        if layout is not None and dtype is not None and not is_group:
            obj = Dataset(self, name, shape if shape is not None else (), dtype, layout, filters, attrs)
        else:
            obj = Group(self, name, links, attrs)
        self._cache[addr] = obj
        return obj

    # ---- groups
    def _heap_string(self, heap_addr, off):
        p = self.base + heap_addr
        if self._buf[p:p + 4] != b"HEAP":
            raise HDF5Error("bad local heap")
        data = self._u(p + 8 + 2 * self.L, self.O)
        s = self.base + data + off
        e = self._buf.index(b"\0", s)
        return self._buf[s:e].decode("utf8")

    def _symtab(self, btree, heap, links):
        O, L, b = self.O, self.L, self._buf
        p = self.base + btree
        if b[p:p + 4] != b"TREE":
            raise HDF5Error("bad group B-tree node")
        level, used = b[p + 5], self._u(p + 6, 2)
        q = p + 8 + 2 * O
        for i in range(used):
            child = self._u(q + L + i * (L + O), O)
            if level > 0:
                self._symtab(child, heap, links)
                continue
            s = self.base + child
            if b[s:s + 4] != b"SNOD":
                raise HDF5Error("bad symbol table node")
            n = self._u(s + 6, 2)
            e = s + 8
            for _ in range(n):
                name_off, hdr = self._u(e, O), self._u(e + O, O)
                links[self._heap_string(heap, name_off)] = hdr
                e += 2 * O + 24

    def _link(self, body):
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nb = 1 << (flags & 3)
        ln = int.from_bytes(body[p:p + nb], "little"); p += nb
        name = bytes(body[p:p + ln]).decode("utf8"); p += ln
        if ltype != 0:
            return name, None                                   # soft / external links are not followed
        return name, int.from_bytes(body[p:p + self.O], "little")

    # ---- dataset pieces
    def _dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(body[p + i * self.L:p + (i + 1) * self.L], "little") for i in range(rank))

    def _datatype(self, body):
        cls = body[0] & 0x0F
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size = int.from_bytes(body[4:8], "little")
        order = ">" if bits & 1 else "<"
        if cls == 0:                                            # fixed point
            signed = bool(bits & 0x08)
            return _Dtype(cls, size, np.dtype(f"{order}{'i' if signed else 'u'}{size}")), 8 + 4
        if cls == 1:                                            # IEEE float (only the standard layouts)
            if size not in (2, 4, 8):
                return _Dtype(cls, size), 8 + 12
            return _Dtype(cls, size, np.dtype(f"{order}f{size}")), 8 + 12
        if cls == 3:                                            # fixed-length string
            return _Dtype(cls, size, np.dtype(f"S{size}")), 8
        if cls == 9:                                            # variable length
            base, used = self._datatype(body[8:])
            is_str = (bits & 0x0F) == 1
            return _Dtype(cls, size, None, vlen_string=is_str, base=base), 8 + used
        if cls == 7:                                            # reference
            return _Dtype(cls, size, np.dtype(f"<u{size}") if size in (4, 8) else None), 8
        return _Dtype(cls, size), 8

    def _layout(self, body):
        ver = body[0]
        O, L = self.O, self.L
        if ver == 3:
            cls = body[1]
            if cls == 0:
                n = int.from_bytes(body[2:4], "little")
                return ("compact", bytes(body[4:4 + n]))
            if cls == 1:
                return ("contiguous", int.from_bytes(body[2:2 + O], "little"), int.from_bytes(body[2 + O:2 + O + L], "little"))
            if cls == 2:
                rank = body[2]
                addr = int.from_bytes(body[3:3 + O], "little")
                dims = [int.from_bytes(body[3 + O + 4 * i:7 + O + 4 * i], "little") for i in range(rank)]
                return ("chunked", addr, tuple(dims[:-1]))
        elif ver in (1, 2):
            rank, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                addr = int.from_bytes(body[p:p + O], "little"); p += O
            dims = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank)]
            p += 4 * rank
            if cls == 0:
                n = int.from_bytes(body[p:p + 4], "little")
                return ("compact", bytes(body[p + 4:p + 4 + n]))
            if cls == 1:
                return ("contiguous", addr, 0)
            return ("chunked", addr, tuple(dims[:-1]))
        raise HDF5Error(f"unsupported data layout message (version {ver})")

    def _filters(self, body):
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(body[p:p + 2], "little")
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(body[p + 2:p + 4], "little"); p += 4
            else:
                nlen = 0; p += 2
            nvals = int.from_bytes(body[p + 2:p + 4], "little"); p += 4
            if ver == 1:
                nlen = (nlen + 7) & ~7
            p += nlen
            vals = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(nvals)]
            p += 4 * nvals
            if ver == 1 and nvals % 2:
                p += 4
            out.append((fid, vals))
        return out

    def _chunk_btree(self, addr, rank):
        O, b = self.O, self._buf
        p = self.base + addr
        if b[p:p + 4] != b"TREE":
            raise HDF5Error("bad chunk B-tree node")
        level, used = b[p + 5], self._u(p + 6, 2)
        q = p + 8 + 2 * O
        ksz = 8 + 8 * (rank + 1)
        for i in range(used):
            k = q + i * (ksz + O)
            size, mask = self._u(k, 4), self._u(k + 4, 4)
            offs = tuple(self._u(k + 8 + 8 * d, 8) for d in range(rank))
            child = self._u(k + ksz, O)
            if level > 0:
                yield from self._chunk_btree(child, rank)
            else:
                yield offs, size, mask, child

    # ---- attributes
    def _attribute(self, body):
        ver = body[0]
        nsz, dsz, ssz = (int.from_bytes(body[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
        p = 8 if ver < 3 else 9
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = bytes(body[p:p + nsz]).split(b"\0")[0].decode("utf8"); p += pad(nsz)
        dt = self._datatype(body[p:p + dsz])[0]; p += pad(dsz)
        shape = self._dataspace(body[p:p + ssz]) if ssz else (); p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        data = bytes(body[p:])
        if dt.cls == 9 and dt.vlen_string:
            vals = []
            for i in range(n):
                e = data[i * (4 + self.O + 4):]
                ln = int.from_bytes(e[:4], "little")
                col = int.from_bytes(e[4:4 + self.O], "little")
                idx = int.from_bytes(e[4 + self.O:8 + self.O], "little")
                vals.append(self._global_heap(col, idx)[:ln])
            return name, (vals[0] if not shape else np.array(vals, dtype=object).reshape(shape))
        if dt.np_dtype is None:
            return name, None
        arr = np.frombuffer(data[:n * dt.size], dtype=dt.np_dtype).reshape(shape).copy()
        if dt.cls == 3:
            return name, (np.char.rstrip(arr, b"\0") if shape else np.bytes_(bytes(arr[()]).rstrip(b"\0")))
        return name, (arr if shape else arr[()])

    def _global_heap(self, addr, idx):
        b, L = self._buf, self.L
        p = self.base + addr
        if b[p:p + 4] != b"GCOL":
            raise HDF5Error("bad global heap collection")
        size = self._u(p + 8, L)
        q, end = p + 8 + L, p + size
        while q + 8 + L <= end:
            oi = self._u(q, 2)
            osz = self._u(q + 8, L)
            if oi == idx:
                return b[q + 8 + L:q + 8 + L + osz]
            if oi == 0:
                break
            q += 8 + L + ((osz + 7) & ~7)
        raise HDF5Error("global heap object not found")


# ======================================================================================= Keras layout

def _as_str(v):
    return v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def read_keras_weights(path):
    """-> (list of arrays in ``model.get_weights()`` order, list of their Keras weight names).  Accepts a
    weights-only file (``save_weights``, root attribute ``layer_names``) or a full model file (``model.save``,
    group ``model_weights``) - keras/engine/saving.py ``load_weights_from_hdf5_group``."""
    f = File(path)
    g = f
    if "layer_names" not in f.attrs and "model_weights" in f:
        g = f["model_weights"]
    if "layer_names" not in g.attrs:
        raise HDF5Error(f"{path}: no 'layer_names' attribute - not a Keras weight file")
    arrays, names = [], []
    for layer in np.atleast_1d(g.attrs["layer_names"]):
        lg = g[_as_str(layer)]
        for wn in np.atleast_1d(lg.attrs.get("weight_names", [])):
            wn = _as_str(wn)
            arrays.append(np.asarray(lg[wn].read(), dtype=np.float32))
            names.append(wn)
    return arrays, names


# ======================================================================================= writer

def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dt_float32() -> bytes:
    # class 1 (float) version 1; bit field: little-endian, mantissa normalisation 2 (implied msb), sign at bit 31
    return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_string(n: int) -> bytes:
    return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, n)          # class 3 version 1, null-padded (NumPy 'S'), ASCII


def _ds_simple(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_msg(name: str, value) -> bytes:
    nm = name.encode("utf8") + b"\0"
    if isinstance(value, (bytes, str)):
        v = value.encode("utf8") if isinstance(value, str) else value
        dt, ds, data = _dt_string(max(len(v), 1)), struct.pack("<BBBB4x", 1, 0, 0, 0), v.ljust(max(len(v), 1), b"\0")
    else:
        items = [x.encode("utf8") if isinstance(x, str) else bytes(x) for x in value]
        n = max([len(x) for x in items] + [1])
        dt, ds, data = _dt_string(n), _ds_simple((len(items),)), b"".join(x.ljust(n, b"\0") for x in items)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    return _msg(0x0C, body)


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


class _Writer:
    LEAF_K = 16             # symbol table nodes hold up to 2 * LEAF_K entries (recorded in the superblock)
    INTERNAL_K = 16

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)                       # superblock v0 (filled in at the end)

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        a = np.ascontiguousarray(arr, dtype="<f4")
        addr = self.alloc(a.tobytes()) if a.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, a.nbytes)
        return self.alloc(_object_header([_msg(0x01, _ds_simple(a.shape)), _msg(0x03, _dt_float32(), 1),
                                          _msg(0x08, layout)]))

    def group(self, children: dict, attrs: dict) -> int:
        """children: name -> object header address.  One B-tree leaf node + one symbol table node + a local heap."""
        names = sorted(children)
        if len(names) > 2 * self.LEAF_K:
            raise HDF5Error("too many links for one symbol table node")
        heap = bytearray(b"\0" * 8)                            # offset 0: the empty string
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))       # free-list head 1 = none
        snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(names)))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
        snod += b"\0" * (8 + 2 * self.LEAF_K * 40 - len(snod))
        snod_addr = self.alloc(bytes(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        tree += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]] if names else 0)
        tree += b"\0" * (24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8 - len(tree))
        tree_addr = self.alloc(bytes(tree))
        msgs = [_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + [_attr_msg(k, v) for k, v in attrs.items()]
        return self.alloc(_object_header(msgs)), tree_addr, heap_addr

    def finish(self, root) -> bytes:
        root_hdr, tree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", tree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, layers, full_model=False, model_config=None, keras_version="2.2.4"):
    """layers: ordered ``[(layer_name, [(weight_name, array), ...]), ...]`` - every layer of ``model.layers`` in
    order, weight-less ones (InputLayer, ZeroPadding2D) with an empty list, as Keras lists them in ``layer_names``."""
    w = _Writer()
    top = {}
    for lname, weights in layers:
        inner = {}
        for wname, arr in weights:
            parts = wname.split("/")
            node = inner
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = w.dataset(arr)

        def build(tree):
            return w.group({k: (build(v) if isinstance(v, dict) else v) for k, v in tree.items()}, {})[0]
        kids = {k: (build(v) if isinstance(v, dict) else v) for k, v in inner.items()}
        top[lname] = w.group(kids, {"weight_names": [n for n, _ in weights]})[0]
    attrs = {"layer_names": [n for n, _ in layers], "backend": "tensorflow", "keras_version": keras_version}
    if full_model:
        mw = w.group(top, attrs)[0]
        root = w.group({"model_weights": mw}, {"keras_version": keras_version, "backend": "tensorflow",
                                                "model_config": model_config or "{}"})
    else:
        root = w.group(top, attrs)
    data = w.finish(root)
    with open(path, "wb") as fh:
        fh.write(data)
