"""A minimal pure-Python HDF5 reader and writer: just enough of the format for the ``model.weights.h5``
member of a Keras-3 ``.keras`` archive (h5py is not installable in this image).

What Keras 3.3.3 writes through ``h5py.File(path, "w")`` with h5py's defaults (``saving_lib.H5IOStore``,
used by ``Model.save`` / ``ModelCheckpoint`` at /root/reference/Super_resolution/code/train_adaptive_unet.py:617 and
read back by ``load_weights`` / ``load_model`` at :511-516 and evaluate_model.py:57-91) is the *classic* subset of
the format, restated here from the HDF5 File Format Specification (version 2.0):

  * superblock version 0 (8-byte offsets and lengths), optionally behind a user block (base address);
  * groups as symbol tables: a version-1 object header with a Symbol Table message (0x0011) pointing at a
    version-1 B-tree ("TREE", node type 0) whose leaves are symbol-table nodes ("SNOD") and at a local heap
    ("HEAP") holding the link names;
  * datasets: version-1 object headers with Dataspace (0x0001), Datatype (0x0003: fixed-point / IEEE float,
    little endian), Fill Value (0x0005) and Data Layout (0x0008, version 3: contiguous or compact) messages,
    continuation blocks (0x0010) followed; attributes and other messages are skipped.

Not supported (clear errors): version-2 object headers / new-style (link-message) groups, chunked or filtered
datasets, big-endian, compound / string / variable-length types.

The reader is pinned against a file written by the real HDF5 library (scipy ships a MATLAB v7.3 = HDF5 test
file; tests/test_h5lite_cpu.py), the writer against the reader and against the structural rules of the
specification (``validate``).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16          # group B-tree parameters written into (and read from) the superblock

Tree = Dict[str, Union["Tree", np.ndarray]]


class H5Error(ValueError):
    pass


# =============================================================================================== reader
class _Dataset:
    def __init__(self, shape, dtype, address, size, inline=None):
        self.shape, self.dtype, self.address, self.size, self.inline = shape, dtype, address, size, inline


class H5Reader:
    """Read-only view of an HDF5 file held in memory: ``tree()`` -> nested dict of numpy arrays."""

    def __init__(self, data: bytes):
        self.d = bytes(data)
        off = 0
        while True:                      # the superblock sits at 0, 512, 1024, 2048, ... (user block in front)
            if self.d[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.d):
                raise H5Error("not an HDF5 file (no superblock signature)")
        sb = off
        ver = self.d[sb + 8]
        if ver not in (0, 1):
            raise H5Error(f"HDF5 superblock version {ver}: only the classic versions 0 / 1 (h5py's default) are read; "
                          "re-save with libver='earliest'")
        so, sl = self.d[sb + 13], self.d[sb + 14]
        if (so, sl) != (8, 8):
            raise H5Error(f"offsets / lengths of {so} / {sl} bytes are not supported (8 / 8 expected)")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.d, sb + 16)
        p = sb + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", self.d, p)
        if self.base == 0 and sb != 0:
            self.base = sb               # files whose base address field is 0 behind a user block
        ste = p + 32
        _name_off, self.root_addr, cache, _res = struct.unpack_from("<QQII", self.d, ste)
        self._objs: Dict[int, object] = {}

    # ---- low level ------------------------------------------------------------------------------
    def _at(self, addr: int) -> int:
        if addr == UNDEF:
            raise H5Error("undefined address dereferenced")
        a = self.base + addr
        if a >= len(self.d):
            raise H5Error(f"address {addr:#x} lies beyond the end of the file")
        return a

    def _messages(self, addr: int) -> List[Tuple[int, int, bytes]]:
        """(type, flags, data) of every message of the version-1 object header at `addr` (continuations followed)."""
        a = self._at(addr)
        if self.d[a:a + 4] == b"OHDR":
            raise H5Error("version-2 object header: the file was written with libver='latest'; only the classic "
                          "format (h5py default) is supported")
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", self.d, a)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr:#x}")
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.d, p)
                body = self.d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                                   # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._at(caddr), clen))
                out.append((mtype, flags, body))
        return out

    # ---- groups ------------------------------------------------------------------------------------
    def _heap_name(self, heap_data: int, off: int) -> str:
        a = heap_data + off
        e = self.d.index(b"\0", a)
        return self.d[a:e].decode("utf-8")

    def _group_entries(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        h = self._at(heap)
        if self.d[h:h + 4] != b"HEAP":
            raise H5Error(f"local heap signature missing at {heap:#x}")
        _seg_size, _free, seg_addr = struct.unpack_from("<QQQ", self.d, h + 8)
        heap_data = self._at(seg_addr)
        out: List[Tuple[str, int]] = []

        def walk(node):
            a = self._at(node)
            if self.d[a:a + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", self.d, a + 6)[0]
                for i in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", self.d, a + 8 + 40 * i)
                    out.append((self._heap_name(heap_data, noff), ohdr))
                return
            if self.d[a:a + 4] != b"TREE":
                raise H5Error(f"B-tree node signature missing at {node:#x}")
            ntype, _level, used = struct.unpack_from("<BBH", self.d, a + 4)
            if ntype != 0:
                raise H5Error("chunked-dataset B-tree where a group B-tree was expected")
            p = a + 24
            for i in range(used):
                child = struct.unpack_from("<Q", self.d, p + 8 + 16 * i)[0]
                walk(child)

        walk(btree)
        return out

    def _object(self, addr: int):
        if addr in self._objs:
            return self._objs[addr]
        msgs = self._messages(addr)
        types = {t for t, _, _ in msgs}
        if 0x0011 in types:
            body = next(b for t, _, b in msgs if t == 0x0011)
            btree, heap = struct.unpack_from("<QQ", body, 0)
            obj = dict(self._group_entries(btree, heap))
        elif 0x0002 in types or 0x0006 in types:
            raise H5Error("new-style group (link messages): the file was written with libver='latest' or track_order; "
                          "only symbol-table groups (h5py default) are supported")
        elif 0x0008 in types:
            obj = self._dataset(msgs)
        else:
            raise H5Error(f"object at {addr:#x} is neither a group nor a dataset")
        self._objs[addr] = obj
        return obj

    # ---- datasets ------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(body: bytes) -> np.dtype:
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits0 = body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        if bits0 & 1:
            raise H5Error("big-endian datasets are not supported")
        if cls == 0:                     # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(f"<{'i' if signed else 'u'}{size}")
        if cls == 1:                     # IEEE floating point
            if size not in (2, 4, 8):
                raise H5Error(f"{size}-byte floating-point type")
            return np.dtype(f"<f{size}")
        raise H5Error(f"datatype class {cls} (only fixed-point and floating-point datasets are supported)")

    def _dataset(self, msgs) -> _Dataset:
        shape, dtype, layout = (), None, None
        for t, _f, b in msgs:
            if t == 0x0001:
                ver, rank = b[0], b[1]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", b, p) if rank else ()
            elif t == 0x0003:
                dtype = self._dtype(b)
            elif t == 0x0008:
                layout = b
            elif t == 0x000B:
                raise H5Error("filtered (compressed) datasets are not supported")
        if dtype is None or layout is None:
            raise H5Error("dataset without datatype or layout message")
        ver = layout[0]
        if ver in (1, 2):                # HDF5 <= 1.6: version, dimensionality, class, 5 reserved, [address], 4-byte sizes
            ndim, cls = layout[1], layout[2]
            count = int(np.prod(shape)) if shape else 1
            if cls == 1:
                return _Dataset(shape, dtype, struct.unpack_from("<Q", layout, 8)[0], count * dtype.itemsize)
            if cls == 0:
                p = 8 + 4 * ndim
                n = struct.unpack_from("<I", layout, p)[0]
                return _Dataset(shape, dtype, None, n, layout[p + 4:p + 4 + n])
            raise H5Error("chunked datasets are not supported (keras writes contiguous ones)")
        if ver != 3:
            raise H5Error(f"data layout message version {ver} (1-3 expected)")
        cls = layout[1]
        if cls == 0:                     # compact: data inside the header
            n = struct.unpack_from("<H", layout, 2)[0]
            return _Dataset(shape, dtype, None, n, layout[4:4 + n])
        if cls == 1:                     # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            return _Dataset(shape, dtype, addr, size)
        raise H5Error("chunked datasets are not supported (keras writes contiguous ones)")

    def _array(self, ds: _Dataset) -> np.ndarray:
        count = int(np.prod(ds.shape)) if ds.shape else 1
        nbytes = count * ds.dtype.itemsize
        if ds.inline is not None:
            raw = ds.inline[:nbytes]
        elif count == 0 or ds.address == UNDEF:
            raw = b"\0" * nbytes         # never written: the fill value (zeros)
        else:
            a = self._at(ds.address)
            raw = self.d[a:a + nbytes]
        if len(raw) != nbytes:
            raise H5Error("dataset extends beyond the end of the file")
        return np.frombuffer(raw, dtype=ds.dtype).reshape(ds.shape).copy()

    # ---- public ----------------------------------------------------------------------------------------
    def tree(self) -> Tree:
        def build(addr, depth=0):
            if depth > 64:
                raise H5Error("group nesting too deep (cycle?)")
            obj = self._object(addr)
            if isinstance(obj, _Dataset):
                return self._array(obj)
            return {name: build(child, depth + 1) for name, child in obj.items()}

        return build(self.root_addr)

    def datasets(self) -> Iterator[Tuple[str, np.ndarray]]:
        def walk(node, prefix):
            for k, v in node.items():
                path = f"{prefix}/{k}" if prefix else k
                if isinstance(v, dict):
                    yield from walk(v, path)
                else:
                    yield path, v

        yield from walk(self.tree(), "")


def read_h5(data: bytes) -> Tree:
    return H5Reader(data).tree()


# =============================================================================================== writer
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


_FLOAT_PROPS = {   # size: (sign location, exponent location, exponent size, mantissa size, bias)
    2: (15, 10, 5, 10, 15), 4: (31, 23, 8, 23, 127), 8: (63, 52, 11, 52, 1023),
}


def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise H5Error("big-endian arrays are not written")
    if dt.kind == "f":
        sign, eloc, esize, msize, bias = _FLOAT_PROPS[dt.itemsize]
        # class 1, version 1; bit field: little endian, mantissa normalisation 2 (implied MSB), sign location
        head = bytes([0x11, 0x20, sign, 0x00]) + struct.pack("<I", dt.itemsize)
        props = struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esize, 0, msize, bias)
        return head + props
    if dt.kind in "iu":
        head = bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise H5Error(f"dtype {dt} is not written (float16/32/64 and integers only)")


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: List[bytes]) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)         # superblock (version 0, 8-byte offsets / lengths): filled in at the end

    def _alloc(self, blob: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += blob
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.asarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = np.ascontiguousarray(arr).tobytes()
        data_addr = self._alloc(raw) if raw else UNDEF
        rank = arr.ndim
        space = struct.pack("<BBBB4x", 1, rank, 0, 0) + struct.pack(f"<{rank}Q", *arr.shape)
        msgs = [
            _message(0x0001, space),
            _message(0x0003, _datatype_message(arr.dtype), flags=1),
            # fill value, byte for byte what the HDF5 library wrote into the real file the reader is pinned on:
            # version 1, late allocation, written "if set", default value (defined, size 0)
            _message(0x0005, bytes([1, 2, 2, 1, 0, 0, 0, 0]), flags=1),
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),
        ]
        return self._alloc(_object_header(msgs))

    def group(self, children: Dict[str, int]) -> Tuple[int, int, int]:
        """Symbol-table group over {name: object header address} -> (object header, B-tree, heap) addresses."""
        names = sorted(children, key=lambda s: s.encode("utf-8"))     # B-tree order = strcmp on the raw bytes
        # local heap: offset 0 = the empty string (8 bytes), then the names, each NUL-terminated and 8-byte aligned,
        # then one free block (>= 16 bytes: "next" = 1 = none, size)
        seg = bytearray(8)
        offs = {}
        for n in names:
            offs[n] = len(seg)
            b = n.encode("utf-8") + b"\0"
            seg += b + b"\0" * (_pad8(len(b)) - len(b))
        free_off = len(seg)
        free_size = max(32, _pad8(len(seg) // 4))
        seg += struct.pack("<QQ", 1, free_size) + b"\0" * (free_size - 16)
        seg_addr = self._alloc(bytes(seg))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # leaves: symbol-table nodes of at most 2*LEAF_K entries (all slots allocated)
        leaves = []
        per = 2 * LEAF_K
        for i in range(0, len(names), per):     # (an empty group is a B-tree node with no entries and no leaf)
            part = names[i:i + per]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                body += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            body += b"\0" * (40 * (per - len(part)))
            leaves.append((self._alloc(body), offs[part[-1]] if part else 0))

        def node(level: int, kids: List[Tuple[int, int]]) -> Tuple[int, int]:
            """One B-tree node over (child address, heap offset of the child's greatest name) pairs."""
            body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(kids), UNDEF, UNDEF) + struct.pack("<Q", 0)
            for addr, last in kids:
                body += struct.pack("<QQ", addr, last)
            body += b"\0" * (16 * (2 * INTERNAL_K - len(kids)))
            return self._alloc(body), (kids[-1][1] if kids else 0)

        level, nodes, fan = 0, leaves, 2 * INTERNAL_K
        while True:
            nodes = [node(level, nodes[i:i + fan]) for i in range(0, max(len(nodes), 1), fan)]
            if len(nodes) == 1:
                break
            level += 1
        btree_addr = nodes[0][0]
        ohdr = self._alloc(_object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr), flags=1),
                                           _message(0x0000, b"")]))
        return ohdr, btree_addr, heap_addr

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        while len(self.buf) % 8:
            self.buf.append(0)
        ohdr, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(tree: Tree) -> bytes:
    """Nested dict (groups) of numpy arrays (datasets) -> the bytes of a classic-format HDF5 file."""
    w = _Writer()

    def emit(node) -> Tuple[int, int, int]:
        kids = {}
        for name, v in node.items():
            if not name or "/" in name:
                raise H5Error(f"invalid link name {name!r}")
            kids[name] = emit(v)[0] if isinstance(v, dict) else w.dataset(v)
        return w.group(kids)

    return w.finish(emit(tree))


# =============================================================================================== validator
def validate(data: bytes) -> int:
    """Structural checks of the specification over every object reachable from the root (stricter than the reader):
    8-byte alignment of every structure, B-tree keys sorted and equal to the greatest name of the child to their left,
    symbol-table nodes within their 2K capacity and sorted, heap offsets inside the data segment and NUL-terminated,
    object-header sizes consistent, dataset extents inside the file, end-of-file address == file size.
    Returns the number of objects visited."""
    r = H5Reader(data)
    d = r.d
    if r.eof not in (len(d), len(d) - r.base):     # (the library stores it as an absolute address behind a user block)
        raise H5Error(f"end-of-file address {r.eof:#x} != file size {len(d):#x}")
    seen = set()

    def check_group(btree, heap):
        h = r._at(heap)
        assert heap % 8 == 0 and btree % 8 == 0, "unaligned group structure"
        seg_size, free_head, seg_addr = struct.unpack_from("<QQQ", d, h + 8)
        seg = r._at(seg_addr)
        assert seg + seg_size <= len(d), "heap data segment beyond end of file"
        assert d[seg] == 0, "heap offset 0 must hold the empty string"
        f = free_head                    # the free list must stay inside the segment and terminate
        hops = 0
        while f != 1:
            assert f % 8 == 0 and f + 16 <= seg_size, "heap free block outside the data segment"
            nxt, size = struct.unpack_from("<QQ", d, seg + f)
            assert size >= 16 and f + size <= seg_size, "heap free block size"
            f, hops = nxt, hops + 1
            assert hops < 1 << 16, "heap free list does not terminate"

        def name_at(off):
            assert off < seg_size, "name offset outside the heap"
            e = d.index(b"\0", seg + off)
            assert e < seg + seg_size, "unterminated name"
            return d[seg + off:e]

        def walk(node, level_expected):
            a = r._at(node)
            assert node % 8 == 0
            if d[a:a + 4] == b"SNOD":
                assert level_expected in (None, -1), "leaf at a wrong level"
                ver, _, nsym = struct.unpack_from("<BBH", d, a + 4)
                assert ver == 1 and nsym <= 2 * r.leaf_k, "symbol-table node over capacity"
                names = [name_at(struct.unpack_from("<Q", d, a + 8 + 40 * i)[0]) for i in range(nsym)]
                assert names == sorted(names) and len(set(names)) == len(names), "symbol-table node not sorted"
                return names
            assert d[a:a + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", d, a + 4)
            assert ntype == 0 and used <= 2 * r.internal_k, "B-tree node over capacity"
            if level_expected is not None and level_expected >= 0:
                assert level == level_expected, "B-tree level"
            keys = [struct.unpack_from("<Q", d, a + 24 + 16 * i)[0] for i in range(used + 1)]
            allnames = []
            for i in range(used):
                child = struct.unpack_from("<Q", d, a + 32 + 16 * i)[0]
                names = walk(child, level - 1 if level > 0 else -1)
                if names:
                    assert name_at(keys[i + 1]) == names[-1], "B-tree key is not the greatest name of its left child"
                    if i > 0 or keys[0] != 0:
                        assert name_at(keys[i]) < names[0], "B-tree key order"
                allnames += names
            assert allnames == sorted(allnames), "B-tree children out of order"
            return allnames

        walk(btree, None)

    def visit(addr):
        if addr in seen:
            return
        seen.add(addr)
        assert addr % 8 == 0, "unaligned object header"
        a = r._at(addr)
        ver, _, nmsg, _ref, hsize = struct.unpack_from("<BBHII", d, a)
        assert ver == 1 and a + 16 + hsize <= len(d)
        msgs = r._messages(addr)
        assert len(msgs) == nmsg, "message count"
        assert all(len(b) % 8 == 0 for _, _, b in msgs), "message size not a multiple of 8"
        obj = r._object(addr)
        if isinstance(obj, _Dataset):
            count = int(np.prod(obj.shape)) if obj.shape else 1
            if obj.inline is None and count:
                assert obj.size == count * obj.dtype.itemsize, "layout size != extent"
                assert r._at(obj.address) + obj.size <= len(d), "dataset beyond end of file"
        else:
            body = next(b for t, _, b in msgs if t == 0x0011)
            check_group(*struct.unpack_from("<QQ", body, 0))
            for child in obj.values():
                visit(child)

    visit(r.root_addr)
    return len(seen)
