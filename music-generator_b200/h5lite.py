"""A small reader and writer for the subset of HDF5 that Keras weight files use.

The reference stores its weights as Keras HDF5 (`out/model.h5`: train.py:23 `ModelCheckpoint(MODEL_FILE, ...)`,
util.py:19 `models[0].load_weights(MODEL_FILE)`).  h5py / libhdf5 are not part of this image, so the drop-in reads
and writes the format itself.  What a Keras 2 file written through h5py's default (`libver='earliest'`) contains,
and therefore what is implemented here (HDF5 File Format Specification, version 1.1 / 2.0 numbering):

  * superblock version 0 or 1, 8-byte offsets and lengths (III.A / "Disk Format: Level 0A");
  * old-style groups: symbol-table message -> v1 B-tree of group nodes ("TREE", node type 0) -> symbol-table nodes
    ("SNOD") -> names in a local heap ("HEAP")                                   (Level 1A/1B/1C/1D);
  * version-1 object headers with continuation messages                         (Level 2A);
  * dataspace v1/v2, datatype classes 0 (integer), 1 (float) and 3 (fixed string), fill value, data layout
    v3 contiguous / compact, and chunked layout through the v1 raw-data B-tree with the deflate and shuffle
    filters (only read; Keras never writes chunked weights, other producers do) (Level 2A2);
  * attribute messages v1-v3 holding numeric arrays or arrays of fixed-length strings (`layer_names`,
    `weight_names`); variable-length strings (`backend`, `keras_version`) are resolved through the global heap.

`File(path)` gives a read-only tree: `f.attrs`, `f.keys()`, `f[name]` -> `Group` or `Dataset` (`.shape`, `.dtype`,
`.read()`), `f.visititems(fn)` like h5py.  `write(path, tree)` writes nested dicts of numpy arrays; a dict may carry
attributes under the key `"@attrs"`.  Nothing here is on the GPU path.
"""
from __future__ import annotations

import struct
import zlib
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


# ================================================================================================ reading
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        base = 0
        while base < len(buf) and buf[base:base + 8] != SIG:          # the superblock may sit behind a user block
            base = 512 if base == 0 else base * 2
        if base >= len(buf):
            raise H5Error("not an HDF5 file (no superblock signature at 0, 512, 1024, ...)")
        ver = buf[base + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} (new-style file, libver='latest') is not supported; "
                          "Keras / h5py default files are version 0")
        so, sl = buf[base + 13], buf[base + 14]
        if (so, sl) != (8, 8):
            raise H5Error(f"only 8-byte offsets/lengths are supported (file has {so}/{sl})")
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_addr, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header, cache, _r = struct.unpack_from("<QQII", buf, p)
        self.root_scratch = struct.unpack_from("<QQ", buf, p + 24) if cache == 1 else None

    def at(self, addr: int) -> int:
        return self.base_addr + addr       # file addresses are relative to the base address (= the user block size)

    # ---- object headers (version 1)
    def messages(self, addr: int) -> List[Tuple[int, bytes, int]]:
        b, p = self.b, self.at(addr)
        if b[p:p + 4] == b"OHDR":
            raise H5Error("version-2 object headers (libver='latest') are not supported")
        ver, _, nmsg, _refs, hsize = struct.unpack_from("<BBHII", b, p)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr:#x}")
        out = []
        blocks = [(p + 16, hsize)]
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, q)
                body = b[q + 8:q + 8 + msize]
                if mtype == 0x0010:                                     # continuation: (offset, length)
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(off), ln))
                out.append((mtype, body, flags))
                q += 8 + msize
                left -= 8 + msize
        return out

    # ---- groups
    def group_links(self, btree: int, heap: int) -> Dict[str, int]:
        b = self.b
        hp = self.at(heap)
        if b[hp:hp + 4] != b"HEAP":
            raise H5Error(f"no local heap at {heap:#x}")
        _dsize, _free, daddr = struct.unpack_from("<QQQ", b, hp + 8)
        data0 = self.at(daddr)

        def name(off):
            end = b.index(b"\0", data0 + off)
            return b[data0 + off:end].decode("utf8")

        links: Dict[str, int] = {}

        def walk(addr):
            p = self.at(addr)
            if b[p:p + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, p + 6)[0]
                for i in range(n):
                    noff, ohdr = struct.unpack_from("<QQ", b, p + 8 + 40 * i)
                    links[name(noff)] = ohdr
                return
            if b[p:p + 4] != b"TREE":
                raise H5Error(f"no B-tree / symbol node at {addr:#x}")
            ntype, _lvl, used = struct.unpack_from("<BBH", b, p + 4)
            if ntype != 0:
                raise H5Error("group B-tree expected")
            q = p + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, q + 8 + 16 * i)[0]   # key_i (8), child_i (8), ...
                walk(child)

        walk(btree)
        return links

    # ---- datatypes
    @staticmethod
    def dtype_of(body: bytes):
        cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", body, 0)
        cls = cv & 0x0F
        order = ">" if (b0 & 1) else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if (b0 & 8) else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:
            return ("vlen", size, _Reader.dtype_of(body[8:]) if (b0 & 0x0F) == 0 else "str")
        return ("opaque", size)

    @staticmethod
    def shape_of(body: bytes) -> Tuple[int, ...]:
        ver, rank = body[0], body[1]
        p = 8 if ver == 1 else 4
        return struct.unpack_from("<" + "Q" * rank, body, p) if rank else ()

    def global_heap_object(self, addr: int, index: int) -> bytes:
        b, p = self.b, self.at(addr)
        if b[p:p + 4] != b"GCOL":
            raise H5Error(f"no global heap collection at {addr:#x}")
        size = struct.unpack_from("<Q", b, p + 8)[0]
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, _refs, _r, osize = struct.unpack_from("<HHIQ", b, q)
            if idx == 0:
                break
            if idx == index:
                return b[q + 16:q + 16 + osize]
            q += 16 + ((osize + 7) & ~7)
        raise H5Error("global heap object not found")

    def decode(self, dt, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, np.dtype):
            a = np.frombuffer(raw, dtype=dt, count=n).reshape(shape)
            return a.copy()
        if dt[0] == "vlen":
            out = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                data = self.global_heap_object(gaddr, gidx) if ln else b""
                out.append(data[:ln] if dt[2] == "str" else np.frombuffer(data, dtype=dt[2], count=ln).copy())
            return out[0] if not shape else out
        return None

    def attribute(self, body: bytes):
        ver = body[0]
        nsz, dsz, ssz = struct.unpack_from("<HHH", body, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        name = body[p:p + nsz].split(b"\0")[0].decode("utf8")
        p += pad(nsz)
        dt = self.dtype_of(body[p:p + dsz])
        p += pad(dsz)
        shape = self.shape_of(body[p:p + ssz])
        p += pad(ssz)
        return name, self.decode(dt, shape, body[p:])

    # ---- chunked storage (v1 B-tree, node type 1)
    def read_chunked(self, btree: int, shape, chunk, dt: np.dtype, filters) -> np.ndarray:
        b = self.b
        out = np.zeros(shape, dtype=dt)
        rank = len(shape)

        def walk(addr):
            p = self.at(addr)
            if b[p:p + 4] != b"TREE":
                raise H5Error(f"no chunk B-tree at {addr:#x}")
            ntype, lvl, used = struct.unpack_from("<BBH", b, p + 4)
            ksz = 8 + 8 * (rank + 1)
            q = p + 24
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, q)
                offs = struct.unpack_from("<" + "Q" * (rank + 1), b, q + 8)[:rank]
                child = struct.unpack_from("<Q", b, q + ksz)[0]
                if lvl > 0:
                    walk(child)
                else:
                    raw = b[self.at(child):self.at(child) + csize]
                    for k, (fid, _cd) in enumerate(reversed(filters)):
                        if fmask & (1 << (len(filters) - 1 - k)):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = dt.itemsize
                            raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                        else:
                            raise H5Error(f"unsupported filter {fid}")
                    tile = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunk))).reshape(chunk)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                    out[sl] = tile[tuple(slice(0, s.stop - s.start) for s in sl)]
                q += ksz + 8

        if btree != UNDEF:
            walk(btree)
        return out


class _Node:
    def __init__(self, rd: _Reader, addr: int, name: str):
        self._rd, self._addr, self.name = rd, addr, name
        self._msgs = rd.messages(addr)
        self.attrs: Dict[str, object] = {}
        for t, body, _ in self._msgs:
            if t == 0x000C:
                try:
                    k, v = rd.attribute(body)
                    self.attrs[k] = v
                except (H5Error, struct.error, ValueError):
                    pass

    def _msg(self, t) -> Optional[bytes]:
        for mt, body, _ in self._msgs:
            if mt == t:
                return body
        return None


class Dataset(_Node):
    def __init__(self, rd, addr, name):
        super().__init__(rd, addr, name)
        self.dtype = rd.dtype_of(self._msg(0x0003))
        self.shape = tuple(rd.shape_of(self._msg(0x0001)))

    def read(self) -> np.ndarray:
        rd, lay = self._rd, self._msg(0x0008)
        if not isinstance(self.dtype, np.dtype):
            raise H5Error(f"{self.name}: datatype {self.dtype} is not readable as an array")
        ver, cls = lay[0], lay[1]
        n = int(np.prod(self.shape)) if self.shape else 1
        if ver in (1, 2):           # HDF5 1.6 writers: dimensionality, class, 5 reserved bytes, address, 4-byte dimensions
            nd, cls = lay[1], lay[2]
            if cls == 1:
                p = rd.at(struct.unpack_from("<Q", lay, 8)[0])
                return np.frombuffer(rd.b[p:p + n * self.dtype.itemsize], dtype=self.dtype, count=n).reshape(self.shape).copy()
            if cls == 2:
                btree = struct.unpack_from("<Q", lay, 8)[0]
                dims = struct.unpack_from("<" + "I" * nd, lay, 16)
                return rd.read_chunked(btree, self.shape, dims[:-1], self.dtype, self._filters())
            size = struct.unpack_from("<I", lay, 8 + 4 * nd)[0]
            return np.frombuffer(lay[12 + 4 * nd:12 + 4 * nd + size], dtype=self.dtype, count=n).reshape(self.shape).copy()
        if ver != 3:
            raise H5Error(f"{self.name}: data layout message version {ver}")
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == UNDEF:
                return np.zeros(self.shape, self.dtype)
            p = rd.at(addr)
            return np.frombuffer(rd.b[p:p + size], dtype=self.dtype, count=n).reshape(self.shape).copy()
        if cls == 0:
            size = struct.unpack_from("<H", lay, 2)[0]
            return np.frombuffer(lay[4:4 + size], dtype=self.dtype, count=n).reshape(self.shape).copy()
        if cls == 2:
            rank1 = lay[2]
            btree = struct.unpack_from("<Q", lay, 3)[0]
            dims = struct.unpack_from("<" + "I" * rank1, lay, 11)
            return rd.read_chunked(btree, self.shape, dims[:-1], self.dtype, self._filters())
        raise H5Error(f"{self.name}: layout class {cls}")

    def _filters(self):
        """Filter pipeline message (0x000B) -> [(filter id, client data)] in application order."""
        fm = self._msg(0x000B)
        if fm is None:
            return []
        fver, nf = fm[0], fm[1]
        p = 8 if fver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", fm, p)[0]
            if fver == 2 and fid < 256:          # v2 drops the name length of the predefined filters
                _flags, ncd = struct.unpack_from("<HH", fm, p + 2)
                nlen, p = 0, p + 6
            else:
                nlen, _flags, ncd = struct.unpack_from("<HHH", fm, p + 2)
                p += 8
            p += ((nlen + 7) & ~7) if fver == 1 else nlen
            cd = struct.unpack_from("<" + "I" * ncd, fm, p)
            p += 4 * ncd + (4 if (fver == 1 and ncd % 2) else 0)
            out.append((fid, cd))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)


class Group(_Node):
    def __init__(self, rd, addr, name, scratch=None):
        super().__init__(rd, addr, name)
        st = self._msg(0x0011)
        if st is not None:
            btree, heap = struct.unpack_from("<QQ", st, 0)
        elif scratch is not None:
            btree, heap = scratch
        else:
            raise H5Error(f"{name or '/'}: not an old-style group (no symbol table message)")
        self._links = rd.group_links(btree, heap)

    def keys(self) -> List[str]:
        return sorted(self._links)

    def __contains__(self, k: str) -> bool:
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._child(part)
        return node

    def _child(self, k: str):
        rd, addr = self._rd, self._links[k]
        name = f"{self.name}/{k}" if self.name else k
        types = {t for t, _, _ in rd.messages(addr)}
        if 0x0011 in types:
            return Group(rd, addr, name)
        if 0x0008 in types:
            return Dataset(rd, addr, name)
        raise H5Error(f"{name}: neither a group nor a dataset (committed datatype?)")

    def visititems(self, fn: Callable[[str, object], object]):
        for k in self.keys():
            try:
                c = self._child(k)
            except H5Error:
                continue
            r = fn(c.name, c)
            if r is not None:
                return r
            if isinstance(c, Group):
                r = c.visititems(fn)
                if r is not None:
                    return r
        return None


class File(Group):
    """Read-only view of an HDF5 file (the subset in the module docstring)."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            rd = _Reader(f.read())
        super().__init__(rd, rd.root_header, "", rd.root_scratch)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ================================================================================================ writing
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        exp_bits, mant_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bits = 8 * dt.itemsize
        return (struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize) +
                struct.pack("<HHBBBBI", 0, bits, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1))
    if dt.kind in "iu":
        return (struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) +
                struct.pack("<HH", 0, 8 * dt.itemsize))
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)         # null-padded ASCII, like h5py for numpy 'S'
    raise H5Error(f"cannot write dtype {dt}")


def _space_msg(shape) -> bytes:
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _attr_msg(name: str, value) -> bytes:
    if isinstance(value, (bytes, str)):
        v = value.encode("utf8") if isinstance(value, str) else value
        a = np.array(v, dtype=f"S{max(len(v), 1)}")
    elif isinstance(value, (list, tuple)) and value and isinstance(value[0], (bytes, str)):
        vals = [x.encode("utf8") if isinstance(x, str) else x for x in value]
        a = np.array(vals, dtype=f"S{max(max(len(x) for x in vals), 1)}")
    else:
        a = np.asarray(value)
        if a.dtype == np.float64 and not isinstance(value, np.ndarray):
            a = a.astype(np.float64)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    nm = name.encode("utf8") + b"\0"
    dtm, spm = _dtype_msg(a.dtype), _space_msg(a.shape)
    body = (struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) +
            np.ascontiguousarray(a).tobytes())
    if len(body) > 65000:
        raise H5Error(f"attribute {name} is too large for an object-header message")
    return body


def _header(msgs: List[Tuple[int, bytes]]) -> bytes:
    """Version-1 object header holding `msgs` [(type, body)] in one chunk."""
    data = b"".join(struct.pack("<HHBBBB", t, len(_pad8(b)), 0, 0, 0, 0) + _pad8(b) for t, b in msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(data)) + b"\0" * 4 + data


class _Writer:
    LEAF_K = 64         # symbol-table node capacity 2K = 128 entries: one node per group is enough for a Keras file
    INTERNAL_K = 16

    def __init__(self):
        self.buf = bytearray(96)     # superblock filled in at the end

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, a: np.ndarray, attrs: dict) -> int:
        a = np.asarray(a)
        shape = a.shape                      # (ascontiguousarray would turn a scalar into a 1-element vector)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        raw = np.ascontiguousarray(a).tobytes()
        daddr = self.alloc(raw) if raw else UNDEF
        msgs = [(0x0001, _space_msg(shape)), (0x0003, _dtype_msg(a.dtype)),
                (0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),                      # fill value v2: late alloc, if-set, default
                (0x0008, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]                # contiguous layout v3
        msgs += [(0x000C, _attr_msg(k, v)) for k, v in attrs.items()]
        return self.alloc(_header(msgs))

    def group(self, tree: dict) -> Tuple[int, int, int]:
        """Returns (object header address, B-tree address, local heap address)."""
        attrs = tree.get("@attrs", {})
        children = {}
        for k, v in tree.items():
            if k == "@attrs":
                continue
            if isinstance(v, dict):
                children[k] = self.group(v)[0]
            else:
                children[k] = self.dataset(np.asarray(v), {})
        if len(children) > 2 * self.LEAF_K:
            raise H5Error(f"a group with {len(children)} members exceeds one symbol-table node ({2 * self.LEAF_K})")
        names = sorted(children, key=lambda s: s.encode("utf8"))      # B-tree order: strcmp on the raw bytes
        # local heap data: the empty string at offset 0, then every name, 8-byte aligned
        heap = bytearray(8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), 1, heap_data))   # 1: no free block
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for n in names:
            snod += struct.pack("<QQII", offs[n], children[n], 0, 0) + b"\0" * 16
        snod += b"\0" * (8 + 40 * 2 * self.LEAF_K - len(snod))
        snod_addr = self.alloc(bytes(snod))
        K = self.INTERNAL_K
        tree_node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            tree_node += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])     # key0 = "", child0, key1 = last name
        tree_node += b"\0" * (24 + (2 * K + 1) * 8 + 2 * K * 8 - len(tree_node))
        btree_addr = self.alloc(bytes(tree_node))
        msgs = [(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [(0x000C, _attr_msg(k, v)) for k, v in attrs.items()]
        return self.alloc(_header(msgs)), btree_addr, heap_addr

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        self.buf += b"\0" * (-len(self.buf) % 8)
        hdr, btree, heap = root
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write(path: str, tree: dict) -> None:
    """Write nested dicts of arrays as an HDF5 file (superblock v0, old-style groups, contiguous datasets).  A dict's
    `"@attrs"` entry holds that group's attributes (bytes / str / list of them / numeric arrays)."""
    w = _Writer()
    data = w.finish(w.group(tree))
    with open(path, "wb") as f:
        f.write(data)
