"""Minimal HDF5 reader (and a fixture writer) for the captured view sets -- SURVEY.md 8(f) N4.

The reference loads its captures with h5py (captured_data.py:94-108, 136-149: `h5py.File(path, 'r')`, then `h5data[name][i]` /
`h5data[name][:]` on six numeric datasets).  h5py / libhdf5 are not in this image, so this module reads the subset of the HDF5
file format those files use, straight from the published format specification ("HDF5 File Format Specification Version 3.0"):

  superblock v0-v3 (user block: searched at 0, 512, 1024, ...)     object headers v1 and v2 (+ continuation blocks)
  old-style groups (symbol table: B-tree v1 + local heap + SNOD)   new-style compact groups (link messages)
  dataspace v1 / v2 (simple)                                       datatypes: fixed point, IEEE float, enum of those (h5py bool)
  data layout v1-v3: compact, contiguous, chunked (B-tree v1)      filters: deflate, shuffle, fletcher32 (checksum skipped)

Not covered (raises NotImplementedError, never guesses): dense groups (fractal heap), layout v4 chunk indices, compound /
variable-length / string types, szip / n-bit / scale-offset / third-party filters.

`File` is a read-only mapping with h5py's call shapes as the loaders use them: `f['name']` -> `Dataset` with `.shape`, `.dtype`
and `ds[i]`, `ds[:]`, `ds[a:b]`, `ds[()]`; contiguous datasets are served from a memory map (reading view i of a 2 GB
`screen_position` touches only that view), chunked ones decompress only the chunks that overlap the request.

`write_h5` writes the same subset (superblock v0, one symbol-table root group, contiguous or chunked + shuffle + deflate
datasets); tests use it for fixtures and `tools`-style conversions of .npz view sets.  It has not been checked against libhdf5
(absent here); the READER is checked against a file libhdf5 itself wrote (tests/test_h5lite_cpu.py).
"""
import mmap
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class File:
    def __init__(self, path, mode="r"):
        if mode != "r":
            raise ValueError("h5lite.File is read-only")
        self._fh = open(path, "rb")
        self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self.filename = str(path)
        self._parse_superblock()
        self._root = Group(self, self._root_addr, "/")

    # -- mapping interface (h5py.File) --
    def __getitem__(self, name):
        return self._root[name]

    def __contains__(self, name):
        return name in self._root

    def keys(self):
        return self._root.keys()

    def __iter__(self):
        return iter(self._root.keys())

    def __len__(self):
        return len(self._root.keys())

    def close(self):
        if self._buf is not None:
            try:
                self._buf.close()
            except BufferError:  # arrays handed out still view the map: leave it to the garbage collector
                pass
            self._buf = None
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- format --
    def _parse_superblock(self):
        buf, size = self._buf, len(self._buf)
        off = 0
        while off + 8 <= size and buf[off:off + 8] != SIGNATURE:
            off = 512 if off == 0 else off * 2
        if off + 8 > size:
            raise H5FormatError("not an HDF5 file (no superblock signature)")
        ver = buf[off + 8]
        if ver in (0, 1):
            self.O, self.L = buf[off + 13], buf[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = _u(buf, p, self.O)
            p += 4 * self.O  # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
            self._root_addr = _u(buf, p + self.O, self.O)
        elif ver in (2, 3):
            self.O, self.L = buf[off + 9], buf[off + 10]
            p = off + 12
            self.base = _u(buf, p, self.O)
            self._root_addr = _u(buf, p + 3 * self.O, self.O)
        else:
            raise H5FormatError(f"unknown superblock version {ver}")
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5FormatError("unsupported size of offsets / lengths")
        # the base address of a file with a user block is the superblock's own offset; some writers leave the field at 0
        if self.base == 0 and off:
            self.base = off if self._looks_like_header(self._root_addr + off) else 0

    def _looks_like_header(self, pos):
        return pos + 4 <= len(self._buf) and (self._buf[pos:pos + 4] == b"OHDR" or self._buf[pos] == 1)

    def _undef(self, a):
        return a == (1 << (8 * self.O)) - 1

    def _messages(self, addr):
        """-> list of (type, flags, bytes) of the object header at `addr` (continuation blocks followed)."""
        buf = self._buf
        pos = addr + self.base
        out = []
        if buf[pos:pos + 4] == b"OHDR":  # version 2
            flags = buf[pos + 5]
            p = pos + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            n = 1 << (flags & 3)
            chunk = _u(buf, p, n)
            p += n
            blocks = [(p, p + chunk)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = buf[p], _u(buf, p + 1, 2), buf[p + 3]
                    p += 4 + (2 if track_order else 0)
                    if p + msize > end:
                        break
                    data = bytes(buf[p:p + msize])
                    p += msize
                    if mtype == 0x10:
                        o, ln = _u(data, 0, self.O), _u(data, self.O, self.L)
                        q = o + self.base
                        if buf[q:q + 4] != b"OCHK":
                            raise H5FormatError("bad object header continuation block")
                        blocks.append((q + 4, q + ln - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        if buf[pos] != 1:
            raise H5FormatError(f"no object header at {addr:#x}")
        n_msgs, hsize = _u(buf, pos + 2, 2), _u(buf, pos + 8, 4)
        blocks = [(pos + 16, pos + 16 + hsize)]
        while blocks and n_msgs > 0:
            p, end = blocks.pop(0)
            while p + 8 <= end and n_msgs > 0:
                mtype, msize, mflags = _u(buf, p, 2), _u(buf, p + 2, 2), buf[p + 4]
                data = bytes(buf[p + 8:p + 8 + msize])
                p += 8 + msize
                n_msgs -= 1
                if mtype == 0x10:
                    o, ln = _u(data, 0, self.O), _u(data, self.O, self.L)
                    blocks.append((o + self.base, o + self.base + ln))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out


class Group:
    def __init__(self, f, addr, name):
        self._f, self._addr, self.name = f, addr, name
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        f = self._f
        links = {}
        for mtype, _, d in f._messages(self._addr):
            if mtype == 0x11:  # symbol table: B-tree v1 + local heap
                btree, heap = _u(d, 0, f.O), _u(d, f.O, f.O)
                links.update(self._symbol_table(btree, heap))
            elif mtype == 0x06:  # link message
                name, addr = self._link(d)
                if addr is not None:
                    links[name] = addr
            elif mtype == 0x02:  # link info: dense storage when it names a fractal heap
                fl = d[1]
                p = 2 + (8 if fl & 1 else 0)
                if not f._undef(_u(d, p, f.O)):
                    raise NotImplementedError("dense link storage (fractal heap) is not supported by h5lite")
        self._links = links
        return links

    def _link(self, d):
        f = self._f
        flags = d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        n = 1 << (flags & 3)
        ln = _u(d, p, n)
        p += n
        name = d[p:p + ln].decode("utf-8")
        p += ln
        return name, (_u(d, p, f.O) if ltype == 0 else None)  # soft / external links are not followed

    def _symbol_table(self, btree, heap):
        f, buf = self._f, self._f._buf
        hp = heap + f.base
        if buf[hp:hp + 4] != b"HEAP":
            raise H5FormatError("bad local heap")
        seg = _u(buf, hp + 8 + 2 * f.L, f.O) + f.base
        out = {}

        def name_at(off):
            end = buf.find(b"\0", seg + off)
            return bytes(buf[seg + off:end]).decode("utf-8")

        def walk(addr):
            p = addr + f.base
            sig = bytes(buf[p:p + 4])
            if sig == b"TREE":
                level, used = buf[p + 5], _u(buf, p + 6, 2)
                q = p + 8 + 2 * f.O
                for k in range(used):
                    child = _u(buf, q + f.L + k * (f.L + f.O), f.O)
                    walk(child)
                _ = level
            elif sig == b"SNOD":
                n = _u(buf, p + 6, 2)
                q = p + 8
                esz = 2 * f.O + 24
                for k in range(n):
                    e = q + k * esz
                    out[name_at(_u(buf, e, f.O))] = _u(buf, e + f.O, f.O)
            else:
                raise H5FormatError("bad group B-tree node")

        walk(btree)
        return out

    def keys(self):
        return sorted(self._load())

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        node = self
        for part in [p for p in name.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(name)
            links = node._load()
            if part not in links:
                raise KeyError(f"{name!r} not in {node.name!r} (has {sorted(links)})")
            node = node._open(part, links[part])
        return node

    def _open(self, part, addr):
        f = self._f
        types = {m[0] for m in f._messages(addr)}
        path = (self.name.rstrip("/") + "/" + part)
        if 0x08 in types:
            return Dataset(f, addr, path)
        return Group(f, addr, path)


def _parse_datatype(d):
    cls, ver = d[0] & 0x0F, d[0] >> 4
    bits = d[1] | (d[2] << 8) | (d[3] << 16)
    size = _u(d, 4, 4)
    if cls == 0:  # fixed point
        return np.dtype((">" if bits & 1 else "<") + ("i" if bits & 8 else "u") + str(size))
    if cls == 1:  # floating point: IEEE sizes only
        if size not in (2, 4, 8):
            raise NotImplementedError(f"{size}-byte floating point")
        return np.dtype((">" if bits & 1 else "<") + "f" + str(size))
    if cls == 8:  # enumeration (h5py stores numpy bool as an enum of int8): read as the base type
        return _parse_datatype(d[8:])
    raise NotImplementedError(f"HDF5 datatype class {cls} (version {ver}) is not supported by h5lite")


class Dataset:
    def __init__(self, f, addr, name):
        self._f, self.name = f, name
        self.shape = self.dtype = None
        self._layout = None
        self._filters = []
        self.chunks = None
        for mtype, _, d in f._messages(addr):
            if mtype == 0x01:
                self.shape = self._dataspace(d)
            elif mtype == 0x03:
                self.dtype = _parse_datatype(d)
            elif mtype == 0x08:
                self._layout = self._parse_layout(d)
            elif mtype == 0x0B:
                self._filters = self._parse_filters(d)
        if self.shape is None or self.dtype is None or self._layout is None:
            raise H5FormatError(f"{name}: dataset header lacks dataspace, datatype or layout")
        self._chunk_index = None
        if self._layout[0] == "chunked":
            self.chunks = tuple(self._chunk_dims[:len(self.shape)])  # the key dimensionality counts the element size as one more axis

    # -- header messages --
    def _dataspace(self, d):
        f = self._f
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                raise NotImplementedError("null dataspace")
            p = 4
        else:
            raise H5FormatError(f"dataspace version {ver}")
        _ = flags
        return tuple(_u(d, p + k * f.L, f.L) for k in range(rank))

    def _parse_layout(self, d):
        f = self._f
        ver = d[0]
        if ver == 3:
            cls = d[1]
            if cls == 0:
                n = _u(d, 2, 2)
                return ("compact", d[4:4 + n])
            if cls == 1:
                return ("contiguous", _u(d, 2, f.O), _u(d, 2 + f.O, f.L))
            if cls == 2:
                nd = d[2]
                bt = _u(d, 3, f.O)
                self._chunk_dims = tuple(_u(d, 3 + f.O + 4 * k, 4) for k in range(nd))
                return ("chunked", bt, nd)
            raise H5FormatError(f"layout class {cls}")
        if ver in (1, 2):
            nd, cls = d[1], d[2]
            p = 8
            addr = None
            if cls != 0:
                addr = _u(d, p, f.O)
                p += f.O
            dims = tuple(_u(d, p + 4 * k, 4) for k in range(nd))
            p += 4 * nd
            if cls == 2:
                self._chunk_dims = dims
                return ("chunked", addr, nd)
            if cls == 1:
                return ("contiguous", addr, None)
            n = _u(d, p, 4)
            return ("compact", d[p + 4:p + 4 + n])
        raise NotImplementedError(f"data layout message version {ver} (written with libver='latest'?) is not supported by h5lite")

    def _parse_filters(self, d):
        ver, n = d[0], d[1]
        out = []
        p = 8 if ver == 1 else 2
        for _ in range(n):
            fid = _u(d, p, 2)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = _u(d, p, 2)
                p += 2
            p += 2  # flags
            ncd = _u(d, p, 2)
            p += 2
            if ver == 1:
                p += (nlen + 7) // 8 * 8
            else:
                p += nlen
            cd = [_u(d, p + 4 * k, 4) for k in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    # -- data --
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.shape[0]

    def _contiguous_view(self):
        f = self._f
        kind = self._layout[0]
        if kind == "compact":
            return np.frombuffer(self._layout[1], dtype=self.dtype, count=self.size).reshape(self.shape)
        addr = self._layout[1]
        if f._undef(addr):  # never written: fill value (0)
            return np.zeros(self.shape, self.dtype)
        return np.frombuffer(f._buf, dtype=self.dtype, count=self.size, offset=addr + f.base).reshape(self.shape)

    def _chunks(self):
        """-> list of (offsets tuple, file position, stored size, filter mask) of every stored chunk"""
        if self._chunk_index is not None:
            return self._chunk_index
        f, buf = self._f, self._f._buf
        _, bt, nd = self._layout  # nd = rank + 1
        out = []
        ksz = 8 + 8 * nd

        def walk(addr):
            p = addr + f.base
            if bytes(buf[p:p + 4]) != b"TREE" or buf[p + 4] != 1:
                raise H5FormatError("bad chunk B-tree node")
            level, used = buf[p + 5], _u(buf, p + 6, 2)
            q = p + 8 + 2 * f.O
            for k in range(used):
                e = q + k * (ksz + f.O)
                child = _u(buf, e + ksz, f.O)
                if level:
                    walk(child)
                else:
                    offs = tuple(_u(buf, e + 8 + 8 * j, 8) for j in range(nd - 1))
                    out.append((offs, child + f.base, _u(buf, e, 4), _u(buf, e + 4, 4)))

        if not f._undef(bt):
            walk(bt)
        self._chunk_index = out
        return out

    def _decode_chunk(self, pos, nbytes, mask):
        raw = bytes(self._f._buf[pos:pos + nbytes])
        esz = self.dtype.itemsize
        for k in range(len(self._filters) - 1, -1, -1):
            if mask & (1 << k):
                continue
            fid, _ = self._filters[k]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // esz
                raw = np.frombuffer(raw, np.uint8, n * esz).reshape(esz, n).T.tobytes() + raw[n * esz:]
            elif fid == 3:
                raw = raw[:-4]  # fletcher32 checksum: not verified
            else:
                raise NotImplementedError(f"HDF5 filter {fid} is not supported by h5lite")
        return np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(self.chunks, dtype=np.int64))).reshape(self.chunks)

    def _read_box(self, lo, hi):
        """the hyperslab [lo, hi) of a chunked dataset"""
        out = np.zeros(tuple(h - l for l, h in zip(lo, hi)), self.dtype)
        cs = self.chunks
        for offs, pos, nbytes, mask in self._chunks():
            a = [max(o, l) for o, l in zip(offs, lo)]
            b = [min(o + c, h, s) for o, c, h, s in zip(offs, cs, hi, self.shape)]
            if any(x >= y for x, y in zip(a, b)):
                continue
            chunk = self._decode_chunk(pos, nbytes, mask)
            src = tuple(slice(x - o, y - o) for x, y, o in zip(a, b, offs))
            dst = tuple(slice(x - l, y - l) for x, y, l in zip(a, b, lo))
            out[dst] = chunk[src]
        return out

    def __getitem__(self, key):
        if self._layout[0] != "chunked":
            return np.array(self._contiguous_view()[key])  # a copy, like h5py returns
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = [k is Ellipsis for k in key].index(True)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        lo, hi, rest = [], [], []
        for k, n in zip(key, self.shape):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (n if k < 0 else 0)
                if not 0 <= k < n:
                    raise IndexError(f"index {k} out of range for axis of size {n}")
                lo.append(k); hi.append(k + 1); rest.append(0)
            elif isinstance(k, slice):
                a, b, st = k.indices(n)
                if st > 0 and b > a:
                    lo.append(a); hi.append(b); rest.append(slice(None, None, st))
                else:  # empty or reversed: read the whole axis and let numpy index it
                    lo.append(0); hi.append(n); rest.append(k)
            else:
                lo.append(0); hi.append(n); rest.append(k)
        return np.array(self._read_box(lo, hi)[tuple(rest)])

    def __array__(self, dtype=None, copy=None):
        a = self[()] if self.ndim == 0 else self[:]
        return a.astype(dtype) if dtype is not None else a


# ------------------------------------------------------------------------------------------------------------------
# writer (fixtures and conversions; superblock v0, one root group, numeric datasets)
# ------------------------------------------------------------------------------------------------------------------
def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt == np.bool_:
        dt = np.dtype("u1")
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits = be | (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # IEEE: sign position, exponent location / size, mantissa location / size, exponent bias
        sign, eloc, esz, msz, bias = (31, 23, 8, 23, 127) if dt.itemsize == 4 else (63, 52, 11, 52, 1023)
        bits = be | 0x20  # mantissa normalisation: msb implied
        return struct.pack("<BBBBI", 0x11, bits, sign, 0, dt.itemsize) + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, 0, msz, bias)
    raise NotImplementedError(f"h5lite.write_h5: dtype {dt}")


def _message(mtype, data, flags=0):
    pad = (-len(data)) % 8
    return struct.pack("<HHB3x", mtype, len(data) + pad, flags) + data + b"\0" * pad


def write_h5(path, arrays, chunks=None, compression=None, userblock=0):
    """arrays: {name: ndarray}.  chunks: {name: chunk shape} -> chunked layout for those (B-tree v1), with
    compression='gzip' also shuffle + deflate; everything else contiguous.  userblock: 0 or a power of two >= 512."""
    chunks = chunks or {}
    O = L = 8
    out = bytearray()

    def align(n=8):
        out.extend(b"\0" * ((-len(out)) % n))

    def put(b):
        align()
        pos = len(out)
        out.extend(b)
        return pos

    names = sorted(arrays)
    SB = 24 + 4 * O + (2 * O + 24)  # superblock v0 with the root symbol table entry
    out.extend(b"\0" * SB)

    # datasets
    headers = {}
    for name in names:
        a = np.asarray(arrays[name])
        a = np.ascontiguousarray(a) if a.ndim else a
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        msgs = _message(0x01, struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", n) for n in a.shape))
        msgs += _message(0x03, _dtype_message(a.dtype), flags=1)
        msgs += _message(0x05, struct.pack("<BBBB", 2, 2, 2, 0))  # fill value v2: allocate late, write at allocation, undefined
        if name in chunks:
            cs = tuple(int(c) for c in chunks[name])
            assert len(cs) == a.ndim and all(c > 0 for c in cs)
            entries = []
            grid = [range(0, max(n, 1), c) for n, c in zip(a.shape, cs)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = tuple(g[i] for g, i in zip(grid, offs))
                block = np.zeros(cs, a.dtype)
                sl = tuple(slice(x, min(x + c, n)) for x, c, n in zip(o, cs, a.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                raw = block.tobytes()
                if compression == "gzip":
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, a.dtype.itemsize).T.tobytes()
                    raw = zlib.compress(raw, 4)
                entries.append((o, put(raw), len(raw)))
            nd = a.ndim + 1

            def node(level, items):  # items: (first key offsets, size, child address); the final key closes the node
                body = struct.pack("<4sBBH", b"TREE", 1, level, len(items)) + struct.pack("<QQ", UNDEF, UNDEF)
                for o, size, child in items:
                    body += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", x) for x in o) + struct.pack("<Q", 0)
                    body += struct.pack("<Q", child)
                last = tuple(a.shape) if a.ndim else ()
                body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", x) for x in last) + struct.pack("<Q", 0)
                return put(body)

            K2 = 64  # 2K entries per node with the default K = 32
            leaves = []
            for k in range(0, len(entries), K2):
                part = entries[k:k + K2]
                leaves.append((part[0][0], 0, node(0, [(o, size, pos) for o, pos, size in part])))
            level = 1
            while len(leaves) > 1:
                leaves = [(leaves[k][0], 0, node(level, leaves[k:k + K2])) for k in range(0, len(leaves), K2)]
                level += 1
            btree = leaves[0][2] if leaves else UNDEF
            if compression == "gzip":
                flt = struct.pack("<BB6x", 1, 2)
                flt += struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\0" + struct.pack("<II", a.dtype.itemsize, 0)
                flt += struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", 4, 0)
                msgs += _message(0x0B, flt)
            msgs += _message(0x08, struct.pack("<BBB", 3, 2, nd) + struct.pack("<Q", btree) + b"".join(struct.pack("<I", c) for c in cs)
                             + struct.pack("<I", a.dtype.itemsize))
            n_msgs = 5 if compression == "gzip" else 4
        else:
            pos = put(a.tobytes()) if a.size else UNDEF
            msgs += _message(0x08, struct.pack("<BB", 3, 1) + struct.pack("<QQ", pos, a.nbytes))
            n_msgs = 4
        headers[name] = put(struct.pack("<BBHII4x", 1, 0, n_msgs, 1, len(msgs)) + msgs)

    # root group: local heap with the names, one SNOD, one B-tree node, object header with the symbol table message
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for name in names:
        name_off[name] = len(heap_data)
        heap_data.extend(name.encode() + b"\0")
        heap_data.extend(b"\0" * ((-len(heap_data)) % 8))
    heap_seg = put(bytes(heap_data))
    heap = put(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap_data), UNDEF, heap_seg))
    leaf_k = max(4, (len(names) + 1) // 2)
    snod = struct.pack("<4sBBH", b"SNOD", 1, 0, len(names))
    for name in names:
        snod += struct.pack("<QQII16x", name_off[name], headers[name], 0, 0)
    snod += b"\0" * ((2 * leaf_k - len(names)) * 40)
    snod_pos = put(snod)
    last = name_off[names[-1]] if names else 0
    bt = put(struct.pack("<4sBBH", b"TREE", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_pos, last))
    root_msgs = _message(0x11, struct.pack("<QQ", bt, heap))
    root = put(struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs)
    align()
    eof = len(out)
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, O, L, 0) + struct.pack("<HHI", leaf_k, 16, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, eof, UNDEF)  # base address = the superblock's own offset; addresses are relative to it
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, heap)
    assert len(sb) == SB
    out[:SB] = sb
    with open(path, "wb") as fh:
        if userblock:
            assert userblock >= 512 and userblock & (userblock - 1) == 0
            fh.write(b"\0" * userblock)
        fh.write(bytes(out))
