"""Minimal pure-Python HDF5 reader / writer for the files the reference reads and writes (no libhdf5 / h5py in this image).

Scope = what HDF5.jl + MatsubaraFunctions.jl's `save!` / `load_mesh_function` produce with the library defaults (checked against
the 34 files under the reference's data/ directory): superblock version 0, "old style" groups (symbol table: v1 B-tree + local
heap + SNOD nodes), version-1 object headers with continuation blocks, contiguous or compact dataset layout, no filters,
datatypes: IEEE floats, fixed-point integers, fixed-length strings, compounds of those (complex numbers are the compound
{r: f64, i: f64}), attributes (versions 1-3).  Anything else raises H5Error rather than guessing.

The writer emits the same subset (one symbol-table group per Python dict level, contiguous datasets, version-1 attributes), so a
file written here is read back by this reader and, being plain HDF5 1.8 "earliest" format, by libhdf5.

Reference call sites: src/ParquetSolver.jl:309-330 (`save!` of a solver), src/nonlocal/ParquetSolver.jl:332-346 (`load_solver!`),
src/channel.jl:377-402, src/refvertex.jl:219-250, src/vertex.jl (save!/load of vertices), src/utility/load_triqs.jl:298-308.
Format: "HDF5 File Format Specification Version 2.0" (superblock 0, sections III.A-III.D, IV.A).
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


# ================================================================================================ reader
class Dataset:
    def __init__(self, f, name, dtype, shape, layout, attrs):
        self._f, self.name, self.dtype, self.shape, self._layout, self.attrs = f, name, dtype, shape, layout, attrs

    def read(self):
        """numpy array with the HDF5 (row-major) dimension order; Julia arrays appear with their dimensions reversed"""
        kind, a, b = self._layout
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = n * self.dtype.itemsize
        if kind == "contiguous":
            if a == UNDEF:
                raw = b"\x00" * nbytes
            else:
                raw = self._f._buf[a:a + nbytes]
        elif kind == "compact":
            raw = a[:nbytes]
        else:
            raise H5Error(f"dataset {self.name}: layout {kind} not supported")
        if len(raw) != nbytes:
            raise H5Error(f"dataset {self.name}: short read")
        arr = np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape)
        return _complexify(arr)

    def __repr__(self):
        return f"<Dataset {self.name} shape={self.shape} dtype={self.dtype}>"


def _complexify(arr):
    """{r, i} / {re, im} float compounds -> complex; everything else unchanged"""
    names = arr.dtype.names
    if names and len(names) == 2 and set(n.lower() for n in names) in ({"r", "i"}, {"re", "im"}, {"real", "imag"}):
        re = [n for n in names if n.lower() in ("r", "re", "real")][0]
        im = [n for n in names if n != re][0]
        return (arr[re] + 1j * arr[im]).astype(np.complex128)
    return arr


class Group:
    def __init__(self, f, name, children, attrs):
        self._f, self.name, self._children, self.attrs = f, name, children, attrs

    def keys(self):
        return list(self._children)

    def __contains__(self, k):
        return k.split("/")[0] in self._children if "/" in k else k in self._children

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._children:
                raise KeyError(path)
            node = node._f._object(node._children[part], (node.name.rstrip("/") + "/" + part))
        return node

    def __repr__(self):
        return f"<Group {self.name} {self.keys()}>"


class File(Group):
    def __init__(self, path):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        b = self._buf
        if b[:8] != SIG:
            raise H5Error("not an HDF5 file")
        if b[8] != 0:
            raise H5Error(f"superblock version {b[8]} not supported (the reference's files are version 0)")
        self._so, self._sl = b[13], b[14]
        if (self._so, self._sl) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        self._base = struct.unpack_from("<Q", b, 24)[0]
        root_entry = 24 + 4 * 8
        _, ohdr, _, _ = struct.unpack_from("<QQII", b, root_entry)
        self._cache = {}
        root = self._object(ohdr, "/")
        super().__init__(self, "/", root._children, root.attrs)

    # ---- object headers
    def _messages(self, addr):
        b = self._buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} not supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off + self._base, ln))
                out.append((mtype, mflags, data))
        return out

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr + self._base if addr < len(self._buf) else addr)
        attrs, stab, dspace, dtype, layout, links = {}, None, None, None, None, None
        for mtype, mflags, data in msgs:
            if mflags & 2:
                raise H5Error(f"{name}: shared header messages are not supported")
            if mtype == 0x11:
                stab = struct.unpack_from("<QQ", data, 0)
            elif mtype == 0x01:
                dspace = _parse_dataspace(data)
            elif mtype == 0x03:
                dtype, _ = _parse_datatype(data, 0)
            elif mtype == 0x08:
                layout = _parse_layout(data)
            elif mtype == 0x0B:
                raise H5Error(f"{name}: filtered (compressed) datasets are not supported")
            elif mtype == 0x0C:
                k, v = _parse_attribute(data)
                attrs[k] = v
            elif mtype == 0x02:     # link info: "compact" groups keep their links as link messages in the header
                flags = data[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", data, p)[0]
                if fheap != UNDEF:
                    raise H5Error(f"{name}: densely stored groups (fractal heap) are not supported")
                links = {} if links is None else links
            elif mtype == 0x06:
                k, a = _parse_link(data)
                links = {} if links is None else links
                links[k] = a
        if stab is not None and links:
            raise H5Error(f"{name}: group with both a symbol table and link messages")
        if links is not None and stab is None and dtype is None:
            obj = Group(self, name, links, attrs)
        elif stab is not None:
            obj = Group(self, name, self._group_entries(*stab), attrs)
        elif dtype is not None and layout is not None:
            obj = Dataset(self, name, dtype, dspace if dspace is not None else (), layout, attrs)
        else:
            raise H5Error(f"{name}: neither a group nor a dataset")
        self._cache[addr] = obj
        return obj

    # ---- old-style groups: v1 B-tree of symbol-table nodes + local heap of names
    def _group_entries(self, btree, heap):
        b = self._buf
        heap += self._base
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        data_addr = struct.unpack_from("<Q", b, heap + 24)[0] + self._base
        out = {}

        def name_at(off):
            s = data_addr + off
            return b[s:b.index(b"\x00", s)].decode("utf-8")

        def walk(node):
            node += self._base
            if b[node:node + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, node + 6)[0]
                for i in range(n):
                    e = node + 8 + 40 * i
                    noff, ohdr = struct.unpack_from("<QQ", b, e)
                    out[name_at(noff)] = ohdr
                return
            if b[node:node + 4] != b"TREE":
                raise H5Error("B-tree signature missing")
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            if ntype != 0:
                raise H5Error("not a group B-tree")
            pos = node + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8)[0]      # key_i (8) child_i (8) ... key_n
                walk(child)
                pos += 16
        walk(btree)
        return out


def _parse_link(d):
    ver, flags = d[0], d[1]
    if ver != 1:
        raise H5Error(f"link message version {ver}")
    p = 2
    ltype = 0
    if flags & 0x08:
        ltype = d[p]; p += 1
    if flags & 0x04:
        p += 8                      # creation order
    if flags & 0x10:
        p += 1                      # character set
    nb = 1 << (flags & 3)
    n = int.from_bytes(d[p:p + nb], "little"); p += nb
    name = d[p:p + n].decode("utf-8"); p += n
    if ltype != 0:
        raise H5Error(f"link {name}: only hard links are supported")
    return name, struct.unpack_from("<Q", d, p)[0]


def _parse_dataspace(d):
    ver, rank, flags = d[0], d[1], d[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if d[3] == 2:       # null dataspace
            return (0,)
    else:
        raise H5Error(f"dataspace version {ver}")
    return tuple(struct.unpack_from("<Q", d, off + 8 * i)[0] for i in range(rank))


def _parse_datatype(d, pos):
    """(numpy dtype, bytes consumed)"""
    cv = d[pos]
    cls, ver = cv & 0x0F, cv >> 4
    bits = d[pos + 1] | (d[pos + 2] << 8) | (d[pos + 3] << 16)
    size = struct.unpack_from("<I", d, pos + 4)[0]
    p = pos + 8
    if cls == 0:      # fixed point: bit offset (2), precision (2)
        signed = bool(bits & 0x08)
        if bits & 1:
            raise H5Error("big-endian integers not supported")
        return np.dtype(("<i" if signed else "<u") + str(size)), p + 4 - pos
    if cls == 1:      # floating point: 12 bytes of properties
        if bits & 1:
            raise H5Error("big-endian floats not supported")
        return np.dtype("<f" + str(size)), p + 12 - pos
    if cls == 3:      # fixed-length string
        return np.dtype("S" + str(size)), p - pos
    if cls == 6:      # compound
        nmem = bits & 0xFFFF
        fields = []
        for _ in range(nmem):
            e = d.index(b"\x00", p)
            name = d[p:e].decode("utf-8")
            if ver < 3:
                p += ((e - p) // 8 + 1) * 8                      # name padded to a multiple of 8 (including the terminator)
                off = struct.unpack_from("<I", d, p)[0]
                p += 4
                if ver == 1:
                    p += 1 + 3 + 4 + 4 + 16                       # dimensionality, reserved, permutation, reserved, 4 dim sizes
            else:
                p = e + 1
                nb = 1 if size < 256 else (2 if size < 65536 else (3 if size < 16777216 else 4))
                off = int.from_bytes(d[p:p + nb], "little")
                p += nb
            mt, used = _parse_datatype(d, p)
            p += used
            fields.append((name, mt, off))
        return np.dtype({"names": [f[0] for f in fields], "formats": [f[1] for f in fields], "offsets": [f[2] for f in fields],
                         "itemsize": size}), p - pos
    if cls == 9:
        raise H5Error("variable-length datatypes are not supported")
    raise H5Error(f"datatype class {cls} not supported")


def _parse_layout(d):
    ver = d[0]
    if ver == 3:
        cls = d[1]
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", d, 2)
            return ("contiguous", addr, size)
        if cls == 0:
            n = struct.unpack_from("<H", d, 2)[0]
            return ("compact", bytes(d[4:4 + n]), n)
        return ("chunked", None, None)
    if ver in (1, 2):
        rank, cls = d[1], d[2]
        if cls == 1:
            addr = struct.unpack_from("<Q", d, 8)[0]
            return ("contiguous", addr, None)
        return ("chunked" if cls == 2 else "compact-v1", None, None)
    raise H5Error(f"data layout version {ver}")


def _pad8(n):
    return (n + 7) & ~7


def _parse_attribute(d):
    ver = d[0]
    if ver == 1:
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8
        name = d[p:p + nsz].split(b"\x00")[0].decode("utf-8"); p += _pad8(nsz)
        dt, _ = _parse_datatype(d, p); p += _pad8(tsz)
        shape = _parse_dataspace(d[p:p + ssz]); p += _pad8(ssz)
    elif ver in (2, 3):
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8 + (1 if ver == 3 else 0)
        name = d[p:p + nsz].split(b"\x00")[0].decode("utf-8"); p += nsz
        dt, _ = _parse_datatype(d, p); p += tsz
        shape = _parse_dataspace(d[p:p + ssz]); p += ssz
    else:
        raise H5Error(f"attribute version {ver}")
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    arr = _complexify(np.frombuffer(d[p:p + n * dt.itemsize], dtype=dt, count=n).reshape(shape))
    if arr.dtype.kind == "S":
        vals = [x.split(b"\x00")[0].decode("utf-8") for x in arr.reshape(-1)]
        return name, (vals[0] if not shape else vals)
    return name, (arr.reshape(-1)[0].item() if not shape else arr)


# ================================================================================================ writer
class _Out:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def align(self, n=8):
        self.b += b"\x00" * ((-len(self.b)) % n)

    def write(self, data):
        pos = len(self.b)
        self.b += data
        return pos


def _dt_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        h = dt.itemsize // 2
        return _dt_message(np.dtype({"names": ["r", "i"], "formats": [f"<f{h}", f"<f{h}"], "offsets": [0, h], "itemsize": dt.itemsize}))
    if dt.kind == "f":
        if dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)      # little endian, mantissa normalisation "implied", sign bit location 63
        elif dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            raise H5Error("float size")
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], dt.itemsize) + props
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated, ASCII
    if dt.names:
        body = b""
        for name in dt.names:
            ft, off = dt.fields[name][0], dt.fields[name][1]
            nm = name.encode("utf-8") + b"\x00"
            nm += b"\x00" * ((-len(nm)) % 8)
            body += nm + struct.pack("<IB3xII4I", off, 0, 0, 0, 0, 0, 0, 0) + _dt_message(ft)
        n = len(dt.names)
        return struct.pack("<BBBBI", 0x16, n & 0xFF, (n >> 8) & 0xFF, 0, dt.itemsize) + body
    raise H5Error(f"cannot write dtype {dt}")


def _ds_message(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _msg(mtype, data, flags=0):
    data = bytes(data) + b"\x00" * ((-len(data)) % 8)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_message(name, value):
    if isinstance(value, str):
        raw = value.encode("utf-8") + b"\x00"
        dt, shape, data = np.dtype("S" + str(len(raw))), (), raw
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            w = max(len(s.encode("utf-8")) for s in arr.reshape(-1)) + 1
            arr = np.array([s.encode("utf-8") for s in arr.reshape(-1)], dtype="S" + str(w)).reshape(arr.shape)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if arr.dtype.kind == "i" and arr.dtype.itemsize != 8 and not arr.shape:
            arr = arr.astype(np.int64)
        dt, shape, data = arr.dtype, arr.shape, _raw_bytes(arr)
    nm = name.encode("utf-8") + b"\x00"
    t, s = _dt_message(dt), _ds_message(shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(t), len(s))
    body += nm + b"\x00" * ((-len(nm)) % 8) + t + b"\x00" * ((-len(t)) % 8) + s + b"\x00" * ((-len(s)) % 8) + data
    return _msg(0x0C, body)


def _raw_bytes(arr):
    arr = np.ascontiguousarray(arr)
    if arr.dtype.kind == "c":
        return arr.view(f"<f{arr.dtype.itemsize // 2}").tobytes()
    return arr.tobytes()


def _object_header(out, msgs):
    body = b"".join(msgs)
    out.align(8)
    pos = out.write(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)
    return pos


class DatasetSpec:
    """leaf of the tree handed to write_file: an array plus attributes"""

    def __init__(self, data, attrs=None):
        self.data, self.attrs = np.asarray(data), dict(attrs or {})


class GroupSpec(dict):
    """dict of children (GroupSpec / DatasetSpec / array) plus attributes"""

    def __init__(self, children=None, attrs=None):
        super().__init__(children or {})
        self.attrs = dict(attrs or {})


def _write_dataset(out, spec):
    arr = spec.data
    out.align(8)
    addr = out.write(_raw_bytes(arr)) if arr.size else UNDEF
    nbytes = arr.size * arr.dtype.itemsize
    msgs = [_msg(0x01, _ds_message(arr.shape)), _msg(0x03, _dt_message(arr.dtype), flags=1),
            _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                      # fill value: version 2, allocate late, write never, undefined
            _msg(0x08, struct.pack("<BBQQ", 3, 1, addr, nbytes))]
    msgs += [_attr_message(k, v) for k, v in spec.attrs.items()]
    return _object_header(out, msgs)


def _write_group(out, spec):
    attrs = getattr(spec, "attrs", {})
    entries = []
    for name in sorted(spec):          # symbol table nodes hold their entries sorted by name
        child = spec[name]
        if isinstance(child, dict):
            entries.append((name, _write_group(out, child if isinstance(child, GroupSpec) else GroupSpec(child))[0]))
        else:
            entries.append((name, _write_dataset(out, child if isinstance(child, DatasetSpec) else DatasetSpec(child))))
    # local heap: offset 0 holds the empty string
    heap_data = bytearray(b"\x00" * 8)
    offs = []
    for name, _ in entries:
        offs.append(len(heap_data))
        nm = name.encode("utf-8") + b"\x00"
        heap_data += nm + b"\x00" * ((-len(nm)) % 8)
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)       # one free block: next = 1 (none), size 16
    out.align(8)
    data_addr = out.write(bytes(heap_data))
    out.align(8)
    heap_addr = out.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
    # symbol table nodes of at most 2 * leaf_k = 8 entries, one B-tree level above them
    LEAF = 8
    snods = []
    for i in range(0, max(len(entries), 1), LEAF):
        chunk = list(zip(offs[i:i + LEAF], entries[i:i + LEAF]))
        body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
        for noff, (_, ohdr) in chunk:
            body += struct.pack("<QQII16x", noff, ohdr, 0, 0)
        body += b"\x00" * (40 * (LEAF - len(chunk)))
        out.align(8)
        snods.append((out.write(body), chunk[-1][0] if chunk else 0))
    if len(snods) > 32:
        raise H5Error("group too large for the single-level B-tree this writer emits")
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for addr, last_off in snods:
        node += struct.pack("<QQ", addr, last_off)
    node += b"\x00" * (16 * (32 - len(snods)))
    out.align(8)
    btree_addr = out.write(node)
    msgs = [_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + [_attr_message(k, v) for k, v in attrs.items()]
    return _object_header(out, msgs), btree_addr, heap_addr


def write_file(path, tree):
    """tree: GroupSpec / dict of (GroupSpec | dict | DatasetSpec | array)"""
    out = _Out()
    out.write(b"\x00" * 96)                     # superblock, filled in last
    root = tree if isinstance(tree, GroupSpec) else GroupSpec(tree)
    ohdr, btree, heap = _write_group(out, root)
    out.align(8)
    eof = out.tell()
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96, len(sb)
    out.b[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(out.b))
