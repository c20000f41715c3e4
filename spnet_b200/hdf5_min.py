"""A minimal HDF5 reader / writer for Keras weight files (`weights.hdf5`, `spnet.model`, `full_model.h5` — the
reference's checkpoint formats: spnet/models.py:475-485 load_weights, spnet/callbacks.py:35-41 save_weights / save,
train_spnet.py:145-150). h5py is not available in this image, so the subset of the HDF5 file format those files use is
implemented here from the format specification:

  * superblock version 0 (also accepts 2 / 3), 8-byte offsets and lengths;
  * "old style" groups: symbol-table message -> version-1 B-tree of symbol-table nodes ("SNOD") + local heap;
    "new style" compact groups (link messages) are read as well;
  * version-1 object headers with continuation blocks (version-2 "OHDR" headers are read as well);
  * datasets: contiguous or compact layout (and chunked WITHOUT filters), little-endian IEEE floats / integers;
  * attributes (message versions 1-3): fixed-length strings, variable-length strings through the global heap,
    numeric scalars / arrays.

That is what h5py writes with its default `libver='earliest'` for Keras 2.x `save_weights` / `save` (no compression,
no chunking). The writer emits the same structures (superblock 0, symbol-table groups, contiguous datasets, fixed-length
string attributes), so files written here follow the layout h5py produces for Keras; it has been validated by reading
back with this reader only - no h5py in the image - and says so in DESIGN.md.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


# =====================================================================================================================
# reader
# =====================================================================================================================
class H5Error(IOError):
    pass


class Reader:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        # the superblock sits at offset 0 or, behind a user block, at 512, 1024, 2048, ...
        s0 = 0
        while self.b[s0:s0 + 8] != SIG:
            s0 = 512 if s0 == 0 else 2 * s0
            if s0 + 8 > len(self.b):
                raise H5Error("%s is not an HDF5 file" % path)
        ver = self.b[s0 + 8]
        if ver in (0, 1):
            so, sl = self.b[s0 + 13], self.b[s0 + 14]
            if (so, sl) != (8, 8):
                raise H5Error("only 8-byte offsets / lengths are supported")
            p = s0 + 24 + (4 if ver == 1 else 0)
            self.base = self.u64(p)
            # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
            p += 32
            self.root = self.u64(p + 8)
        elif ver in (2, 3):
            if (self.b[s0 + 9], self.b[s0 + 10]) != (8, 8):
                raise H5Error("only 8-byte offsets / lengths are supported")
            self.base = self.u64(s0 + 12)
            self.root = self.u64(s0 + 12 + 24)
        else:
            raise H5Error("unsupported superblock version %d" % ver)
        self._gheap = {}

    # ---- primitives
    def u8(self, p):
        return self.b[p]

    def u16(self, p):
        return struct.unpack_from("<H", self.b, p)[0]

    def u32(self, p):
        return struct.unpack_from("<I", self.b, p)[0]

    def u64(self, p):
        return struct.unpack_from("<Q", self.b, p)[0]

    # ---- object headers -> list of (type, flags, payload offset, size)
    def messages(self, addr):
        addr += self.base
        out = []
        if self.b[addr:addr + 4] == b"OHDR":
            self._messages_v2(addr, out)
            return out
        if self.b[addr] != 1:
            raise H5Error("unsupported object header version %d at %d" % (self.b[addr], addr))
        nmsg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u16(p), self.u16(p + 2), self.b[p + 4]
                body = p + 8
                if mtype == 0x0010:  # continuation
                    blocks.append((self.u64(body) + self.base, self.u64(body + 8)))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    def _messages_v2(self, addr, out):
        flags = self.b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szf = 1 << (flags & 3)
        chunk0 = int.from_bytes(self.b[p:p + szf], "little")
        p += szf
        blocks = [(p, chunk0)]
        track = bool(flags & 0x04)
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = self.b[p], self.u16(p + 1), self.b[p + 3]
                body = p + 4 + (2 if track else 0)
                if body + msize > end:
                    break
                if mtype == 0x10:
                    ca, cl = self.u64(body) + self.base, self.u64(body + 8)
                    blocks.append((ca + 4, cl - 8))  # "OCHK" signature in front, checksum behind
                elif mtype != 0:
                    out.append((mtype, mflags, body, msize))
                p = body + msize

    # ---- groups
    def links(self, addr):
        """{name: object header address} of the group at addr."""
        out = {}
        for mtype, _, body, size in self.messages(addr):
            if mtype == 0x0011:  # symbol table: B-tree + local heap
                self._walk_btree(self.u64(body), self._heap_data(self.u64(body + 8)), out)
            elif mtype == 0x0006:  # link message (new-style compact group)
                name, target = self._link_message(body)
                if target is not None:
                    out[name] = target
        return out

    def _heap_data(self, addr):
        addr += self.base
        if self.b[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap at %d" % addr)
        return self.u64(addr + 24) + self.base

    def _cstr(self, p):
        e = self.b.index(b"\x00", p)
        return self.b[p:e].decode("utf8")

    def _walk_btree(self, addr, heap, out):
        addr += self.base
        if self.b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree node at %d" % addr)
        level, used = self.b[addr + 5], self.u16(addr + 6)
        p = addr + 24
        for i in range(used):
            child = self.u64(p + 8)       # key i (8 bytes) precedes child i
            p += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                self._snod(child, heap, out)

    def _snod(self, addr, heap, out):
        addr += self.base
        if self.b[addr:addr + 4] != b"SNOD":
            raise H5Error("bad symbol-table node at %d" % addr)
        n = self.u16(addr + 6)
        p = addr + 8
        for _ in range(n):
            out[self._cstr(heap + self.u64(p))] = self.u64(p + 8)
            p += 40

    def _link_message(self, p):
        ver, flags = self.b[p], self.b[p + 1]
        p += 2
        ltype = 0
        if flags & 0x08:
            ltype = self.b[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        ls = 1 << (flags & 3)
        n = int.from_bytes(self.b[p:p + ls], "little")
        p += ls
        name = self.b[p:p + n].decode("utf8")
        p += n
        return name, (self.u64(p) if ltype == 0 else None)

    # ---- datatypes / dataspaces
    def _dtype(self, p):
        """-> (numpy dtype or ('S', n) or ('vlen_str',), size in bytes)."""
        cv = self.b[p]
        cls, bits0 = cv & 0x0F, self.b[p + 1]
        size = self.u32(p + 4)
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(("<" if not bits0 & 1 else ">") + ("i" if signed else "u") + str(size)), size
        if cls == 1:  # float
            return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size)), size
        if cls == 3:  # fixed-length string
            return ("S", size), size
        if cls == 9:  # variable length; type 1 = string
            if (bits0 & 0x0F) == 1:
                return ("vlen_str",), size
            raise H5Error("variable-length sequences are not supported")
        raise H5Error("unsupported datatype class %d" % cls)

    def _dspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
            if self.b[p + 3] == 2:  # null dataspace
                return None
        else:
            raise H5Error("unsupported dataspace version %d" % ver)
        return tuple(self.u64(q + 8 * i) for i in range(rank))

    def _global_heap_obj(self, addr, index):
        addr += self.base
        if addr not in self._gheap:
            if self.b[addr:addr + 4] != b"GCOL":
                raise H5Error("bad global heap at %d" % addr)
            size = self.u64(addr + 8)
            objs, p = {}, addr + 16
            while p + 16 <= addr + size:
                idx, osz = self.u16(p), self.u64(p + 8)
                if idx == 0:
                    break
                objs[idx] = self.b[p + 16:p + 16 + osz]
                p += 16 + ((osz + 7) // 8) * 8
            self._gheap[addr] = objs
        return self._gheap[addr][index]

    def _decode(self, dt, shape, raw_at, nbytes=None):
        n = 1
        for s in (shape or ()):
            n *= s
        if shape is None:
            return None
        if isinstance(dt, tuple) and dt[0] == "S":
            L = dt[1]
            vals = [self.b[raw_at + i * L:raw_at + (i + 1) * L].split(b"\x00")[0] for i in range(n)]
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        if isinstance(dt, tuple):  # variable-length strings: (length u32, global heap address u64, index u32)
            vals = []
            for i in range(n):
                q = raw_at + 16 * i
                ln, ga, gi = self.u32(q), self.u64(q + 4), self.u32(q + 12)
                vals.append(bytes(self._global_heap_obj(ga, gi)[:ln]) if ln else b"")
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=raw_at).reshape(shape)
        return a[()] if shape == () else a

    # ---- attributes and datasets
    def attrs(self, addr):
        out = {}
        for mtype, _, body, size in self.messages(addr):
            if mtype != 0x000C:
                continue
            ver = self.b[body]
            nsz, tsz, ssz = self.u16(body + 2), self.u16(body + 4), self.u16(body + 6)
            p = body + 8 + (1 if ver == 3 else 0)
            pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
            name = self.b[p:p + nsz].split(b"\x00")[0].decode("utf8")
            p += pad(nsz)
            dt, _ = self._dtype(p)
            p += pad(tsz)
            shape = self._dspace(p)
            p += pad(ssz)
            out[name] = self._decode(dt, shape, p)
        return out

    def is_dataset(self, addr):
        return any(m[0] == 0x0008 for m in self.messages(addr))

    def dataset(self, addr):
        dt = shape = layout = None
        for mtype, _, body, size in self.messages(addr):
            if mtype == 0x0003:
                dt, esz = self._dtype(body)
            elif mtype == 0x0001:
                shape = self._dspace(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                if self.u16(body + 2) if self.b[body] == 1 else self.b[body + 1]:
                    raise H5Error("filtered (compressed) datasets are not supported")
        if dt is None or layout is None:
            raise H5Error("object at %d is not a dataset" % addr)
        if self.b[layout] in (1, 2):  # libhdf5 1.6 and older: version, rank, class, 5 reserved, [address], dims
            cls = self.b[layout + 2]
            if cls == 1:
                a = self.u64(layout + 8)
                return np.zeros(shape, dt) if a == UNDEF else np.array(self._decode(dt, shape, a + self.base))
            raise H5Error("unsupported layout class %d in a version-%d layout message" % (cls, self.b[layout]))
        if self.b[layout] != 3:
            raise H5Error("unsupported data-layout message version %d" % self.b[layout])
        cls = self.b[layout + 1]
        if cls == 0:  # compact
            return np.array(self._decode(dt, shape, layout + 4))
        if cls == 1:  # contiguous
            a = self.u64(layout + 2)
            if a == UNDEF:
                return np.zeros(shape, dt)
            return np.array(self._decode(dt, shape, a + self.base))
        if cls == 2:  # chunked, no filters
            rank1 = self.b[layout + 2]
            bt = self.u64(layout + 3)
            cdims = [self.u32(layout + 11 + 4 * i) for i in range(rank1)]
            out = np.zeros(shape, dt)
            self._chunks(bt, rank1, cdims[:-1], out)
            return out
        raise H5Error("unsupported layout class %d" % cls)

    def _chunks(self, addr, rank1, cdims, out):
        addr += self.base
        if addr - self.base == UNDEF:
            return
        level, used = self.b[addr + 5], self.u16(addr + 6)
        ksz = 8 + 8 * rank1
        p = addr + 24
        for i in range(used):
            csize = self.u32(p)
            offs = [self.u64(p + 8 + 8 * d) for d in range(rank1 - 1)]
            child = self.u64(p + ksz)
            p += ksz + 8
            if level > 0:
                self._chunks(child, rank1, cdims, out)
            else:
                chunk = np.frombuffer(self.b, dtype=out.dtype, count=int(np.prod(cdims)), offset=child + self.base).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

    def resolve(self, path, start=None):
        addr = self.root if start is None else start
        for part in [q for q in path.split("/") if q]:
            ln = self.links(addr)
            if part not in ln:
                raise KeyError(path)
            addr = ln[part]
        return addr


def read_keras_weights(path):
    """{'<layer>/<weight>': array} from a Keras `save_weights` file or the `model_weights` group of a `model.save` file.
    Keras stores weight 'conv2d_1/kernel:0' of layer 'conv2d_1' as dataset /conv2d_1/conv2d_1/kernel:0 and lists it in
    the layer group's `weight_names` attribute; the ':0' suffix is dropped."""
    r = Reader(path)
    root = r.root
    top = r.links(root)
    if "model_weights" in top:
        root = top["model_weights"]
        top = r.links(root)
    names = r.attrs(root).get("layer_names")
    layer_names = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in (list(names.ravel()) if names is not None else sorted(top))]
    out = {}
    for lname in layer_names:
        if lname not in top:
            continue
        g = top[lname]
        wn = r.attrs(g).get("weight_names")
        if wn is None:
            continue
        for w in (list(wn.ravel()) if hasattr(wn, "ravel") else [wn]):
            w = w.decode("utf8") if isinstance(w, bytes) else str(w)
            arr = r.dataset(r.resolve(w, start=g))
            key = w[:-2] if w.endswith(":0") else w
            out[key] = np.ascontiguousarray(arr)
    return out, r.attrs(r.root)


# =====================================================================================================================
# writer (superblock 0, symbol-table groups, contiguous datasets, fixed-length string attributes)
# =====================================================================================================================
class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def align(self, n=8):
        while len(self.b) % n:
            self.b.append(0)

    def write(self, data):
        p = len(self.b)
        self.b += data
        return p


def _dtype_msg(dt):
    if isinstance(dt, tuple):  # ('S', n): fixed-length, null-padded ASCII
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt[1])
    dt = np.dtype(dt)
    if dt.kind == "f":
        sz = dt.itemsize
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[sz]
        return struct.pack("<BBBBI", 0x11, 0x20, 8 * sz - 1, 0, sz) + struct.pack("<HHBBBBI", 0, 8 * sz, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError("unsupported dtype %s" % dt)


def _dspace_msg(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _pad8(b):
    return b + b"\x00" * ((-len(b)) % 8)


def _attr_msg(name, value):
    nm = name.encode("utf8") + b"\x00"
    if isinstance(value, (bytes, str)):
        v = value.encode("utf8") if isinstance(value, str) else value
        dt, shape, data = ("S", max(1, len(v))), (), v.ljust(max(1, len(v)), b"\x00")
    elif isinstance(value, (list, tuple)) and (len(value) == 0 or isinstance(value[0], (bytes, str))):
        vs = [x.encode("utf8") if isinstance(x, str) else x for x in value]
        L = max([len(x) for x in vs] + [1])
        dt, shape, data = ("S", L), (len(vs),), b"".join(x.ljust(L, b"\x00") for x in vs)
    else:
        a = np.ascontiguousarray(value)
        dt, shape, data = a.dtype, a.shape, a.tobytes()
    t, s = _dtype_msg(dt), _dspace_msg(shape)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(t), len(s)) + _pad8(nm) + _pad8(t) + _pad8(s) + data


def _object_header(buf, msgs):
    """msgs: [(type, payload bytes)] -> address of a version-1 object header holding them."""
    body = b""
    for mtype, payload in msgs:
        payload = _pad8(payload)
        body += struct.pack("<HHBBBB", mtype, len(payload), 0, 0, 0, 0) + payload
    buf.align(8)
    addr = buf.write(struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4)
    buf.write(body)
    return addr


class Writer:
    """w = Writer(); g = w.group(parent, name, attrs); w.dataset(g, name, array); w.save(path). Parents are handles returned
    by group(); the root handle is w.root."""
    LEAF_K, NODE_K = 4, 16

    def __init__(self):
        self.root = {"children": {}, "attrs": {}}

    def group(self, parent, name, attrs=None):
        g = {"children": {}, "attrs": dict(attrs or {})}
        parent["children"][name] = g
        return g

    def dataset(self, parent, name, array):
        parent["children"][name] = np.ascontiguousarray(array)

    def _emit_dataset(self, buf, a):
        buf.align(8)
        data_at = buf.write(a.tobytes()) if a.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, data_at, a.nbytes)
        # fill value (old, version 2: allocate late, never write, undefined) keeps h5py / libhdf5 happy
        fill = struct.pack("<BBBB", 2, 2, 0, 0)
        return _object_header(buf, [(0x0001, _dspace_msg(a.shape)), (0x0003, _dtype_msg(a.dtype)), (0x0005, fill), (0x0008, layout)])

    def _emit_group(self, buf, g):
        entries = []
        for name in sorted(g["children"], key=lambda s: s.encode("utf8")):
            c = g["children"][name]
            entries.append((name, self._emit_dataset(buf, c) if isinstance(c, np.ndarray) else self._emit_group(buf, c)))
        # local heap: offset 0 is the empty string (the B-tree's first key), then the names
        heap = bytearray(b"\x00" * 8)
        offs = []
        for name, _ in entries:
            offs.append(len(heap))
            heap += name.encode("utf8") + b"\x00"
            while len(heap) % 8:
                heap.append(0)
        free_at = len(heap)
        heap += struct.pack("<QQ", 1, 16)            # one free block: next = 1 (end of list), size 16
        buf.align(8)
        heap_data_at = buf.write(bytes(heap))
        buf.align(8)
        heap_at = buf.write(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free_at, heap_data_at))
        # symbol-table nodes of at most 2 * LEAF_K entries, one leaf-level B-tree node above them
        snods, keys = [], [0]
        per = 2 * self.LEAF_K
        for i in range(0, max(len(entries), 1), per):
            chunk = list(zip(entries[i:i + per], offs[i:i + per]))
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for (name, addr), off in chunk:
                body += struct.pack("<QQII", off, addr, 0, 0) + b"\x00" * 16
            body += b"\x00" * (40 * (per - len(chunk)))
            buf.align(8)
            snods.append(buf.write(body))
            keys.append(chunk[-1][1] if chunk else 0)
        # B-tree of the symbol-table nodes: leaf-level nodes hold up to 2 * NODE_K of them, further levels above as needed
        # (key i = heap offset of the largest name in child i - 1; key 0 = the empty string)
        level, nodes = 0, [(keys[i], keys[i + 1], s) for i, s in enumerate(snods)]   # (left key, right key, address)
        cap = 2 * self.NODE_K
        while True:
            parents = []
            for i in range(0, len(nodes), cap):
                ch = nodes[i:i + cap]
                tree = b"TREE" + struct.pack("<BBHQQ", 0, level, len(ch), UNDEF, UNDEF)
                for lk, rk, a in ch:
                    tree += struct.pack("<QQ", lk, a)
                tree += struct.pack("<Q", ch[-1][1])
                tree += b"\x00" * (16 * (cap - len(ch)))
                buf.align(8)
                parents.append((ch[0][0], ch[-1][1], buf.write(tree)))
            for j, (_, _, a) in enumerate(parents):   # sibling links of the level just written
                left = parents[j - 1][2] if j > 0 else UNDEF
                right = parents[j + 1][2] if j + 1 < len(parents) else UNDEF
                buf.b[a + 8:a + 24] = struct.pack("<QQ", left, right)
            if len(parents) == 1:
                tree_at = parents[0][2]
                break
            nodes, level = parents, level + 1
        msgs = [(0x0011, struct.pack("<QQ", tree_at, heap_at))] + [(0x000C, _attr_msg(k, v)) for k, v in g["attrs"].items()]
        g["_tree"], g["_heap"] = tree_at, heap_at
        return _object_header(buf, msgs)

    def save(self, path):
        buf = _Buf()
        buf.write(b"\x00" * 96)  # superblock (56 bytes) + root symbol-table entry (40 bytes), filled in below
        root_at = self._emit_group(buf, self.root)
        buf.align(8)
        eof = buf.tell()
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.NODE_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_at, 1, 0) + struct.pack("<QQ", self.root["_tree"], self.root["_heap"])
        assert len(sb) == 96, len(sb)
        buf.b[:96] = sb
        with open(path, "wb") as f:
            f.write(bytes(buf.b))


def write_keras_weights(path, layers, weights, full_model_config=None):
    """layers: ordered [(layer name, [weight keys '<layer>/<weight>'])] INCLUDING weight-less layers (Keras lists every
    layer in `layer_names`); weights: {key: array}. full_model_config (a JSON string) writes the `model.save` layout:
    the weights go under /model_weights and the root carries `model_config`."""
    w = Writer()
    root = w.root
    if full_model_config is not None:
        w.root["attrs"].update({"keras_version": b"2.1.3", "backend": b"tensorflow", "model_config": full_model_config})
        root = w.group(w.root, "model_weights")
    root["attrs"].update({"layer_names": [n for n, _ in layers], "backend": b"tensorflow", "keras_version": b"2.1.3"})
    for lname, keys in layers:
        g = w.group(root, lname, {"weight_names": [k + ":0" for k in keys]})
        if keys:
            inner = w.group(g, lname)   # 'conv2d_1/kernel:0' inside group 'conv2d_1' -> /conv2d_1/conv2d_1/kernel:0
            for k in keys:
                assert k.split("/")[0] == lname, (k, lname)
                w.dataset(inner, k.split("/", 1)[1] + ":0", np.asarray(weights[k], np.float32))
    w.save(path)
