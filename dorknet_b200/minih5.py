"""minih5: the subset of HDF5 (and of the h5py API) that Dorknet checkpoints use, in pure Python + NumPy.

The reference saves / loads weights through h5py (network/feed_forward_network.py:90-139 and every layer's
save_to_h5 / load_from_h5, e.g. layers/convolution.py:226-281, layers/batch_norm.py:176-232).  h5py is not part of this
image, and a checkpoint reader / writer has no business on the GPU anyway, so this module implements the file format
itself (HDF5 File Format Specification 1.x, "earliest" library-version layout -- what h5py writes by default and every
HDF5 library reads):

    superblock v0 - old-style groups (symbol-table message, v1 B-tree, SNOD nodes, local heap) - v1 object headers -
    dataspace v1 (simple) / v2 (null) - datatypes: IEEE floats, integers, fixed strings, variable-length strings (global
    heap), the int8 enum h5py uses for bool - contiguous (and, reading, compact) data layout - attribute messages v1-v3 -
    object header continuation blocks.

API (h5py-compatible for what the reference touches):  File(name, "r" | "w") as a context manager; f.create_dataset(path,
shape=None, dtype=None, data=None) with intermediate groups created on the way; f[path] -> Dataset | Group; dset[:] /
dset[...] read, dset[:] = array write; dset.shape / .dtype; obj.attrs[key] get / set, .attrs.get(key, default), `in`,
.keys(); group.keys() / `in`.  Chunked, compressed, v2-B-tree / fractal-heap ("latest" format) files are rejected with a
clear error.  dorknet_b200.dropin installs this module as `h5py` when the real one is missing.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
LEAF_K, INT_K = 4, 16


def _pad8(n):
    return (n + 7) & ~7


class Empty:
    """value of a dataset / attribute with a NULL dataspace (h5py.Empty)"""

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)


# =========================================================================================== datatypes
def _enc_dtype(dt):
    """numpy dtype (or the markers 'vstr', ('fstr', n), 'bool') -> datatype message bytes"""
    if dt == "vstr":  # variable-length UTF-8 string: class 9, base type = 1-byte UTF-8 string
        base = struct.pack("<B3BI", 0x13, 0x10, 0, 0, 1)
        return struct.pack("<B3BI", 0x19, 0x01, 0x01, 0, 16) + base
    if isinstance(dt, tuple) and dt[0] == "fstr":  # fixed-length ASCII, null-padded
        return struct.pack("<B3BI", 0x13, 0x01, 0, 0, max(int(dt[1]), 1))
    if dt == "bool":  # h5py: enum of int8 {FALSE = 0, TRUE = 1}
        base = struct.pack("<B3BIHH", 0x10, 0x08, 0, 0, 1, 0, 8)
        names = b"FALSE\0\0\0" + b"TRUE\0\0\0\0"
        return struct.pack("<B3BI", 0x18, 2, 0, 0, 1) + base + names + b"\x00\x01"
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 4:
            return struct.pack("<B3BIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        if dt.itemsize == 8:
            return struct.pack("<B3BIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        if dt.itemsize == 2:
            return struct.pack("<B3BIHHBBBBI", 0x11, 0x20, 15, 0, 2, 0, 16, 10, 5, 0, 10, 15)
    if dt.kind in "iu":
        return struct.pack("<B3BIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    raise TypeError("minih5: no HDF5 equivalent for dtype %r" % (dt,))


def _dec_dtype(buf, off=0):
    """-> (descr, message length).  descr: np.dtype | 'vstr' | ('fstr', n) | 'bool' | ('enum', base dtype)"""
    cv, b0, b1, b2, size = struct.unpack_from("<B3BI", buf, off)
    cls, ver = cv & 0x0F, cv >> 4
    if cls == 0:
        if b0 & 1:
            raise NotImplementedError("minih5: big-endian integers")
        return np.dtype("%s%d" % ("i" if b0 & 8 else "u", size)).newbyteorder("<"), 12
    if cls == 1:
        if b0 & 1:
            raise NotImplementedError("minih5: big-endian floats")
        return np.dtype("<f%d" % size), 20
    if cls == 3:
        return ("fstr", size), 8
    if cls == 9:
        base, blen = _dec_dtype(buf, off + 8)
        if (b0 & 0x0F) == 1:
            return "vstr", 8 + blen
        raise NotImplementedError("minih5: variable-length sequences")
    if cls == 8:
        nmemb = b0 | (b1 << 8)
        base, blen = _dec_dtype(buf, off + 8)
        p = off + 8 + blen
        names = []
        for _ in range(nmemb):
            e = buf.index(b"\0", p)
            names.append(bytes(buf[p:e]))
            p = e + 1 if ver >= 3 else p + _pad8(e + 1 - p)
        p += nmemb * base.itemsize
        if sorted(names) == [b"FALSE", b"TRUE"] and base.itemsize == 1:
            return "bool", p - off
        return ("enum", base), p - off
    raise NotImplementedError("minih5: datatype class %d" % cls)


def _np_dtype(descr):
    if isinstance(descr, np.dtype):
        return descr
    if descr == "bool":
        return np.dtype(bool)
    if descr == "vstr":
        return np.dtype(object)
    if descr[0] == "fstr":
        return np.dtype("S%d" % descr[1])
    return descr[1]


# =========================================================================================== reading
class _Reader:
    def __init__(self, data):
        self.d = data
        base = 0
        while True:  # the superblock may sit at 0, 512, 1024, ... (user block)
            if self.d[base:base + 8] == SIG:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.d):
                raise OSError("minih5: not an HDF5 file")
        ver = self.d[base + 8]
        if ver > 1:
            raise NotImplementedError("minih5: superblock version %d (file written with libver='latest'?)" % ver)
        so, sl = self.d[base + 13], self.d[base + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError("minih5: offsets / lengths of %d / %d bytes" % (so, sl))
        p = base + 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", self.d, p)[0]
        root = p + 32  # base, free-space, eof, driver addresses, then the root symbol-table entry
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", self.d, root)
        self.root_header += self.base

    # -- object headers ------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, payload bytes)] of a version-1 object header, continuation blocks followed"""
        if self.d[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("minih5: version-2 object headers (file written with libver='latest')")
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.d, addr)
        if ver != 1:
            raise OSError("minih5: bad object header at %d" % addr)
        out, blocks = [], [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.d, p)
                body = self.d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    o, l = struct.unpack_from("<QQ", body)
                    blocks.append((o + self.base, l))
                out.append((mtype, flags, body))
        return out

    def heap_name(self, heap_addr, off):
        assert self.d[heap_addr:heap_addr + 4] == b"HEAP"
        seg = struct.unpack_from("<Q", self.d, heap_addr + 24)[0] + self.base
        e = self.d.index(b"\0", seg + off)
        return self.d[seg + off:e].decode("utf-8")

    def group_entries(self, btree, heap):
        """{name: object header address} of an old-style group"""
        out = {}

        def walk(node):
            sig = self.d[node:node + 4]
            if sig == b"SNOD":
                n = struct.unpack_from("<H", self.d, node + 6)[0]
                for i in range(n):
                    noff, ohdr = struct.unpack_from("<QQ", self.d, node + 8 + 40 * i)
                    out[self.heap_name(heap, noff)] = ohdr + self.base
                return
            if sig != b"TREE":
                raise OSError("minih5: bad group B-tree node at %d" % node)
            used = struct.unpack_from("<H", self.d, node + 6)[0]
            for i in range(used):
                walk(struct.unpack_from("<Q", self.d, node + 24 + 8 + 16 * i)[0] + self.base)
        walk(btree)
        return out

    def gheap(self, coll, index):
        coll += self.base
        if self.d[coll:coll + 4] != b"GCOL":
            raise OSError("minih5: bad global heap collection at %d" % coll)
        size = struct.unpack_from("<Q", self.d, coll + 8)[0]
        p, end = coll + 16, coll + size
        while p + 16 <= end:
            idx, _, _, osize = struct.unpack_from("<HHIQ", self.d, p)
            if idx == index:
                return self.d[p + 16:p + 16 + osize]
            if idx == 0:
                break
            p += 16 + _pad8(osize)
        raise OSError("minih5: global heap object %d not found" % index)

    # -- values --------------------------------------------------------------------------------------
    @staticmethod
    def dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            return tuple(struct.unpack_from("<%dQ" % rank, body, 8)) if rank else ()
        if ver == 2:
            if body[3] == 2:
                return None  # NULL dataspace
            return tuple(struct.unpack_from("<%dQ" % rank, body, 4)) if rank else ()
        raise NotImplementedError("minih5: dataspace version %d" % ver)

    def decode(self, descr, shape, raw):
        if shape is None:
            return Empty(_np_dtype(descr))
        n = int(np.prod(shape)) if shape else 1
        if descr == "vstr":
            vals = []
            for i in range(n):
                ln, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(bytes(self.gheap(coll, idx)[:ln]).decode("utf-8") if ln else "")
            if shape == ():
                return vals[0]
            a = np.empty(n, dtype=object)
            a[:] = vals
            return a.reshape(shape)
        dt = _np_dtype(descr)
        a = np.frombuffer(raw, dtype=dt if descr != "bool" else np.int8, count=n).reshape(shape)
        if descr == "bool":
            a = a.astype(bool)
        if isinstance(descr, tuple) and descr[0] == "fstr" and shape == ():
            return np.bytes_(bytes(a[()]).rstrip(b"\0"))
        return a[()] if shape == () else a.copy()

    def attribute(self, body):
        ver = body[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8
            name = bytes(body[p:p + nsz]).split(b"\0")[0].decode("utf-8")
            p += _pad8(nsz)
            t0 = p
            p += _pad8(tsz)
            s0 = p
            p += _pad8(ssz)
        elif ver in (2, 3):
            if body[1] & 3:
                raise NotImplementedError("minih5: shared attribute datatypes")
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = bytes(body[p:p + nsz]).split(b"\0")[0].decode("utf-8")
            p += nsz
            t0 = p
            p += tsz
            s0 = p
            p += ssz
        else:
            raise NotImplementedError("minih5: attribute message version %d" % ver)
        descr, _ = _dec_dtype(body, t0)
        shape = self.dataspace(body[s0:s0 + ssz])
        return name, self.decode(descr, shape, body[p:])


class _ReadAttrs:
    def __init__(self, d):
        self._d = d

    def __getitem__(self, k):
        return self._d[k]

    def get(self, k, default=None):
        return self._d.get(k, default)

    def __contains__(self, k):
        return k in self._d

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)


class _ReadObject:
    def __init__(self, rd, addr, name):
        self._rd, self._addr, self.name = rd, addr, name
        self._msgs = rd.messages(addr)
        self.attrs = _ReadAttrs(dict(rd.attribute(b) for t, _, b in self._msgs if t == 0x000C))

    def _open(self, name, addr):
        msgs = self._rd.messages(addr)
        if any(t == 0x0011 for t, _, _ in msgs):
            return ReadGroup(self._rd, addr, name)
        if any(t == 0x0002 for t, _, _ in msgs) and not any(t == 0x0008 for t, _, _ in msgs):
            raise NotImplementedError("minih5: new-style (link message) groups: file written with libver='latest'")
        return ReadDataset(self._rd, addr, name)


class ReadGroup(_ReadObject):
    def __init__(self, rd, addr, name):
        super().__init__(rd, addr, name)
        st = [b for t, _, b in self._msgs if t == 0x0011]
        if not st:
            raise NotImplementedError("minih5: group without a symbol-table message (libver='latest' file?)")
        bt, hp = struct.unpack_from("<QQ", st[0])
        self._entries = rd.group_entries(bt + rd.base, hp + rd.base)

    def keys(self):
        return sorted(self._entries)

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        obj = self
        for part in [q for q in path.split("/") if q]:
            if not isinstance(obj, ReadGroup) or part not in obj._entries:
                raise KeyError("Unable to open object (object %r doesn't exist)" % path)
            obj = obj._open((obj.name.rstrip("/") + "/" + part), obj._entries[part])
        return obj


class ReadDataset(_ReadObject):
    def __init__(self, rd, addr, name):
        super().__init__(rd, addr, name)
        m = {t: b for t, _, b in self._msgs}
        self._descr, _ = _dec_dtype(m[0x0003])
        self.shape = rd.dataspace(m[0x0001])
        self.dtype = _np_dtype(self._descr)
        lay = m[0x0008]
        if lay[0] != 3:
            raise NotImplementedError("minih5: data layout message version %d" % lay[0])
        if lay[1] == 1:
            a, n = struct.unpack_from("<QQ", lay, 2)
            self._raw = b"" if a == UNDEF else rd.d[a + rd.base:a + rd.base + n]
        elif lay[1] == 0:
            n = struct.unpack_from("<H", lay, 2)[0]
            self._raw = lay[4:4 + n]
        else:
            raise NotImplementedError("minih5: chunked datasets")
        if any(t == 0x000B for t in m):
            raise NotImplementedError("minih5: filtered (compressed) datasets")

    def __getitem__(self, key):
        v = self._rd.decode(self._descr, self.shape, self._raw)
        if isinstance(v, Empty) or key is Ellipsis or key == () or (isinstance(key, slice) and key == slice(None)):
            return v
        return v[key]

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self[...])
        return a.astype(dtype) if dtype is not None else a


# =========================================================================================== writing
class _WAttrs(dict):
    def get(self, k, default=None):
        return dict.get(self, k, default)


class WDataset:
    def __init__(self, name, shape, dtype, data=None):
        self.name = name
        self.dtype = np.dtype(dtype if dtype is not None else np.float32)
        self.shape = None if shape is None else tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
        self.attrs = _WAttrs()
        self._data = None
        if self.shape is not None:
            self._data = np.zeros(self.shape, self.dtype)
            if data is not None:
                self._data[...] = data

    def __setitem__(self, key, value):
        if self.shape is None:
            raise TypeError("minih5: cannot write to an empty (NULL dataspace) dataset")
        self._data[key] = np.asarray(value)

    def __getitem__(self, key):
        return Empty(self.dtype) if self.shape is None else self._data[key]


class WGroup:
    def __init__(self, name):
        self.name = name
        self.attrs = _WAttrs()
        self._children = {}

    def _walk(self, path, create):
        g = self
        parts = [q for q in path.split("/") if q]
        for part in parts[:-1]:
            nxt = g._children.get(part)
            if nxt is None:
                if not create:
                    raise KeyError(path)
                nxt = g._children[part] = WGroup(g.name.rstrip("/") + "/" + part)
            if not isinstance(nxt, WGroup):
                raise ValueError("minih5: %r is a dataset, not a group" % part)
            g = nxt
        return g, parts[-1]

    def create_group(self, path):
        g, last = self._walk(path, True)
        if last in g._children:
            raise ValueError("Unable to create group (name already exists)")
        g._children[last] = WGroup(g.name.rstrip("/") + "/" + last)
        return g._children[last]

    def create_dataset(self, path, shape=None, dtype=None, data=None):
        if data is not None and shape is None:
            data = np.asarray(data)
            shape, dtype = data.shape, dtype or data.dtype
        g, last = self._walk(path, True)
        if last in g._children:
            raise ValueError("Unable to create dataset (name already exists)")
        d = g._children[last] = WDataset(g.name.rstrip("/") + "/" + last, shape, dtype, data)
        return d

    def __getitem__(self, path):
        g, last = self._walk(path, False)
        if last not in g._children:
            raise KeyError(path)
        return g._children[last]

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return sorted(self._children)


class _Writer:
    """lays the whole tree out in one buffer at close()"""

    def __init__(self):
        self.buf = bytearray()
        self.gheap_objs = []  # payloads of the single global heap collection (vlen strings)

    def alloc(self, n, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf)
        self.buf.extend(b"\0" * n)
        return off

    def put(self, off, data):
        self.buf[off:off + len(data)] = data

    # -- attribute values -> (datatype msg, dataspace msg, data, vlen fixups) ------------------------------
    def _attr_parts(self, value):
        if isinstance(value, Empty):
            return _enc_dtype(value.dtype), struct.pack("<BBBB", 2, 0, 0, 2), b"", []
        if isinstance(value, str):
            return _enc_dtype("vstr"), struct.pack("<BBB5x", 1, 0, 0), b"\0" * 16, [(0, value.encode("utf-8"))]
        if isinstance(value, (bytes, np.bytes_)):
            b = bytes(value)
            return _enc_dtype(("fstr", len(b))), struct.pack("<BBB5x", 1, 0, 0), b if b else b"\0", []
        if isinstance(value, (bool, np.bool_)):
            return _enc_dtype("bool"), struct.pack("<BBB5x", 1, 0, 0), b"\x01" if value else b"\x00", []
        if isinstance(value, (list, tuple)) and value and all(isinstance(v, str) for v in value):
            space = struct.pack("<BBB5xQ", 1, 1, 0, len(value))
            return _enc_dtype("vstr"), space, b"\0" * (16 * len(value)), [(16 * i, v.encode("utf-8")) for i, v in enumerate(value)]
        a = np.asarray(value)
        if a.dtype == object or a.dtype.kind in "US":
            raise TypeError("minih5: attribute value %r has no HDF5 equivalent here" % (value,))
        if a.dtype == bool:
            dt, raw = _enc_dtype("bool"), a.astype(np.int8).tobytes()
        else:
            if a.dtype.kind == "i" and isinstance(value, int):
                a = a.astype(np.int64)
            dt, raw = _enc_dtype(a.dtype), np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tobytes()
        if a.ndim == 0:
            space = struct.pack("<BBB5x", 1, 0, 0)
        else:
            space = struct.pack("<BBB5x%dQ" % a.ndim, 1, a.ndim, 0, *a.shape)
        return dt, space, raw, []

    def _attr_msg(self, name, value):
        dt, space, raw, fix = self._attr_parts(value)
        nm = name.encode("utf-8") + b"\0"
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(space))
        body = bytearray(head + nm.ljust(_pad8(len(nm)), b"\0") + dt.ljust(_pad8(len(dt)), b"\0") + space.ljust(_pad8(len(space)), b"\0"))
        data_off = len(body)
        body += raw
        fixups = []
        for off, payload in fix:  # vlen element: length, collection address (patched later), object index
            self.gheap_objs.append(payload)
            idx = len(self.gheap_objs)
            struct.pack_into("<I", body, data_off + off, len(payload))
            struct.pack_into("<I", body, data_off + off + 12, idx)
            fixups.append(data_off + off + 4)
        return bytes(body), fixups

    def _header(self, msgs):
        """msgs: [(type, payload, gheap fixup offsets)] -> address; gheap address fields are recorded for patching"""
        size = sum(8 + _pad8(len(b)) for _, b, _ in msgs)
        addr = self.alloc(16 + size)
        self.put(addr, struct.pack("<BBHII", 1, 0, len(msgs), 1, size))
        p = addr + 16
        for t, b, fix in msgs:
            self.put(p, struct.pack("<HHB3x", t, _pad8(len(b)), 0))
            self.put(p + 8, b)
            for f in fix:
                self.gheap_fixups.append(p + 8 + f)
            p += 8 + _pad8(len(b))
        return addr

    gheap_fixups = None

    def write_dataset(self, d):
        msgs = []
        if d.shape is None:
            msgs.append((0x0001, struct.pack("<BBBB", 2, 0, 0, 2), []))
        else:
            msgs.append((0x0001, struct.pack("<BBB5x%dQ" % len(d.shape), 1, len(d.shape), 0, *d.shape), []))
        msgs.append((0x0003, _enc_dtype(d.dtype), []))
        msgs.append((0x0005, struct.pack("<BBBB", 2, 2, 2, 0), []))  # fill value: allocate late, never written, undefined
        if d.shape is None or d._data.size == 0:
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, UNDEF, 0), []))
        else:
            raw = np.ascontiguousarray(d._data).astype(d.dtype.newbyteorder("<")).tobytes()
            a = self.alloc(len(raw))
            self.put(a, raw)
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, a, len(raw)), []))
        for k, v in d.attrs.items():
            b, fix = self._attr_msg(k, v)
            msgs.append((0x000C, b, fix))
        return self._header(msgs)

    def write_group(self, g):
        names = sorted(g._children, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * LEAF_K * 2 * INT_K:
            raise NotImplementedError("minih5: more than %d links in one group" % (2 * LEAF_K * 2 * INT_K))
        child_addr = {}
        for n in names:
            c = g._children[n]
            child_addr[n] = self.write_group(c)[0] if isinstance(c, WGroup) else self.write_dataset(c)
        # local heap: "" at offset 0, then the names
        seg = bytearray(b"\0" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(seg)
            b = n.encode("utf-8") + b"\0"
            seg += b.ljust(_pad8(len(b)), b"\0")
        seg_size = max(_pad8(len(seg)) + 16, 88)
        free_off = len(seg)
        seg = seg.ljust(seg_size, b"\0")
        struct.pack_into("<QQ", seg, free_off, 1, seg_size - free_off)  # one free block: next = 1 (none), size
        seg_addr = self.alloc(seg_size)
        self.put(seg_addr, bytes(seg))
        heap_addr = self.alloc(32)
        self.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, seg_size, free_off, seg_addr))
        # symbol-table nodes of up to 2*LEAF_K entries, one B-tree node above them
        chunks = [names[i:i + 2 * LEAF_K] for i in range(0, len(names), 2 * LEAF_K)]
        snods = []
        for ch in chunks:
            a = self.alloc(8 + 40 * 2 * LEAF_K)
            self.put(a, b"SNOD" + struct.pack("<BBH", 1, 0, len(ch)))
            for i, n in enumerate(ch):
                c = g._children[n]
                if isinstance(c, WGroup):
                    entry = struct.pack("<QQIIQQ", name_off[n], child_addr[n], 1, 0, *self._group_scratch[child_addr[n]])
                else:
                    entry = struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0)
                self.put(a + 8 + 40 * i, entry)
            snods.append(a)
        bt = self.alloc(24 + 8 * (2 * INT_K + 1) + 8 * 2 * INT_K)
        self.put(bt, b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF))
        p = bt + 24
        self.put(p, struct.pack("<Q", 0))
        p += 8
        for a, ch in zip(snods, chunks):
            self.put(p, struct.pack("<QQ", a, name_off[ch[-1]]))
            p += 16
        msgs = [(0x0011, struct.pack("<QQ", bt, heap_addr), [])]
        for k, v in g.attrs.items():
            b, fix = self._attr_msg(k, v)
            msgs.append((0x000C, b, fix))
        hdr = self._header(msgs)
        self._group_scratch[hdr] = (bt, heap_addr)
        return hdr, bt, heap_addr

    def build(self, root):
        self.buf = bytearray(b"\0" * 96)  # superblock v0 (56 bytes + 40-byte root entry)
        self.gheap_fixups = []
        self._group_scratch = {}
        hdr, bt, heap = self.write_group(root)
        if self.gheap_objs:
            body = bytearray()
            for i, payload in enumerate(self.gheap_objs):
                body += struct.pack("<HHIQ", i + 1, 1, 0, len(payload)) + payload.ljust(_pad8(len(payload)), b"\0")
            size = max(4096, _pad8(16 + len(body) + 16))
            coll = self.alloc(size)
            free = size - 16 - len(body)
            self.put(coll, b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body) + struct.pack("<HHIQ", 0, 0, 0, free))
            for f in self.gheap_fixups:
                struct.pack_into("<Q", self.buf, f, coll)
        eof = len(self.buf)
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, bt, heap)
        self.put(0, sb)
        return bytes(self.buf)


# =========================================================================================== File
class File:
    """h5py.File for modes "r" and "w"."""

    def __init__(self, name, mode="r", **_):
        self.filename, self.mode = name, mode
        if mode == "r":
            with open(name, "rb") as f:
                rd = _Reader(memoryview(f.read()).tobytes())
            self._root = ReadGroup(rd, rd.root_header, "/")
        elif mode in ("w", "w-", "x"):
            self._root = WGroup("/")
        else:
            raise ValueError("minih5: mode %r is not supported (use 'r' or 'w')" % mode)
        self.attrs = self._root.attrs

    def create_dataset(self, *a, **k):
        return self._root.create_dataset(*a, **k)

    def create_group(self, path):
        return self._root.create_group(path)

    def __getitem__(self, path):
        return self._root[path]

    def __contains__(self, path):
        return path in self._root

    def keys(self):
        return self._root.keys()

    def close(self):
        if self.mode != "r" and self._root is not None:
            data = _Writer().build(self._root)
            with open(self.filename, "wb") as f:
                f.write(data)
        self._root = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None or self.mode == "r":
            self.close()
        return False
