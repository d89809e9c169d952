"""TEST INFRASTRUCTURE: a minimal structural HDF5 reader (version-0 superblock, v1 object headers, symbol-table
groups, contiguous and chunked / deflated datasets, fixed- and variable-length string attributes with their global heap),
written from the HDF5 File Format Specification independently of the emitter's code paths: it walks the file the way
libhdf5 does (superblock -> root symbol-table entry -> B-tree -> SNOD -> object headers -> chunk B-tree / global heap).
h5py / libhdf5 are not available offline; this reader is what checks the emitted files (tools/verify_h5.py runs the
reference's own accessors wherever h5py exists)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Min:
    def __init__(self, path, userblock=0):
        """userblock: size of a user block in front of the superblock (file addresses are relative to it)."""
        whole = open(path, "rb").read()
        self.b = whole[userblock:]
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "bad signature"
        assert b[8] == 0 and b[13] == 8 and b[14] == 8, "superblock v0 with 8-byte offsets/lengths expected"
        self.k_leaf, self.k_int = struct.unpack_from("<HH", b, 16)
        base, freesp, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        assert base == userblock and freesp == UNDEF and drv == UNDEF and eof == len(whole), (base, freesp, eof, len(whole))
        name_off, hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        assert cache == 1
        self.root = self._object(hdr)

    def _heap_str(self, heap_addr, off):
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b"HEAP" and b[heap_addr + 4] == 0
        size, free, data = struct.unpack_from("<QQQ", b, heap_addr + 8)
        assert free == 1 or free < size
        assert off < size
        end = b.index(b"\0", data + off)
        return b[data + off:end].decode()

    def _group(self, btree, heap):
        b = self.b
        assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0, "group B-tree node expected"
        level, used = b[btree + 5], struct.unpack_from("<H", b, btree + 6)[0]
        left, right = struct.unpack_from("<QQ", b, btree + 8)
        assert left == UNDEF and right == UNDEF and used <= 2 * self.k_int
        out = {}
        p = btree + 24
        prev_key = self._heap_str(heap, struct.unpack_from("<Q", b, p)[0])
        p += 8
        for _ in range(used):
            child, key = struct.unpack_from("<QQ", b, p)
            p += 16
            key_name = self._heap_str(heap, key)
            if level > 0:
                sub = self._group(child, heap)
            else:
                assert b[child:child + 4] == b"SNOD" and b[child + 4] == 1
                n = struct.unpack_from("<H", b, child + 6)[0]
                assert n <= 2 * self.k_leaf
                sub, names = {}, []
                for i in range(n):
                    no, hdr, cache, _ = struct.unpack_from("<QQII", b, child + 8 + 40 * i)
                    nm = self._heap_str(heap, no)
                    names.append(nm)
                    sub[nm] = self._object(hdr)
                assert names == sorted(names), "symbol table entries must be sorted"
                assert all(prev_key < nm <= key_name for nm in names), "B-tree keys must bracket the node's names"
            out.update(sub)
            prev_key = key_name
        return out

    def _messages(self, addr):
        b = self.b
        assert b[addr] == 1, "object header v1"
        nmsg, refc, size = struct.unpack_from("<HII", b, addr + 2)
        p, end, out = addr + 16, addr + 16 + size, []
        for _ in range(nmsg):
            t, sz, flags = struct.unpack_from("<HHB", b, p)
            assert sz % 8 == 0
            out.append((t, b[p + 8:p + 8 + sz]))
            p += 8 + sz
        assert p == end
        return out

    @staticmethod
    def _dataspace(m, want_max=False):
        assert m[0] == 1
        rank, flags = m[1], m[2]
        assert flags in (0, 1)
        dims = list(struct.unpack_from("<%dQ" % rank, m, 8))
        if want_max:
            return dims, (list(struct.unpack_from("<%dQ" % rank, m, 8 + 8 * rank)) if flags & 1 else None)
        return dims

    def _global_heap_object(self, addr, index):
        """Object `index` of the global heap collection at `addr` (walks the collection like H5HG does)."""
        b = self.b
        assert b[addr:addr + 4] == b"GCOL" and b[addr + 4] == 1
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        assert size >= 4096 and addr + size <= len(b)
        p, end, found, seen_free = addr + 16, addr + size, None, False
        while p + 16 <= end:
            idx, refs, _, osz = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:  # free space: its size includes its own header and must run to the end of the collection
                assert p + osz == end
                seen_free = True
                break
            assert refs >= 1
            if idx == index:
                found = b[p + 16:p + 16 + osz]
            p += 16 + ((osz + 7) & ~7)
        assert seen_free or p == end
        assert found is not None, "global heap object %d not found" % index
        return found

    def _chunked(self, lay, dims, pline):
        """Layout class 2: rank-1 chunks through the v1 chunk B-tree (node type 1)."""
        b = self.b
        ndim = lay[2]
        assert ndim == len(dims) + 1 == 2
        bt = struct.unpack_from("<Q", lay, 3)[0]
        cdim, esz = struct.unpack_from("<II", lay, 11)
        assert esz == 8
        out = np.zeros(dims[0])
        if bt == UNDEF:
            assert dims[0] == 0
            return out, cdim
        covered = []

        def walk(node, lo_expect):
            assert b[node:node + 4] == b"TREE" and b[node + 4] == 1
            level, used = b[node + 5], struct.unpack_from("<H", b, node + 6)[0]
            assert 1 <= used <= 64
            p = node + 24
            keys, kids = [], []
            for i in range(used + 1):
                nbytes, mask, off, zero = struct.unpack_from("<IIQQ", b, p)
                assert zero == 0 and off % cdim == 0
                keys.append((nbytes, mask, off)); p += 24
                if i < used:
                    kids.append(struct.unpack_from("<Q", b, p)[0]); p += 8
            assert keys[0][2] == lo_expect and all(keys[i][2] < keys[i + 1][2] for i in range(used))
            for i in range(used):
                if level > 0:
                    walk(kids[i], keys[i][2])
                else:
                    nbytes, mask, off = keys[i]
                    assert mask == 0
                    raw = b[kids[i]:kids[i] + nbytes]
                    if pline:
                        raw = zlib.decompress(raw)
                    assert len(raw) == 8 * cdim
                    v = np.frombuffer(raw, dtype="<f8")
                    n = min(cdim, dims[0] - off)
                    out[off:off + n] = v[:n]
                    assert np.all(v[n:] == 0)
                    covered.append((off, n))
            return keys[-1][2]
        walk(bt, 0)
        assert sum(n for _, n in covered) == dims[0]
        return out, cdim

    @staticmethod
    def _datatype(m):
        cls, ver = m[0] & 0xF, m[0] >> 4
        size = struct.unpack_from("<I", m, 4)[0]
        assert ver == 1
        if cls == 1:
            assert m[1] & 1 == 0 and size == 8
            off, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", m, 8)
            assert (off, prec, eloc, esz, mloc, msz, bias) == (0, 64, 52, 11, 0, 52, 1023) and m[2] == 63
            return ("f64", size)
        if cls == 9:  # variable-length: string type, null-terminated ASCII, base type = one character
            assert m[1] & 0xF == 1 and (m[1] >> 4) == 0 and m[2] & 0xF == 0 and size == 16
            base = m[8:]
            assert base[0] == 0x13 and struct.unpack_from("<I", base, 4)[0] == 1
            return ("vstr", size)
        assert cls == 3
        return ("str", size)

    def _object(self, addr):
        msgs = self._messages(addr)
        types = [t for t, _ in msgs]
        if 0x11 in types:
            bt, hp = struct.unpack_from("<QQ", dict(msgs)[0x11], 0)
            return self._group(bt, hp)
        d = dict((t, m) for t, m in msgs if t != 0x0C)
        dims, maxdims = self._dataspace(d[0x01], True)
        kind, size = self._datatype(d[0x03])
        lay = d[0x08]
        assert kind == "f64" and lay[0] == 3 and lay[1] in (1, 2)
        storage = {"layout": "contiguous", "maxdims": maxdims}
        if lay[1] == 1:
            a, nbytes = struct.unpack_from("<QQ", lay, 2)
            n = int(np.prod(dims)) if dims else 1
            assert nbytes == 8 * n and maxdims is None and 0x0B not in d
            data = np.frombuffer(self.b, dtype="<f8", count=n, offset=a).reshape(dims)
        else:
            pline = None
            if 0x0B in d:  # filter pipeline: exactly {deflate, level}
                fm = d[0x0B]
                assert fm[0] == 1 and fm[1] == 1
                fid, nlen, flags, ncd = struct.unpack_from("<HHHH", fm, 8)
                assert fid == 1 and nlen % 8 == 0 and ncd == 1 and fm[16:16 + nlen].split(b"\0")[0] == b"deflate"
                pline = struct.unpack_from("<I", fm, 16 + nlen)[0]
            assert maxdims == [UNDEF] and d[0x05][1] == 3
            data, cdim = self._chunked(lay, dims, pline)
            storage = {"layout": "chunked", "chunk": cdim, "deflate": pline, "maxdims": maxdims}
        attrs = {}
        for t, m in msgs:
            if t != 0x0C:
                continue
            assert m[0] == 1
            nsz, dtsz, dssz = struct.unpack_from("<HHH", m, 2)
            p = 8
            name = m[p:p + nsz].split(b"\0")[0].decode(); p += (nsz + 7) & ~7
            kind, size = self._datatype(m[p:p + dtsz]); p += (dtsz + 7) & ~7
            assert self._dataspace(m[p:p + dssz]) == []; p += (dssz + 7) & ~7
            if kind == "vstr":
                ln, gaddr, gidx = struct.unpack_from("<IQI", m, p)
                val = self._global_heap_object(gaddr, gidx)
                assert len(val) == ln and b"\0" not in val
                attrs[name] = val.decode()
                storage.setdefault("attr_kinds", set()).add("vlen")
            else:
                assert kind == "str"
                attrs[name] = m[p:p + size].split(b"\0")[0].decode()
                storage.setdefault("attr_kinds", set()).add("fixed")
        return {"data": data, "attrs": attrs, "storage": storage}
