"""TEST INFRASTRUCTURE: a minimal structural HDF5 reader (version-0 superblock, v1 object headers, symbol-table
groups, contiguous datasets, string attributes), written from the HDF5 File Format Specification independently of the
emitter's code paths: it walks the file the way libhdf5 does (superblock -> root symbol-table entry -> B-tree ->
SNOD -> object headers).  h5py / libhdf5 are not available offline; this reader is what checks the emitted files."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Min:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "bad signature"
        assert b[8] == 0 and b[13] == 8 and b[14] == 8, "superblock v0 with 8-byte offsets/lengths expected"
        self.k_leaf, self.k_int = struct.unpack_from("<HH", b, 16)
        base, freesp, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        assert base == 0 and freesp == UNDEF and drv == UNDEF and eof == len(b), (base, freesp, eof, len(b))
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
    def _dataspace(m):
        assert m[0] == 1
        rank, flags = m[1], m[2]
        return list(struct.unpack_from("<%dQ" % rank, m, 8))

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
        assert cls == 3
        return ("str", size)

    def _object(self, addr):
        msgs = self._messages(addr)
        types = [t for t, _ in msgs]
        if 0x11 in types:
            bt, hp = struct.unpack_from("<QQ", dict(msgs)[0x11], 0)
            return self._group(bt, hp)
        d = dict((t, m) for t, m in msgs if t != 0x0C)
        dims = self._dataspace(d[0x01])
        kind, size = self._datatype(d[0x03])
        lay = d[0x08]
        assert kind == "f64" and lay[0] == 3 and lay[1] == 1
        a, nbytes = struct.unpack_from("<QQ", lay, 2)
        n = int(np.prod(dims)) if dims else 1
        assert nbytes == 8 * n
        data = np.frombuffer(self.b, dtype="<f8", count=n, offset=a).reshape(dims)
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
            assert kind == "str"
            attrs[name] = m[p:p + size].split(b"\0")[0].decode()
        return {"data": data, "attrs": attrs}
