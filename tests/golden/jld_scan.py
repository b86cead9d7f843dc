"""Minimal reader for the reference's JLD (HDF5) test fixtures, without libhdf5.

Test infrastructure only.  Reads `test/data/O3.jld` and `test/data/linalg.jld` of the
reference (written by Julia 0.7 / JLD 0.1.2): 512-byte user block, superblock v0, v1 object
headers, contiguous/compact uncompressed datasets.  Link messages are found by a byte scan
(the root group uses dense link storage in a fractal heap) and kept when the address they
carry parses as a v1 object header.

Julia arrays are column-major and HDF5 dims are the reverse of Julia's, so
``np.frombuffer(...).reshape(hdf5_dims)`` is the C-order view of the transposed array;
`load()` returns arrays indexed like Julia (``a[i-1, j-1, k-1]``).
"""
import re
import struct
import numpy as np

BASE = 512


class JLDFile:
    def __init__(self, path):
        self.d = open(path, "rb").read()
        assert self.d[BASE:BASE + 8] == b"\x89HDF\r\n\x1a\n", "not a JLD/HDF5 file with 512B userblock"
        self.objs = {}
        self._scan_links()

    # ---- object headers -------------------------------------------------------------
    def _messages(self, addr):
        d = self.d
        o = BASE + addr
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", d, o)
        if ver != 1:
            raise ValueError("not a v1 object header")
        blocks = [(o + 16, hsize)]
        msgs = []
        bi = 0
        while bi < len(blocks) and len(msgs) < nmsg:
            p, ln = blocks[bi]
            end = p + ln
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _mflags = struct.unpack_from("<HHB", d, p)
                body = d[p + 8:p + 8 + msize]
                msgs.append((mtype, body))
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((BASE + coff, clen))
                p += 8 + msize
            bi += 1
        return msgs

    def _is_header(self, addr):
        o = BASE + addr
        if addr <= 0 or o + 16 > len(self.d):
            return False
        ver, rsv, nmsg, refc, hsize = struct.unpack_from("<BBHII", self.d, o)
        return ver == 1 and rsv == 0 and 0 < nmsg < 64 and 0 < refc < 1000 and 0 < hsize <= 65536

    def _scan_links(self):
        d = self.d
        # link message v1, flags with 1-byte name length (bits0-1 == 0); optional fields by flag bits
        for m in re.finditer(rb"\x01([\x00-\x1f])", d):
            flags = m.group(1)[0]
            if flags & 3:
                continue
            p = m.end()
            if flags & 0x08:
                if d[p] != 0:
                    continue  # hard links only
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            nlen = d[p]
            p += 1
            name = d[p:p + nlen]
            if nlen == 0 or not re.fullmatch(rb"[A-Za-z0-9_]+", name):
                continue
            p += nlen
            if p + 8 > len(d):
                continue
            (addr,) = struct.unpack_from("<Q", d, p)
            if self._is_header(addr):
                try:
                    self._messages(addr)
                except Exception:
                    continue
                self.objs.setdefault(name.decode(), set()).add(addr)

    # ---- datasets -------------------------------------------------------------------
    def _dataset(self, addr):
        dims, raw, dtsize = None, None, None
        for mtype, body in self._messages(addr):
            if mtype == 0x1:  # dataspace
                ver, rank, fl = body[0], body[1], body[2]
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from("<%dQ" % rank, body, off) if rank else ()
            elif mtype == 0x3:  # datatype
                dtsize = struct.unpack_from("<I", body, 4)[0]
            elif mtype == 0x8:  # layout v3
                assert body[0] == 3
                cls = body[1]
                if cls == 1:
                    a, sz = struct.unpack_from("<QQ", body, 2)
                    raw = self.d[BASE + a:BASE + a + sz] if a != 0xFFFFFFFFFFFFFFFF else b""
                elif cls == 0:
                    (sz,) = struct.unpack_from("<H", body, 2)
                    raw = body[4:4 + sz]
                else:
                    raise NotImplementedError("chunked layout")
        return dims, raw, dtsize

    def names(self):
        return sorted(k for k in self.objs if not re.fullmatch(r"\d{8}", k) and not k.startswith("_"))

    def _addr(self, name):
        cands = self.objs[name]
        good = []
        for a in cands:
            try:
                dims, raw, _ = self._dataset(a)
                if dims is not None and raw is not None:
                    good.append(a)
            except Exception:
                pass
        assert good, name
        return good[0]

    def _array(self, addr, kind=None):
        dims, raw, _ = self._dataset(addr)
        cnt = int(np.prod(dims)) if dims else 1
        esz = len(raw) // max(cnt, 1)
        if kind == "int" or (kind is None and False):
            a = np.frombuffer(raw, "<i8", cnt)
        elif esz == 16:
            a = np.frombuffer(raw, "<c16", cnt)
        elif esz == 8:
            a = np.frombuffer(raw, "<f8" if kind != "int" else "<i8", cnt)
        else:
            raise NotImplementedError((dims, len(raw)))
        a = a.reshape(dims) if dims else a.reshape(())
        return np.ascontiguousarray(a.T) if a.ndim > 1 else a.copy()

    def load(self, name, kind=None):
        """Dense array (Julia indexing order) or, for a 40-byte scalar compound, a CSC triple."""
        addr = self._addr(name)
        dims, raw, _ = self._dataset(addr)
        cnt = int(np.prod(dims)) if dims else 1
        if len(raw) == 40 and cnt == 1:  # SparseMatrixCSC{Tv,Int64}
            m, n, r1, r2, r3 = struct.unpack("<qqQQQ", raw)
            colptr = self._array(r1, "int")
            rowval = self._array(r2, "int")
            nzval = self._array(r3)
            return {"m": m, "n": n, "colptr": colptr, "rowval": rowval, "nzval": nzval}
        return self._array(addr, kind)


def csc_to_dense(c):
    a = np.zeros((c["m"], c["n"]), dtype=c["nzval"].dtype)
    for j in range(c["n"]):
        for k in range(c["colptr"][j] - 1, c["colptr"][j + 1] - 1):
            a[c["rowval"][k] - 1, j] = c["nzval"][k]
    return a
