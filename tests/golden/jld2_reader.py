"""Minimal reader for the reference's golden file test/data/gen_test_results.jld2.

Test infrastructure only.  The file is a plain (uncompressed, unchunked) JLD2 / HDF5
container written by /root/reference/test/gen_tst_results.jl:237 (Julia 1.10.2); no
HDF5 library is available in this image, so the few HDF5 structures it uses are parsed
by hand: superblock v2, version-2 object headers ("OHDR"), link messages (0x06),
dataspace (0x01), datatype (0x03, only its size is needed) and data-layout v4 (0x08,
compact or contiguous).  All arrays are Julia column-major Float64/Int64; object
references are 8-byte addresses relative to the base address (512).
"""
import struct
import numpy as np


class JLD2File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        sb = b.find(b"\x89HDF\r\n\x1a\n")
        assert sb >= 0 and b[sb + 8] == 2, "expected an HDF5 v2 superblock"
        self.base, _ext, _eof, root = struct.unpack_from("<QQQQ", b, sb + 12)
        self.base = sb if self.base == 0 else self.base
        self.root = root

    # -- object header ---------------------------------------------------------
    def messages(self, addr):
        """Yield (type, payload-bytes) for every message of the v2 object header at addr."""
        b = self.buf
        o = addr + self.base
        assert b[o:o + 4] == b"OHDR", (addr, b[o:o + 4])
        flags = b[o + 5]
        p = o + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        chunk = int.from_bytes(b[p:p + nsz], "little")
        p += nsz
        yield from self._chunk(p, p + chunk, flags)

    def _chunk(self, p, end, flags):
        b = self.buf
        while p + 4 <= end:
            mtype = b[p]
            msize = struct.unpack_from("<H", b, p + 1)[0]
            p += 4
            if flags & 0x04:
                p += 2
            body = b[p:p + msize]
            p += msize
            if mtype == 0x10:  # continuation: (address, length) of an "OCHK" block
                caddr, clen = struct.unpack_from("<QQ", body, 0)
                co = caddr + self.base
                assert b[co:co + 4] == b"OCHK"
                yield from self._chunk(co + 4, co + clen - 4, flags)
            elif mtype != 0:
                yield mtype, body

    def links(self, addr):
        out = {}
        for mtype, body in self.messages(addr):
            if mtype != 0x06:
                continue
            lflags = body[1]
            p = 2
            if lflags & 0x08:
                p += 1
            if lflags & 0x04:
                p += 8
            if lflags & 0x10:
                p += 1
            nsz = 1 << (lflags & 3)
            nlen = int.from_bytes(body[p:p + nsz], "little")
            p += nsz
            name = body[p:p + nlen].decode("utf8")
            p += nlen
            out[name] = struct.unpack_from("<Q", body, p)[0]
        return out

    def dataset(self, addr):
        """Return (dims in Julia order, element size, raw bytes)."""
        dims, esize, raw = (), None, None
        for mtype, body in self.messages(addr):
            if mtype == 0x01:
                ver, rank, sflags = body[0], body[1], body[2]
                assert ver == 2
                p = 4
                d = struct.unpack_from("<%dQ" % rank, body, p)
                dims = tuple(reversed(d))  # HDF5 stores them reversed w.r.t. Julia
            elif mtype == 0x03:
                esize = struct.unpack_from("<I", body, 4)[0]
            elif mtype == 0x08:
                assert body[0] in (3, 4)
                cls = body[1]
                if cls == 0:
                    sz = struct.unpack_from("<H", body, 2)[0]
                    raw = body[4:4 + sz]
                elif cls == 1:
                    a, sz = struct.unpack_from("<QQ", body, 2)
                    raw = self.buf[a + self.base:a + self.base + sz]
                else:
                    raise NotImplementedError("chunked layout")
        return dims, esize, raw

    def f64(self, addr):
        dims, esize, raw = self.dataset(addr)
        assert esize == 8
        a = np.frombuffer(raw, dtype="<f8")
        return a.reshape(dims, order="F").copy() if dims else a.copy()

    def refs(self, addr):
        dims, esize, raw = self.dataset(addr)
        return list(struct.unpack("<%dQ" % (len(raw) // 8), raw))


STATE_FIELDS = ["tau2", "u", "xi", "gamma", "S", "theta", "Delta", "M", "mu", "lambda", "pi",
                "Sigma_inv", "invC", "mu_t"]


def read_table(jf, addr):
    """A TypedTables.Table is stored as a compact dataset of references in field order."""
    r = jf.refs(addr)
    # Table -> one reference to the NamedTuple of columns, or the column refs directly
    while len(r) == 1:
        r = jf.refs(r[0])
    out = {}
    for name, a in zip(STATE_FIELDS, r):
        out[name] = jf.f64(a)
    return out
