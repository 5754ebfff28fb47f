"""Minimal JLD2 reader for the reference's result artifacts.

JLD2 files are HDF5 (superblock v2 at byte 512, version-2 object headers, link messages in the groups,
contiguous or compact little-endian datasets, Julia arrays stored column-major with the HDF5 dimensions
reversed, `Vector{Vector{T}}` as arrays of 8-byte object references).  h5py is not available in the build
image, so this walks exactly that subset: enough for `source_data/*.jld2` and `suppression/results/*.jld2`
(weights, betas, indices, data tensors).  NamedTuple datasets (compound types with committed datatypes, as in
`data/ohashi.jld2`) are outside the subset; `tests/golden/make_fixtures.py` reads those at fixed offsets.
Host-side I/O only — nothing here touches the GPU path.
"""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class JLD2File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        self.base = self.b.find(b"\x89HDF\r\n\x1a\n")
        if self.base < 0:
            raise ValueError("not an HDF5/JLD2 file")
        sb = self.b[self.base:self.base + 48]
        if sb[8] != 2 or sb[9] != 8 or sb[10] != 8:
            raise ValueError("unsupported superblock (need version 2, 8-byte offsets)")
        _, _, _, self.root = struct.unpack("<QQQQ", sb[12:44])

    # ---- object headers --------------------------------------------------------------------
    def _messages(self, addr):
        """Yield (type, body) for every message of the version-2 object header at `addr`."""
        b, p = self.b, self.base + addr
        if b[p:p + 4] != b"OHDR" or b[p + 4] != 2:
            raise ValueError(f"no v2 object header at {addr}")
        flags = b[p + 5]
        p += 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        size = int.from_bytes(b[p:p + nsz], "little")
        p += nsz
        chunks = [(p, size)]
        track = bool(flags & 0x04)
        while chunks:
            start, length = chunks.pop(0)
            q, end = start, start + length
            while q + 4 <= end:
                mtype = b[q]
                msize = struct.unpack("<H", b[q + 1:q + 3])[0]
                q += 4 + (2 if track else 0)
                body = b[q:q + msize]
                q += msize
                if mtype == 0x10:                                   # continuation -> OCHK block
                    off, ln = struct.unpack("<QQ", body[:16])
                    o = self.base + off
                    if b[o:o + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    chunks.append((o + 4, ln - 8))                  # minus signature and checksum
                elif mtype != 0:
                    yield mtype, body

    def _links(self, addr):
        out = {}
        for mtype, body in self._messages(addr):
            if mtype != 0x06:
                continue
            flags = body[1]
            p = 2
            ltype = 0
            if flags & 0x08:
                ltype = body[p]; p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            nsz = 1 << (flags & 3)
            nlen = int.from_bytes(body[p:p + nsz], "little")
            p += nsz
            name = body[p:p + nlen].decode("utf-8")
            p += nlen
            if ltype == 0:
                out[name] = struct.unpack("<Q", body[p:p + 8])[0]
        return out

    def keys(self):
        return [k for k in self._links(self.root) if not k.startswith("_")]

    # ---- datasets --------------------------------------------------------------------------
    def _dataset(self, addr):
        dims, dtype, ref, data = (), None, False, None
        for mtype, body in self._messages(addr):
            if mtype == 0x01:                                       # dataspace
                ver, rank = body[0], body[1]
                off = 4 if ver == 2 else 8
                dims = struct.unpack("<" + "Q" * rank, body[off:off + 8 * rank]) if rank else ()
            elif mtype == 0x03:                                     # datatype
                cls, size = body[0] & 0x0F, struct.unpack("<I", body[4:8])[0]
                if cls == 1 and size == 8:
                    dtype = "<f8"
                elif cls == 1 and size == 4:
                    dtype = "<f4"
                elif cls == 0:
                    signed = bool(body[1] & 0x08)
                    dtype = ("<i" if signed else "<u") + str(size)
                elif cls == 7:
                    dtype, ref = "<u8", True
                else:
                    dtype = None
            elif mtype == 0x08:                                     # data layout
                ver, lclass = body[0], body[1]
                if ver not in (3, 4):
                    raise ValueError("unsupported layout message version")
                if lclass == 1:
                    a, n = struct.unpack("<QQ", body[2:18])
                    data = None if a == _UNDEF else self.b[self.base + a:self.base + a + n]
                elif lclass == 0:
                    n = struct.unpack("<H", body[2:4])[0]
                    data = body[4:4 + n]
                else:
                    raise ValueError("chunked datasets are outside the supported subset")
        if dtype is None:
            raise ValueError("unsupported datatype (NamedTuple / string / compound)")
        n = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(data or b"", dtype=dtype, count=n if data else 0).copy()
        if ref:
            return [self._dataset(int(a)) for a in arr]             # Vector{Vector{T}}
        if not dims:
            return arr[0].item() if arr.size else None
        return arr.reshape(dims).T.copy() if len(dims) > 1 else arr  # reversed dims + column-major == transpose

    def __getitem__(self, name):
        links = self._links(self.root)
        if name not in links:
            raise KeyError(name)
        return self._dataset(links[name])


def load(path, names=None):
    """Read the named top-level datasets (default: every readable one) into a dict."""
    f = JLD2File(path)
    out = {}
    for k in (names or f.keys()):
        try:
            out[k] = f[k]
        except ValueError:
            if names:
                raise
    return out
