"""Minimal JLD2 reader and writer for the reference's result artifacts.

JLD2 files are HDF5 (superblock v2 at byte 512, version-2 object headers, link messages in the groups,
contiguous or compact little-endian datasets, Julia arrays stored column-major with the HDF5 dimensions
reversed, `Vector{Vector{T}}` as arrays of 8-byte object references).  h5py is not available in the build
image, so this walks exactly that subset: enough for `source_data/*.jld2` and `suppression/results/*.jld2`
(weights, betas, indices, data tensors).  NamedTuple datasets (compound types with committed datatypes, as in
`data/ohashi.jld2`) are outside the subset; `tests/golden/make_fixtures.py` reads those at fixed offsets.
`save` writes the same subset back (Int64 / Float64 scalars and arrays, plain datasets without Julia type attributes,
the way JLD2 itself lays them out: 512-byte text header, superblock v2, v2 object headers with Jenkins lookup3
checksums) — the result files of `02-conditional.jl:44-50` / `07-covariate-inclusion.jl:59-65` with the
`Vector{Vector{Float64}}` fields stored as matrices (one vector per column).
Host-side I/O only — nothing here touches the GPU path.
"""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF
_M32 = 0xFFFFFFFF


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & _M32


def lookup3(data, initval=0):
    """Bob Jenkins' lookup3 `hashlittle` — the checksum HDF5 puts on superblock v2 and version-2 object headers."""
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & _M32
    i = 0
    while n > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & _M32
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & _M32
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 4); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 6); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 8); b = (b + a) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 16); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 19); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 4); b = (b + a) & _M32
        i += 12
        n -= 12
    if n == 0:
        return c
    tail = bytes(data[i:]) + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & _M32
    b = (b + int.from_bytes(tail[4:8], "little")) & _M32
    c = (c + int.from_bytes(tail[8:12], "little")) & _M32
    c ^= b; c = (c - _rot(b, 14)) & _M32
    a ^= c; a = (a - _rot(c, 11)) & _M32
    b ^= a; b = (b - _rot(a, 25)) & _M32
    c ^= b; c = (c - _rot(b, 16)) & _M32
    a ^= c; a = (a - _rot(c, 4)) & _M32
    b ^= a; b = (b - _rot(a, 14)) & _M32
    c ^= b; c = (c - _rot(b, 24)) & _M32
    return c


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
        self.superblock_checksum_ok = lookup3(sb[:44]) == struct.unpack("<I", sb[44:48])[0]

    # ---- object headers --------------------------------------------------------------------
    def _messages(self, addr, verify=False):
        """Yield (type, body) for every message of the version-2 object header at `addr`; verify=True also checks the
        lookup3 checksum that closes every header block."""
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
        chunks = [(p, size, self.base + addr)]
        track = bool(flags & 0x04)
        while chunks:
            start, length, block = chunks.pop(0)
            if verify and lookup3(b[block:start + length]) != struct.unpack("<I", b[start + length:start + length + 4])[0]:
                raise ValueError(f"object header checksum mismatch at {block - self.base}")
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
                    chunks.append((o + 4, ln - 8, o))               # minus signature and checksum
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

    def verify(self):
        """Check the superblock checksum and the checksum of every object header reachable from the root group."""
        if not self.superblock_checksum_ok:
            raise ValueError("superblock checksum mismatch")
        todo, seen = [self.root], set()
        while todo:
            a = todo.pop()
            if a in seen:
                continue
            seen.add(a)
            is_group = False
            for mtype, _ in self._messages(a, verify=True):
                is_group |= mtype == 0x02
            if is_group:
                todo.extend(self._links(a).values())
        return len(seen)

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


# ---- writer ------------------------------------------------------------------------------------
_F64_TYPE = bytes.fromhex("31203f000800000000004000340b0034ff030000")    # IEEE double, little-endian (class 1, version 3)
_I64_TYPE = bytes.fromhex("300800000800000000004000")                    # signed 64-bit integer (class 0, version 3)


def _message(mtype, body):
    return struct.pack("<BHB", mtype, len(body), 0) + body


def _object_header(messages):
    """Version-2 object header holding `messages` in one block, closed by its lookup3 checksum."""
    payload = b"".join(messages)
    if len(payload) < 256:
        head = b"OHDR\x02\x00" + struct.pack("<B", len(payload))
    else:
        head = b"OHDR\x02\x01" + struct.pack("<H", len(payload))
    block = head + payload
    return block + struct.pack("<I", lookup3(block))


def _dataset_header_size(rank, dtype_msg, contiguous):
    n = 4 + 2 + 4 + (4 + 8 * rank) + 4 + len(dtype_msg) + 4 + (18 if contiguous else 0)
    return n


def _encode(value):
    """-> (rank-reversed dims, datatype message body, raw little-endian bytes in Julia's column-major order)."""
    if isinstance(value, (list, tuple)) and value and all(np.ndim(v) == 1 for v in value):
        value = np.stack([np.asarray(v) for v in value], axis=1)         # Vector{Vector} -> one vector per column
    if isinstance(value, (bool, np.bool_)):
        raise TypeError("Bool is outside the supported subset")
    a = np.asarray(value)
    if a.dtype.kind in "iu":
        a, t = a.astype("<i8"), _I64_TYPE
    elif a.dtype.kind == "f":
        a, t = a.astype("<f8"), _F64_TYPE
    else:
        raise TypeError(f"unsupported value type {a.dtype} (Int64 / Float64 scalars and arrays only)")
    return tuple(reversed(a.shape)), t, np.ascontiguousarray(a.T).tobytes()


def save(path, datasets, creator="conditional_ude_b200"):
    """Write `datasets` (name -> int / float / ndarray / list of equal-length vectors) as a JLD2 file."""
    base = 512
    body = bytearray(48)                                                  # superblock, filled in at the end
    links = []
    for name, value in datasets.items():
        dims, tmsg, raw = _encode(value)
        rank = len(dims)
        addr = len(body)
        space = struct.pack("<BBBB", 2, rank, 0, 1 if rank else 0) + b"".join(struct.pack("<Q", d) for d in dims)
        msgs = [_message(0x05, b"\x03\x09"), _message(0x01, space), _message(0x03, tmsg)]
        if rank == 0:
            msgs.append(_message(0x08, struct.pack("<BBH", 4, 0, len(raw)) + raw))          # compact
            body += _object_header(msgs)
        else:
            probe = _object_header(msgs + [_message(0x08, struct.pack("<BBQQ", 4, 1, 0, len(raw)))])
            data_addr = (addr + len(probe) + 7) & ~7
            body += _object_header(msgs + [_message(0x08, struct.pack("<BBQQ", 4, 1, data_addr if raw else _UNDEF, len(raw)))])
            body += b"\0" * (data_addr - len(body)) + raw
        links.append((name, addr))
    root = len(body)
    msgs = [_message(0x02, b"\x00\x00" + b"\xff" * 16), _message(0x0A, b"\x00\x00")]
    for name, addr in links:
        nb = name.encode("utf-8")
        if len(nb) > 255:
            raise ValueError("dataset name too long")
        msgs.append(_message(0x06, b"\x01\x10\x01" + struct.pack("<B", len(nb)) + nb + struct.pack("<Q", addr)))
    body += _object_header(msgs)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", base, _UNDEF, base + len(body), root)
    body[:48] = sb + struct.pack("<I", lookup3(sb))
    text = b"HDF5-based Julia Data Format, version 0.2.0\x00 (" + creator.encode("ascii") + b")\x00"
    with open(path, "wb") as f:
        f.write(text.ljust(base, b"\0") + bytes(body))


def save_neural_parameters(path, neural_network_parameters, betas=None, best_model_index=None, width=4, depth=2):
    """The result file of c-peptide/02-conditional.jl:44-50 (and 07-covariate-inclusion.jl:59-65): width, depth,
    parameters, betas, best_model_index (1-based, as the reference stores it).  `parameters` / `betas` hold one model
    per column; accepts the list of OptimizationSolution that `train` returns."""
    if betas is None and all(hasattr(s, "u") for s in neural_network_parameters):
        sols = neural_network_parameters
        neural_network_parameters = [np.asarray(s.u.neural, dtype=np.float64) for s in sols]
        betas = [np.asarray(s.u.conditional, dtype=np.float64) for s in sols]
    out = {"width": int(width), "depth": int(depth), "parameters": list(neural_network_parameters)}
    if betas is not None:
        out["betas"] = list(betas)
    if best_model_index is not None:
        out["best_model_index"] = int(best_model_index)
    save(path, out)


def load_neural_parameters(path):
    """-> dict(width, depth, parameters [n_models x P], betas [n_models x N] or None, best_model_index or None) from a
    file written by the reference (vectors of vectors) or by `save_neural_parameters` (one model per column)."""
    d = load(path)
    def rows(v):
        if v is None:
            return None
        return np.stack(v) if isinstance(v, list) else (np.asarray(v).T.copy() if np.ndim(v) == 2 else np.asarray(v)[None])
    return {"width": d.get("width"), "depth": d.get("depth"), "parameters": rows(d.get("parameters")),
            "betas": rows(d.get("betas")), "best_model_index": d.get("best_model_index")}
