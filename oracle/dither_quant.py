"""ORACLE - test infrastructure, not product code.

Numpy statement of the dither + quantise epilogue (kernel 5).  The reference never quantises and
its "TPDF" dither is really RPDF float noise from an unseeded RNG (stream_process.py:61-64,
SURVEY.md F6), so this stage is BUILD-DEFINED; the definition lives here and in
``csrc/afp_philox.h`` and the GPU must match it bit for bit:

    counter-based RNG : Philox4x32-10 (Salmon et al., SC'11; Random123 constants)
        key     = (seed & 0xffffffff, seed >> 32)
        counter = (c & 0xffffffff, c >> 32, stream_id, 0)   with c = absolute_sample_index >> 2
        sample i uses output word i & 3 (four samples per call)
    a, b    = word >> 16, word & 0xffff          (two 16-bit integers)
    TPDF    : d = float32(a + b - (2**16 - 1)) * 2**-16          in (-1, 1) LSB, triangular
    RPDF    : d = float32(2*a + 1 - 2**16) * 2**-17              in (-0.5, 0.5) LSB, uniform
              (the reference's amplitude: (u-0.5)*2/2**bits == +-0.5 LSB at LSB = 2**-(bits-1))
    integer : q = clip(rint(fl32(fl32(y * 2**(bits-1)) + d)), -2**(bits-1), 2**(bits-1) - 1)
              rint = round-half-even, every operation rounded to float32 (no fused multiply-add)
    float   : y_out = fl32(y + fl32(d * 2**-(bits-1)))           (dither only, reference style)

"absolute_sample_index" counts output samples of one stream since the last reset, so the result
does not depend on how the signal is cut into blocks.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

DITHER_NONE, DITHER_TPDF, DITHER_RPDF = 0, 1, 2
FMT_F32, FMT_I16, FMT_I24_IN_I32, FMT_I24_PACKED = 0, 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcastable uint32-valued arrays; returns four
    uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _MASK for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def dither_lsb(mode, seed, stream_ids, sample_index):
    """Dither value in LSB units (float32) for each (stream, absolute sample index) pair.
    ``stream_ids`` and ``sample_index`` broadcast against each other."""
    sample_index = np.asarray(sample_index, dtype=np.uint64)
    stream_ids = np.asarray(stream_ids, dtype=np.uint64)
    sample_index, stream_ids = np.broadcast_arrays(sample_index, stream_ids)
    if mode == DITHER_NONE:
        return np.zeros(sample_index.shape, dtype=np.float32)
    ctr = sample_index >> np.uint64(2)
    w = philox4x32_10(ctr & _MASK, ctr >> np.uint64(32), stream_ids, 0,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    word = np.choose((sample_index & np.uint64(3)).astype(np.int64), list(w)).astype(np.int64)
    a, b = word >> 16, word & 0xFFFF
    if mode == DITHER_TPDF:
        t = a + b - ((1 << 16) - 1)
        return (t.astype(np.float32) * np.float32(2.0 ** -16)).astype(np.float32)
    if mode == DITHER_RPDF:
        t = 2 * a + 1 - (1 << 16)
        return (t.astype(np.float32) * np.float32(2.0 ** -17)).astype(np.float32)
    raise ValueError("unknown dither mode")


def quantise(y, bits, d_lsb=None):
    """float32 samples -> int32 codes of a ``bits``-bit signed format (see module docstring)."""
    y = np.asarray(y, dtype=np.float32)
    scale = np.float32(2.0 ** (bits - 1))
    v = (y * scale).astype(np.float32)
    if d_lsb is not None:
        v = (v + np.asarray(d_lsb, dtype=np.float32)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        q = np.rint(v)
    q = np.where(np.isnan(q), 0.0, q)
    q = np.clip(q, -float(scale), float(scale) - 1.0)
    return q.astype(np.int32)


def dither_float(y, bits, d_lsb):
    """Reference-style float dither: noise added, no quantisation (stream_process.py:63-64)."""
    y = np.asarray(y, dtype=np.float32)
    lsb = np.float32(2.0 ** -(bits - 1))
    return (y + (np.asarray(d_lsb, dtype=np.float32) * lsb).astype(np.float32)).astype(np.float32)


def pack_i24(q):
    """int32 codes -> little-endian packed 3-byte samples (uint8[..., 3*n])."""
    q = np.asarray(q, dtype=np.int32)
    u = q.astype(np.uint32)
    out = np.empty(q.shape + (3,), dtype=np.uint8)
    out[..., 0] = u & 0xFF
    out[..., 1] = (u >> 8) & 0xFF
    out[..., 2] = (u >> 16) & 0xFF
    return out.reshape(q.shape[:-1] + (q.shape[-1] * 3,))


def unpack_i24(b):
    b = np.asarray(b, dtype=np.uint8)
    b = b.reshape(b.shape[:-1] + (b.shape[-1] // 3, 3)).astype(np.int32)
    v = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)


def epilogue(y, fmt, dither_mode, seed, stream_ids, sample_base):
    """Full epilogue over ``y[S, L]`` float32: returns what the device writes for ``fmt``."""
    y = np.asarray(y, dtype=np.float32)
    S, L = y.shape
    idx = np.uint64(sample_base) + np.arange(L, dtype=np.uint64)[None, :]
    sid = np.asarray(stream_ids, dtype=np.uint64).reshape(S, 1)
    d = dither_lsb(dither_mode, seed, sid, idx) if dither_mode != DITHER_NONE else None
    if fmt == FMT_F32:
        return y if d is None else dither_float(y, 24, d)
    if fmt == FMT_I16:
        return quantise(y, 16, d).astype(np.int16)
    q = quantise(y, 24, d)
    if fmt == FMT_I24_IN_I32:
        return q
    if fmt == FMT_I24_PACKED:
        return pack_i24(q)
    raise ValueError("unknown output format")
