"""CPU checks of the algorithmic identities the AGC and spectrum kernels rest on (csrc/afp_agc.cuh,
csrc/afp_spectrum.cuh), in numpy at the kernels' precision, against the pinned oracles.  They do not run
the kernels (the -m gpu tests do); they pin the index arithmetic and the reformulations so that a change to
either shows up without a GPU."""
import numpy as np
import pytest

from oracle import agc as oagc
from oracle import spectrum as ospec


def level_steps(n, seed):
    r = np.random.default_rng(seed)
    env = np.repeat(10.0 ** r.uniform(-3.5, 0, 8), n // 8)
    return np.clip(env * r.standard_normal(n), -1, 1).astype(np.float32)


def agc_kernel_form(x, target=0.1, window=512, max_gain=10.0, attack=0.01, release=0.1, fast=True):
    """The kernel's formulation: float32 squares, float64 prefix sums, window sum as a difference with the
    np.convolve 'same' centring (left = W // 2, right = (W - 1) // 2), gain = target * sqrt(W) * rsqrt(sum)
    (fast) or target / (sqrt(sum / W) + 1e-10), float32 recurrence in the max / min form."""
    n = len(x)
    sq = (x * x).astype(np.float32)
    E = np.concatenate([[0.0], np.cumsum(sq.astype(np.float64))])
    i = np.arange(n)
    left, right = window // 2, (window - 1) // 2
    s = (E[np.minimum(i + right, n - 1) + 1] - E[np.maximum(i - left, 0)]).astype(np.float32)
    with np.errstate(divide="ignore"):
        if fast:
            d = np.minimum(np.float32(target * np.sqrt(window)) / np.sqrt(s), np.float32(max_gain))
        else:
            d = np.minimum(np.float32(target) / (np.sqrt(s * np.float32(1.0 / window)) + np.float32(1e-10)),
                           np.float32(max_gain))
    d = d.astype(np.float32)
    _, _, a_att, a_rel = oagc.smoothing_constants(window, attack, release)
    aA, cA, aR, cR = (np.float32(v) for v in (a_att, 1 - a_att, a_rel, 1 - a_rel))
    pick = np.maximum if a_att >= a_rel else np.minimum
    g = np.empty(n, dtype=np.float32)
    g[0] = d[0]
    for k in range(1, n):
        g[k] = pick(cA * g[k - 1] + aA * d[k], cR * g[k - 1] + aR * d[k])
    return x * np.clip(g, np.float32(0.1), np.float32(max_gain))


@pytest.mark.parametrize("kw", [dict(), dict(window=101, target=0.2, max_gain=20.0, attack=0.1, release=0.01),
                                dict(window=256, target=0.05, max_gain=4.0, attack=0.02, release=0.3)])
@pytest.mark.parametrize("fast", [True, False])
def test_agc_kernel_formulation_equals_the_reference_loop(kw, fast):
    x = level_steps(1024, seed=len(kw) + 3)
    ref = oagc.apply_agc(x, target_level=kw.get("target", 0.1), window_size=kw.get("window", 512),
                         max_gain=kw.get("max_gain", 10.0), attack=kw.get("attack", 0.01), release=kw.get("release", 0.1))
    y = agc_kernel_form(x, fast=fast, **kw)
    assert np.abs(y - ref).max() <= 1e-5


def test_fast_gain_criterion():
    """target / max_gain >= 1e-4 (the host's condition for the rsqrt form): wherever the gain is not clamped
    the reference's 1e-10 moves it by less than 1e-6 relative."""
    for target, max_gain in [(0.1, 10.0), (1e-3, 10.0), (0.5, 5000.0)]:
        assert target / max_gain >= 1e-4
        rms = np.geomspace(target / max_gain, 1.0, 200)         # unclamped region
        exact, fast = target / (rms + 1e-10), target / rms
        assert np.abs(fast / exact - 1).max() < 1.1e-6


def spectrum_kernel_form(x):
    """afp_spectrum.cuh in numpy complex64: real-input packing, bit-reversed load, one plain radix-2 stage if
    log2(n/2) is odd, then two stages per pass, unpacking, dB through log2."""
    n = len(x); m = n // 2; L = n.bit_length() - 2
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / (n - 1))).astype(np.float32)
    tw_n = np.exp(-2j * np.pi * np.arange(n // 2) / n).astype(np.complex64)
    tw = tw_n[0::2][:m // 2]
    xw = x.astype(np.float32) * win
    X = np.zeros(m, np.complex64)
    for j in range(m):
        X[int(format(j, f"0{L}b")[::-1], 2)] = xw[2 * j] + 1j * xw[2 * j + 1]
    st = 0
    if L & 1:
        j = np.arange(m // 2); a, b = X[2 * j].copy(), X[2 * j + 1].copy()
        X[2 * j], X[2 * j + 1] = a + b, a - b
        st = 1
    while st < L:
        h = 1 << st
        j = np.arange(m // 4); k = j & (h - 1); i0 = ((j >> st) << (st + 2)) + k
        w1, w2a, w2b = tw[k * (m >> (st + 1))], tw[k * (m >> (st + 2))], tw[(k + h) * (m >> (st + 2))]
        a, b, c, d = X[i0].copy(), X[i0 + h].copy(), X[i0 + 2 * h].copy(), X[i0 + 3 * h].copy()
        tb, td = b * w1, d * w1
        a1, b1, c1, d1 = a + tb, a - tb, c + td, c - td
        tc, te = c1 * w2a, d1 * w2b
        X[i0], X[i0 + 2 * h], X[i0 + h], X[i0 + 3 * h] = a1 + tc, a1 - tc, b1 + te, b1 - te
        st += 2
    k = np.arange(m + 1)
    zk, zr = X[k % m], np.conj(X[(m - k) % m])
    w = np.concatenate([tw_n, [-1.0]]).astype(np.complex64)
    v = 0.5 * (zk + zr) - 0.5j * w * (zk - zr)
    return (6.02059991327962 * np.log2(np.abs(v).astype(np.float32) + np.float32(1e-12))).astype(np.float32)


@pytest.mark.parametrize("n", [64, 128, 256, 1024, 2048, 4096])
def test_spectrum_kernel_formulation_equals_rfft(n):
    r = np.random.default_rng(n)
    x = (0.5 * np.sin(2 * np.pi * 0.123 * np.arange(n)) + 0.01 * r.standard_normal(n)).astype(np.float32)
    got, ref = spectrum_kernel_form(x), ospec.spectrum_db(x)
    lin_g, lin_r = 10.0 ** (got.astype(np.float64) / 20), 10.0 ** (ref / 20)
    assert np.abs(lin_g - lin_r).max() <= 1e-5 * lin_r.max()
    strong = ref >= ref.max() - 60
    assert np.abs(got[strong] - ref[strong]).max() <= 0.01


# ---- split-FP16 tensor kernel (csrc/afp_fir_tc2.cuh): the Toeplitz / Hankel-core form, U = 1 and polyphase ----------------
def _hankel_cores(h, U):
    """tc2_build_cores in numpy: C_q[jj][e] = g[8 q + jj + U e], g[t] = h[t - 63 U]; D + 1 chunks."""
    n, R = len(h), 64
    D = (n - 1 + (R - 1) * U) // (R * U)
    Q = 8 * U * (D + 2) - U
    q, jj, e = np.meshgrid(np.arange(Q), np.arange(8), np.arange(8), indexing="ij")
    k = 8 * q + jj + U * e - U * (R - 1)
    return np.where((k >= 0) & (k < n), h[np.clip(k, 0, n - 1)], 0.0), D


def _tensor_form(h, x, U):
    """What the kernel's MMAs compute, written with the same indices: rows of 64 input samples stored reversed, chunk d joins
    row m - d to row m through the core array only (B descriptor: core 8 U d + U (i / 8) + c / 8, element (c % 8, i % 8))."""
    R = 64
    cores, D = _hankel_cores(h, U)
    rows = len(x) // R
    xr = np.concatenate([np.zeros(D * R), x]).reshape(D + rows, R)[:, ::-1]        # row r of the padded signal, element i = sample R-1-i
    i = np.arange(R)
    c = np.arange(R * U)
    out = np.zeros((rows, R * U))
    for d in range(D + 1):
        qi = (8 * U * d + U * (i[:, None] // 8) + c[None, :] // 8)                    # [i, c] core index
        B = cores[qi, c[None, :] % 8, i[:, None] % 8]                                # W_d[i][c]
        out += xr[D - d: D - d + rows] @ B
    return out.reshape(-1)


@pytest.mark.parametrize("U,n", [(1, 51), (1, 64), (1, 65), (1, 251), (1, 1001), (1, 1025), (2, 96), (2, 251), (2, 1001), (3, 90), (3, 251), (3, 1001), (4, 160),
                                 (4, 251), (4, 1001), (4, 2049)])
def test_hankel_core_form_equals_the_polyphase_fir(U, n):
    """One array of 8 x 8 cores indexed by 8 U d + U (i/8) + c/8 reproduces  out[U m + p] = sum_k h[U k + p] x[m - k]
    (scipy.signal.upfirdn) for every chunk, k-step and column group - the identity behind the resident tap cores, the
    N = 128 walk over [hi | lo] slots and the polyphase form (64 U output columns per 64-sample row)."""
    from scipy import signal
    rng = np.random.default_rng(n * 8 + U)
    h = rng.standard_normal(n) / n
    x = rng.standard_normal(64 * 24)
    ref = signal.upfirdn(h, x, up=U)[: len(x) * U]
    np.testing.assert_allclose(_tensor_form(h, x, U), ref, rtol=0, atol=1e-12)
    cores, D = _hankel_cores(h, U)
    assert D + 1 <= 17 or n > 1025 * U              # the kernel's limit: 17 chunks = 1025 taps per phase
    assert cores.shape[0] == 8 * U * (D + 2) - U    # tc2_n_cores


def test_fp16_split_keeps_22_bits():
    """x = hi + lo 2^-11 with hi = fp16(x), lo = fp16((x - hi) 2^11): three products (hi*hi, hi*lo, lo*hi) leave a relative
    error of ~2^-22 per term - what the kernel's 1e-8 absolute error at audio levels rests on."""
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 4096).astype(np.float32)
    w = (rng.standard_normal(4096) / 64).astype(np.float32)

    def split(v):
        hi = v.astype(np.float16)
        lo = ((v - hi.astype(np.float32)) * np.float32(2048)).astype(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64)

    xh, xl = split(x)
    wh, wl = split(w)
    approx = xh * wh + (xh * wl + xl * wh) / 2048.0
    exact = x.astype(np.float64) * w.astype(np.float64)
    assert np.abs(approx - exact).max() <= 2.0 ** -21 * np.abs(exact).max()
