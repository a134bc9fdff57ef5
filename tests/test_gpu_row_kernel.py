"""Single-row FFMA2 kernel (csrc/afp_fir_row.cuh: bulk-copy / cp.async loaders, tap-pair FFMA2, tail chunk,
warp-transposed stores) against the float64 oracle, the Philox/quantiser oracle (bit-exact) and the row-pair
kernel it replaces on aligned single-table stages.  Tolerance: 1e-5 relative / -100 dBFS (conftest)."""
import os

import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402
from oracle import dither_quant as dq  # noqa: E402

ROW, PAIR = 2, 0


@pytest.fixture(autouse=True)
def _ffma2_kernels_only(monkeypatch):
    """These tests are about the two packed-FFMA2 kernels: U = 1 filters of 176 taps and more would otherwise run on the
    split-FP16 tensor kernel whenever the block is a multiple of 512 frames (tests/test_gpu_tensor2.py covers that)."""
    monkeypatch.setenv("AFP_NO_TC2", "1")


def noise(S, n, seed=99, amp=0.1):
    return np.stack([np.clip(amp * np.random.default_rng(seed + s).standard_normal(n), -1, 1)
                     for s in range(S)]).astype(np.float32)


def taps(N, U, seed=5):
    """Arbitrary (asymmetric) taps with a decaying envelope: every tap index matters in the comparison."""
    r = np.random.default_rng(seed + N)
    return U * r.standard_normal(N) * np.exp(-np.arange(N) / (0.4 * N + 3)) / np.sqrt(N)


def run(eng, x, sizes):
    xd = torch.from_numpy(x).cuda()
    outs, kernels, pos = [], [], 0
    for B in sizes:
        outs.append(eng.process(xd[:, pos:pos + B].contiguous()).cpu().numpy())
        kernels.append(eng.stage_info()[-1]["kernel"])
        pos += B
    return np.concatenate(outs, axis=1), kernels


def oracle(h, x, U):
    y = (rc.streaming_fir(h, x) if U == 1 else rc.polyphase_upsample_whole(h, x, U)).astype(np.float32)
    # upfirdn stops at the last non-zero product: fewer than U taps leaves the final phases of the last
    # sample out, which the streaming kernel writes as the zeros they are
    return np.concatenate([y, np.zeros(len(x) * U - len(y), dtype=np.float32)])


# tap counts chosen around every tile / tail boundary: U=1 tiles of 16 with 4-slot tails (N+1 slots),
# U=2 tiles of 8 (<= 64 taps) / 16 with 2-slot tails, U=3 tiles of 8 (phase-major form), U=4 tiles of 4 (<= 64 taps) / 8
@pytest.mark.parametrize("U,N", [(1, 1), (1, 2), (1, 3), (1, 7), (1, 15), (1, 16), (1, 47), (1, 51), (1, 52), (1, 63),
                                 (1, 64), (1, 101), (1, 251), (1, 385),
                                 (2, 2), (2, 9), (2, 27), (2, 51), (2, 64), (2, 65), (2, 101), (2, 129), (2, 251), (2, 501),
                                 (3, 2), (3, 5), (3, 33), (3, 129), (3, 251), (3, 1001),
                                 (4, 3), (4, 33), (4, 51), (4, 64), (4, 65), (4, 99), (4, 101), (4, 103), (4, 251), (4, 1001)])
def test_row_kernel_matches_oracle(U, N):
    h = taps(N, U)
    S, sizes = 5, [1024, 1024, 512, 1000, 256, 1024, 4, 96]      # whole warps, ragged tails, tiny blocks
    x = noise(S, sum(sizes))
    with BatchFirEngine(h, n_streams=S, max_frames=1024, upsample=U) as eng:
        y, kernels = run(eng, x, sizes)
    assert kernels == [ROW] * len(sizes), "the single-row kernel should have run"
    for s in range(S):
        assert_parity(y[s], oracle(h, x[s], U), f"U={U} N={N} s={s}")


@pytest.mark.parametrize("U,N,B", [(1, 51, 1024), (2, 51, 1024), (4, 51, 1024), (4, 101, 512), (2, 101, 2048), (1, 251, 4096),
                                   (3, 129, 1024), (3, 34, 512)])
def test_row_and_pair_kernels_agree_bit_for_bit(U, N, B):
    """Both kernels accumulate the taps of an output in the same order with fused multiply-adds, so the
    float results must be identical, not merely within tolerance."""
    h = taps(N, U)
    S = 67                                                       # odd: the row-pair kernel pads its last pair
    x = noise(S, B * 3)
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U) as eng:
        y_row, k_row = run(eng, x, [B] * 3)
    os.environ["AFP_NO_ROW"] = "1"
    try:
        with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U) as eng:
            y_pair, k_pair = run(eng, x, [B] * 3)
    finally:
        del os.environ["AFP_NO_ROW"]
    assert k_row == [ROW] * 3 and k_pair == [PAIR] * 3
    np.testing.assert_array_equal(y_row, y_pair)


def test_kernels_interchange_on_one_history():
    """Aligned blocks -> single-row kernel, frames % 4 != 0 -> row-pair kernel; the carried history is
    shared, so a stream may alternate."""
    h = taps(101, 1)
    sizes = [512, 37, 1024, 3, 100, 511, 512]
    x = noise(3, sum(sizes))
    with BatchFirEngine(h, n_streams=3, max_frames=1024) as eng:
        y, kernels = run(eng, x, sizes)
    assert kernels == [ROW, PAIR, ROW, PAIR, ROW, PAIR, ROW]
    for s in range(3):
        assert_parity(y[s], oracle(h, x[s], 1), f"s={s}")


@pytest.mark.parametrize("U", [1, 2, 3, 4])
@pytest.mark.parametrize("fmt,mode", [("i24", "tpdf"), ("i16", "tpdf"), ("i24_packed", "tpdf"), ("i24", "rpdf"),
                                      ("f32", "tpdf"), ("i16", "none")])
def test_row_kernel_epilogues_bit_exact(U, fmt, mode):
    """Dither + quantisation from the single-row kernel (direct and warp-transposed stores, every output
    format) == the numpy Philox / quantiser statement applied to the kernel's own float output."""
    h = taps(51, U)
    S, B, seed, sid0 = 6, 1024, 77, 3
    x = noise(S, B * 2 + 200, amp=0.4)
    sizes = [B, B, 200]
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U) as ef:
        yf, _ = run(ef, x, sizes)
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, out_format=fmt, dither=mode, seed=seed,
                        first_stream_id=sid0) as eq:
        yq, kernels = run(eq, x, sizes)
    assert kernels == [ROW] * 3
    fmt_code = {"f32": dq.FMT_F32, "i16": dq.FMT_I16, "i24": dq.FMT_I24_IN_I32, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    mode_code = {"tpdf": dq.DITHER_TPDF, "rpdf": dq.DITHER_RPDF, "none": dq.DITHER_NONE}[mode]
    ref = dq.epilogue(yf, fmt_code, mode_code, seed, np.arange(S) + sid0, 0)
    np.testing.assert_array_equal(yq, ref)


def test_row_kernel_large_batch_linearity():
    """BASELINE-size batch (4096 streams, U=2, 251 min-phase-like taps): linearity y(a+b) = y(a)+y(b) to
    rounding and 64 streams against the oracle."""
    U, N, S, B = 2, 251, 4096, 1024
    h = taps(N, U)
    g = torch.Generator(device="cuda").manual_seed(4)
    a = 0.1 * torch.randn(S, B * 2, device="cuda", generator=g)
    b = 0.1 * torch.randn(S, B * 2, device="cuda", generator=g)

    def go(x):
        with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U) as eng:
            out = [eng.process(x[:, i:i + B].contiguous()) for i in range(0, x.shape[1], B)]
            assert eng.stage_info()[-1]["kernel"] == ROW
        return torch.cat(out, dim=1)

    ya, yb, yab = go(a), go(b), go(a + b)
    assert float((yab - (ya + yb)).abs().max()) < 2e-6
    an = a[:64].cpu().numpy()
    for s in range(64):
        assert_parity(ya[s].cpu().numpy(), oracle(h, an[s], U), f"stream {s}")


@pytest.mark.parametrize("U,N,fmt", [(4, 51, "f32"), (4, 101, "i16"), (2, 51, "i24"), (2, 129, "f32"), (1, 51, "i24_packed"),
                                     (1, 251, "f32"), (3, 51, "f32"), (3, 129, "i24_packed"), (3, 20, "i16")])
def test_row_kernel_writes_nothing_outside_its_rows(U, N, fmt):
    """Guard bands: the caller's output rows are longer than the block needs (and the pitch is not a
    multiple of a warp's store width); everything past the valid samples of every row must stay untouched,
    for whole-warp blocks (transposed stores) and ragged ones (direct stores), and so must the samples of a
    guard stream row below the batch."""
    S = 7
    h = taps(N, U)
    dt = {"f32": torch.float32, "i16": torch.int16, "i24": torch.int32, "i24_packed": torch.uint8}[fmt]
    k = 3 if fmt == "i24_packed" else 1
    sentinel = 77
    with BatchFirEngine(h, n_streams=S, max_frames=1024, upsample=U, out_format=fmt,
                        dither="tpdf" if fmt != "f32" else "none", seed=9) as eng, \
            BatchFirEngine(h, n_streams=S, max_frames=1024, upsample=U, out_format=fmt,
                           dither="tpdf" if fmt != "f32" else "none", seed=9) as ref:
        for frames in (1024, 1000, 516, 4):
            x = torch.from_numpy(noise(S, frames, seed=frames)).cuda()
            cols = frames * U * k
            pad = 48 if fmt != "i24_packed" else 48 * 3                  # keeps rows 16-byte aligned
            big = torch.full((S + 1, cols + pad), sentinel, dtype=dt, device="cuda")
            eng.process(x, big[:S])
            want = ref.process(x)
            assert eng.stage_info()[-1]["kernel"] == ROW
            assert torch.equal(big[:S, :cols], want), f"frames={frames}"
            assert bool((big[:S, cols:] == sentinel).all()), f"frames={frames}: wrote past the end of a row"
            assert bool((big[S] == sentinel).all()), f"frames={frames}: wrote below the batch"
