"""Tensor-core paths (tcgen05 / TMEM Toeplitz FIR: the split-FP16 form with resident Hankel tap cores, kernel 3,
and the 3xTF32 form, kernel 1) parity: same tolerance as the FFMA2 kernel (1e-5 relative / -100 dBFS absolute
against the float64 oracle), bit-exact integer epilogue, and interchangeability with the FFMA2 kernels block by
block (shared history format).  Every test runs twice: with the default selection (kernel 3 for blocks of a
multiple of 512 frames, kernel 1 for 256-frame blocks) and with AFP_NO_TC2=1 (kernel 1 only)."""
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


@pytest.fixture(autouse=True, params=["tc2", "tc1"])
def tensor_variant(request):
    if request.param == "tc1":
        os.environ["AFP_NO_TC2"] = "1"
    yield request.param
    os.environ.pop("AFP_NO_TC2", None)


def TK(frames):
    """kernel id a tensor-eligible stage reports for a block of `frames`"""
    return 3 if frames % 512 == 0 and not os.environ.get("AFP_NO_TC2") else 1


def remez1001():
    return design.create_fir_filter(method="remez", cutoff=20000, numtaps=1001, filter_type="lowpass",
                                    samplerate=176400)


def noise(S, n, seed=1234, amp=0.1):
    return np.stack([np.clip(amp * np.random.default_rng(seed + s).standard_normal(n), -1, 1)
                     for s in range(S)]).astype(np.float32)


def run(eng, x, sizes):
    xd = torch.from_numpy(x).cuda()
    outs, kernels, pos = [], [], 0
    for B in sizes:
        outs.append(eng.process(xd[:, pos:pos + B].contiguous()).cpu().numpy())
        kernels.append(eng.stage_info()[-1]["kernel"])
        pos += B
    return np.concatenate(outs, axis=1), kernels


@pytest.mark.parametrize("N,B,S", [(1001, 512, 37), (1001, 256, 16), (641, 512, 5), (900, 512, 1), (449, 1024, 9),
                                   (401, 512, 3)])
def test_tensor_path_matches_lfilter(N, B, S):
    h = remez1001() if N == 1001 else design.create_fir_filter(cutoff=9000, numtaps=N, window_type="kaiser",
                                                             samplerate=96000)
    x = noise(S, B * 5, amp=0.3)
    with BatchFirEngine(h, n_streams=S, max_frames=B) as eng:
        y, kernels = run(eng, x, [B] * 5)
    assert kernels == [TK(B)] * 5, "the tcgen05 kernel should have run"
    for s in range(S):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"N={N} B={B} s={s}")


def test_1024_frame_blocks_several_row_groups():
    """frames = 1024 -> two (64-sample rows) or four (32-sample rows) 8-row groups per stream; fits in
    shared memory up to ~600 taps."""
    os.environ["AFP_TC_MIN_TAPS"] = "500"
    try:
        h = design.create_fir_filter(cutoff=9000, numtaps=577, window_type="kaiser", samplerate=96000)
        x = noise(7, 1024 * 3, amp=0.3)
        with BatchFirEngine(h, n_streams=7, max_frames=1024) as eng:
            y, kernels = run(eng, x, [1024] * 3)
    finally:
        del os.environ["AFP_TC_MIN_TAPS"]
    assert kernels == [TK(1024)] * 3
    for s in range(7):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"s={s}")


def test_full_scale_signal_stays_in_tolerance():
    """Worst case for split-precision accumulation: a full-scale sine in the pass band."""
    h = remez1001()
    n = 512 * 6
    t = np.arange(n) / 176400.0
    x = np.stack([0.999 * np.sin(2 * np.pi * f * t) for f in (997.0, 5000.0, 15000.0)]).astype(np.float32)
    with BatchFirEngine(h, n_streams=3, max_frames=512) as eng:
        y, kernels = run(eng, x, [512] * 6)
    assert kernels == [TK(512)] * 6
    for s in range(3):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"sine {s}")


def test_kernels_interchange_block_by_block():
    """512-frame blocks run on the tensor cores, other 16-byte-aligned sizes on the single-row FFMA2 kernel
    (2), odd-sized ones on the row-pair FFMA2 kernel (0); all three read and write the same carried
    history, so a stream may alternate freely."""
    h = remez1001()
    sizes = [512, 100, 256, 512, 37, 1024, 512]
    x = noise(4, sum(sizes))
    with BatchFirEngine(h, n_streams=4, max_frames=1024) as eng:
        y, kernels = run(eng, x, sizes)
    assert kernels == [TK(512), 2, 1, TK(512), 0, TK(1024), TK(512)]   # 1024 frames x 1001 taps on kernel 1: a chain of two 512-frame launches
    for s in range(4):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"stream {s}")


def test_tensor_and_ffma_paths_agree():
    h = remez1001()
    x = noise(20, 512 * 3)
    with BatchFirEngine(h, n_streams=20, max_frames=512) as eng:
        y_tc, k_tc = run(eng, x, [512] * 3)
    os.environ["AFP_NO_TC"] = "1"
    try:
        with BatchFirEngine(h, n_streams=20, max_frames=512) as eng:
            y_ff, k_ff = run(eng, x, [512] * 3)
    finally:
        del os.environ["AFP_NO_TC"]
    assert k_tc == [TK(512)] * 3 and k_ff == [2, 2, 2]      # without the tensor path: the single-row FFMA2 kernel
    assert np.abs(y_tc - y_ff).max() < 2e-6


@pytest.mark.parametrize("fmt,mode", [("i24", "tpdf"), ("i16", "tpdf"), ("i24_packed", "tpdf"), ("i24", "none"), ("f32", "rpdf")])
def test_tensor_path_epilogue_bit_exact(fmt, mode):
    h = remez1001()
    S, seed, sid0 = 18, 77, 5
    x = noise(S, 512 * 3, amp=0.5)
    with BatchFirEngine(h, n_streams=S, max_frames=512) as ef:
        yf, kf = run(ef, x, [512] * 3)
    with BatchFirEngine(h, n_streams=S, max_frames=512, out_format=fmt, dither=mode, seed=seed,
                        first_stream_id=sid0) as eq:
        yq, kq = run(eq, x, [512] * 3)
    assert kf == [TK(512)] * 3 and kq == [TK(512)] * 3
    fmt_code = {"f32": dq.FMT_F32, "i16": dq.FMT_I16, "i24": dq.FMT_I24_IN_I32, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    mode_code = {"none": dq.DITHER_NONE, "tpdf": dq.DITHER_TPDF, "rpdf": dq.DITHER_RPDF}[mode]
    ref = dq.epilogue(yf, fmt_code, mode_code, seed, np.arange(S) + sid0, 0)
    if fmt == "i24_packed":      # per-block packing: compare block by block
        nb = 512 * 3
        yq = np.concatenate([yq[:, i:i + nb] for i in range(0, yq.shape[1], nb)], axis=1)
        ref = np.concatenate([dq.epilogue(yf[:, i * 512:(i + 1) * 512], fmt_code, mode_code, seed,
                                          np.arange(S) + sid0, i * 512) for i in range(3)], axis=1)
    np.testing.assert_array_equal(yq, ref)


def test_tensor_path_state_and_host_pipeline():
    h = remez1001()
    S = 40
    x = noise(S, 512 * 4)
    with BatchFirEngine(h, n_streams=S, max_frames=512, out_format="i24", dither="tpdf", seed=3) as e:
        y_all, _ = run(e, x, [512] * 4)
        e.reset()
        y1, _ = run(e, x[:, :1024], [512] * 2)
        blob = e.get_state()
        e.reset()
        e.set_state(blob)
        y2, _ = run(e, x[:, 1024:], [512] * 2)
        np.testing.assert_array_equal(np.concatenate([y1, y2], axis=1), y_all)
        e.reset()
        blocks = np.ascontiguousarray(x.reshape(S, 4, 512).transpose(1, 0, 2))
        yh = e.process_host_blocks(blocks)
        np.testing.assert_array_equal(np.concatenate(list(yh), axis=1), y_all)


def test_long_fused_eq_runs_on_tensor_cores():
    from scipy.signal import lfilter
    fs = 48000
    bank = design.design_eq_bank(701, fs)
    gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3]
    main = design.create_fir_filter(cutoff=16000, numtaps=801, samplerate=fs)
    S, B = 3, 512
    x = noise(S, B * 4, amp=0.05)
    with BatchFirEngine(main, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=gains, eq_mode="fused") as eng:
        y, _ = run(eng, x, [B] * 4)
        assert [st["kernel"] for st in eng.stage_info()] == [TK(512), TK(512)]
    comb = design.combine_eq_taps(bank, gains)
    for s in range(S):
        ref = lfilter(main, 1.0, lfilter(comb, 1.0, x[s].astype(np.float64)).astype(np.float32).astype(np.float64))
        assert_parity(y[s], ref.astype(np.float32), "fused EQ on tensor cores")


@pytest.mark.parametrize("fmt,dither", [("f32", "none"), ("i24", "tpdf"), ("i24_packed", "tpdf")])
def test_long_blocks_chain_through_the_tensor_kernel(fmt, dither):
    """Blocks too long for one tensor-kernel launch (the reference's default blocksize is 4096) run as a chain
    of 512-frame sub-blocks with the history handed through a scratch buffer: bit-identical to feeding the
    512-frame blocks one by one (same kernel, same dither counters), state left where a single launch
    leaves it, and within tolerance of the oracle."""
    h = remez1001()
    S, n = 21, 4096 + 2048 + 512 + 4096
    x = noise(S, n, amp=0.2)
    kw = dict(out_format=fmt, dither=dither, seed=11)
    with BatchFirEngine(h, n_streams=S, max_frames=4096, **kw) as eng:
        y_big, k_big = run(eng, x, [4096, 2048, 512, 4096])
    with BatchFirEngine(h, n_streams=S, max_frames=512, **kw) as eng:
        y_small, k_small = run(eng, x, [512] * (n // 512))
    assert k_big == [TK(512)] * 4 and set(k_small) == {TK(512)}
    np.testing.assert_array_equal(y_big, y_small)
    if fmt == "f32":
        for s in range(0, S, 5):
            assert_parity(y_big[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"stream {s}")
