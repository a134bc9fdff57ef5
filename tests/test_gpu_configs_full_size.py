"""BASELINE.json configs at the sizes they are quoted on (round-1 review: the EQ tests stopped at 301 taps, the longest
run was 8 blocks, the FFMA2 kernels had no full-scale case, the interleaved entry point was only compared with the
planar one).

  C3  10 bands x 501 taps + 501-tap main, 48 kHz stereo, 1024-frame blocks - per-band, buffer-'same' per band (the exact
      arithmetic of stream_process_EQ.py:119-148, golden from the reference's own callback) and fused
  C1  the whole 10 s sweep: 431 blocks of 1024 frames through the reference chain with TPDF -> int24, against the oracle
      and against a block-at-a-time run resumed from a state blob at block 200
  full-scale (+-0.99) sine and sweep through the single-row and the row-pair FFMA2 kernels at 251 and 1001 taps
  the [frames, channels] wire format against the oracle directly
"""
import os

import numpy as np
import pytest
import scipy.signal

from conftest import load_golden, assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402
from oracle import dither_quant as dq  # noqa: E402


# ------------------------------------------------------------------------------------------------ C3
def test_c3_buffer_same_per_band_matches_reference_callback():
    """stream_process_EQ.audio_callback at C3 size (golden): 10 x 501 taps = 5120 padded taps, right under the
    per-launch table limit; both channels of the stereo pair as two streams of one engine."""
    g = load_golden("cb_stream_process_eq_n501_10band")
    N, B, U, fs = (int(v) for v in g["params"])
    nblk = g["pre_dither"].shape[1]
    x = np.ascontiguousarray(g["x"][:, : nblk * B])
    for mode in ("buffer_same_per_band", "buffer_same"):
        with BatchFirEngine(g["taps"], n_streams=2, max_frames=B, eq_bands=list(g["eq_taps"]), eq_gains=list(g["eq_gains"]),
                            eq_mode=mode) as eng:
            y = np.concatenate([eng.process_host(x[:, i * B:(i + 1) * B]) for i in range(nblk)], axis=1)
        for ch in range(2):
            assert_parity(y[ch], g["pre_dither"][ch].reshape(-1).astype(np.float32), f"C3 {mode} channel {ch}")


@pytest.mark.parametrize("mode", ["per_band", "fused"])
def test_c3_streaming_eq_full_size(mode):
    """The streaming forms (carried state, no block-edge error): 10 x 501-tap bands then the 501-tap main filter,
    stereo sessions; 'fused' folds everything into one 1001-tap filter (tensor kernel), 'per_band' accumulates ten
    tables.  Oracle: the ideal cascade  main * sum_b g_b (h_b * x)  with the EQ output rounded to float32 between the
    stages in the per-band form."""
    g = load_golden("cb_stream_process_eq_n501_10band")
    N, B, U, fs = (int(v) for v in g["params"])
    S, nblk = 6, 5
    x = np.stack([np.clip(0.05 * np.random.default_rng(40 + s).standard_normal(nblk * B), -1, 1) for s in range(S)]).astype(np.float32)
    xb = torch.from_numpy(np.ascontiguousarray(x.reshape(S, nblk, B).transpose(1, 0, 2))).cuda()
    with BatchFirEngine(g["taps"], n_streams=S, max_frames=B, eq_bands=list(g["eq_taps"]), eq_gains=list(g["eq_gains"]),
                        eq_mode=mode) as eng:
        y = np.concatenate(list(eng.process_blocks(xb).cpu().numpy()), axis=1)
        kernels = [st["kernel"] for st in eng.stage_info()]
    assert kernels == ([3] if mode == "fused" else [0, 3])
    comb = design.combine_eq_taps(list(g["eq_taps"]), list(g["eq_gains"]))
    for s in range(S):
        eq = scipy.signal.lfilter(comb, 1.0, x[s].astype(np.float64))
        if mode == "per_band":
            eq = eq.astype(np.float32).astype(np.float64)
        ref = scipy.signal.lfilter(g["taps"], 1.0, eq)
        assert_parity(y[s], ref.astype(np.float32), f"C3 {mode} stream {s}")


# ------------------------------------------------------------------------------------------------ C1, 10 s
def c1_sweep(n, fs=44100, phase=0.0):
    t = np.arange(n) / fs
    return (0.1 * scipy.signal.chirp(t, 20.0, n / fs, 20000.0, method="logarithmic", phi=phase)).astype(np.float32)


def test_c1_full_ten_second_sweep_431_blocks(monkeypatch):
    """C1 as BASELINE.json states it: 251-tap Kaiser band-pass designed at 176.4 kHz, 44.1 kHz mono, 4x up-sample ->
    FIR -> [::4], 1024-frame blocks, the whole 10 s logarithmic sweep = 431 blocks (last one zero padded), TPDF ->
    int24.  (a) float output of stream 0 against the oracle chain; (b) the int24 run is bit-identical to the numpy
    Philox statement over all 441 344 samples (the dither counter crosses 431 block boundaries); (c) the run cut at block
    200, its state saved and restored into a fresh engine, gives the same bits - on the tensor kernel (multi-block calls of the
    collapsed 95-tap chain) and, fed one block at a time, on the FFMA2 kernel (which a single 1024-frame block of a filter
    this short selects; the two kernels agree to the float tolerance, not to the bit, so each is compared with itself)."""
    fs, U, B = 44100, 4, 1024
    h = design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass", samplerate=fs * U)
    h_up = design.design_upsampler(U, fs)
    n = 441000
    nblk = -(-n // B)
    assert nblk == 431
    S = 3
    x = np.zeros((S, nblk * B), dtype=np.float32)
    for s in range(S):
        x[s, :n] = c1_sweep(n, fs, 90.0 * s)
    xb = torch.from_numpy(np.ascontiguousarray(x.reshape(S, nblk, B).transpose(1, 0, 2))).cuda()
    kw = dict(n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=h_up)
    with BatchFirEngine(h, **kw) as ef:
        yf = np.concatenate(list(ef.process_blocks(xb).cpu().numpy()), axis=1)
    ref0 = rc.streaming_fir(h, rc.polyphase_upsample_whole(h_up, x[0], U).astype(np.float32))[::U]
    assert_parity(yf[0], ref0.astype(np.float32), "C1 10 s sweep, stream 0")
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as eq:
        yq = np.concatenate(list(eq.process_blocks(xb).cpu().numpy()), axis=1)
    np.testing.assert_array_equal(yq, dq.epilogue(yf, dq.FMT_I24_IN_I32, dq.DITHER_TPDF, 2024, np.arange(S), 0))
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as e1:
        first = list(e1.process_blocks(xb[:200]).cpu().numpy())
        blob = e1.get_state()
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as e2:
        e2.set_state(blob)
        rest = list(e2.process_blocks(xb[200:]).cpu().numpy())
    np.testing.assert_array_equal(np.concatenate(first + rest, axis=1), yq)
    monkeypatch.setenv("AFP_NO_TC2", "1")
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as e0:
        yq0 = np.concatenate(list(e0.process_blocks(xb).cpu().numpy()), axis=1)
    assert np.abs(yq0.astype(np.int64) - yq).max() <= 1                     # the two kernels: at most one code apart
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as e1:
        first = [e1.process(xb[i]).cpu().numpy() for i in range(200)]
        blob = e1.get_state()
    with BatchFirEngine(h, out_format="i24", dither="tpdf", seed=2024, **kw) as e2:
        e2.set_state(blob)
        rest = [e2.process(xb[i]).cpu().numpy() for i in range(200, nblk)]
    np.testing.assert_array_equal(np.concatenate(first + rest, axis=1), yq0)


def test_dither_counter_past_2_to_32_samples():
    """The Philox counter is the absolute output-sample index (64 bit): a state blob whose counter sits just below 2^32
    continues across it exactly as the numpy statement says."""
    h = design.create_fir_filter(cutoff=9000, numtaps=51, samplerate=48000)
    S, B = 4, 1024
    x = np.clip(0.2 * np.random.default_rng(3).standard_normal((S, 3 * B)), -1, 1).astype(np.float32)
    start = 2 ** 32 - B - 5
    with BatchFirEngine(h, n_streams=S, max_frames=B) as ef:
        yf = np.concatenate([ef.process_host(x[:, i * B:(i + 1) * B]) for i in range(3)], axis=1)
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format="i24", dither="tpdf", seed=77) as e:
        blob = e.get_state()
        hdr = blob[:32].view(np.uint64)
        hdr[2] = start                                    # [magic, size, out_samples, 0] - include/afp.h afp_get_state
        e.set_state(blob)
        yq = np.concatenate([e.process_host(x[:, i * B:(i + 1) * B]) for i in range(3)], axis=1)
    np.testing.assert_array_equal(yq, dq.epilogue(yf, dq.FMT_I24_IN_I32, dq.DITHER_TPDF, 77, np.arange(S), start))


# ------------------------------------------------------------------------------------------------ full scale, FFMA2 kernels
@pytest.mark.parametrize("N", [251, 1001])
@pytest.mark.parametrize("kernel,B", [(2, 1024), (0, 1001)])     # aligned blocks: single-row kernel; odd blocks: row-pair kernel
def test_full_scale_signals_on_the_ffma2_kernels(N, kernel, B, monkeypatch):
    """+-0.99 sine in the pass band, +-0.99 logarithmic sweep and a +-0.99 square wave: the relative branch of the
    tolerance (1e-5 of |y|) at full level, fp32 accumulation over up to 1001 taps."""
    monkeypatch.setenv("AFP_NO_TC", "1")
    fs = 176400
    h = (design.create_fir_filter(method="remez", cutoff=20000, numtaps=1001, filter_type="lowpass", samplerate=fs) if N == 1001
         else design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass", samplerate=fs))
    nblk = 5
    n = nblk * B
    t = np.arange(n) / fs
    x = np.stack([0.99 * np.sin(2 * np.pi * 3000.0 * t),
                  0.99 * scipy.signal.chirp(t, 20.0, n / fs, 20000.0, method="logarithmic"),
                  0.99 * np.sign(np.sin(2 * np.pi * 700.0 * t)),
                  0.99 * np.sin(2 * np.pi * 997.0 * t)]).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    with BatchFirEngine(h, n_streams=4, max_frames=B) as eng:
        y = np.concatenate([eng.process(xd[:, i * B:(i + 1) * B].contiguous()).cpu().numpy() for i in range(nblk)], axis=1)
        assert eng.stage_info()[-1]["kernel"] == kernel
    for s in range(4):
        ref = rc.streaming_fir(h, x[s])
        assert_parity(y[s], ref.astype(np.float32), f"N={N} kernel {kernel} signal {s}")
        assert np.abs(ref).max() > 0.3                    # the case really is at level


# ------------------------------------------------------------------------------------------------ wire format vs oracle
@pytest.mark.parametrize("fmt,dither", [("f32", "none"), ("i16", "tpdf"), ("i24_packed", "tpdf")])
def test_interleaved_entry_point_against_the_oracle(fmt, dither):
    """afp_process_host_interleaved ([session][frame][channel] buffers, stream_process.py:128-136) against the oracle
    itself: every channel of every session is the oracle's streaming FIR of that channel's samples, and the integer
    formats are the numpy Philox/quantiser statement of it with stream id = session * channels + channel."""
    sessions, channels, B, nblk, U = 3, 2, 480, 4, 2
    h = U * design.create_fir_filter(cutoff=7000, numtaps=101, samplerate=48000 * U)
    rng = np.random.default_rng(12)
    x = np.clip(0.3 * rng.standard_normal((nblk, sessions, B, channels)), -1, 1).astype(np.float32)
    S = sessions * channels
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac") as ef:
        yf = np.concatenate([ef.process_host_interleaved(x[i]) for i in range(nblk)], axis=1)      # [sessions, nblk*B*U, channels]
    flat = np.empty((S, nblk * B * U), dtype=np.float32)
    for se in range(sessions):
        for ch in range(channels):
            xs = x[:, se, :, ch].reshape(-1)
            ref = rc.polyphase_upsample_whole(h, xs, U)
            assert_parity(yf[se, :, ch], ref.astype(np.float32), f"session {se} channel {ch}")
            flat[se * channels + ch] = yf[se, :, ch]
    if fmt == "f32":
        return
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac", out_format=fmt, dither=dither, seed=31) as eq:
        yq = [eq.process_host_interleaved(x[i]) for i in range(nblk)]
    code = {"i16": dq.FMT_I16, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    n_out = B * U
    for i in range(nblk):
        ref = dq.epilogue(flat[:, i * n_out:(i + 1) * n_out], code, dq.DITHER_TPDF, 31, np.arange(S), i * n_out)
        if fmt == "i24_packed":
            want = ref.reshape(sessions, channels, n_out, 3).transpose(0, 2, 1, 3)
        else:
            want = ref.reshape(sessions, channels, n_out).transpose(0, 2, 1)
        np.testing.assert_array_equal(yq[i], want, err_msg=f"block {i}")
