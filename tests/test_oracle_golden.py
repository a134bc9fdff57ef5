"""Pins the oracle (oracle/reference_chain.py) to outputs of the reference's own code
(fixtures from tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from scipy import signal

from conftest import load_golden, assert_parity
from oracle import reference_chain as rc

# The restatement performs the same scipy calls on the same data, so agreement is far tighter than
# the 1e-5 contract; 1e-12 absolute leaves room for FFT-plan differences only.
TIGHT = 1e-12


@pytest.mark.parametrize("tag", ["c1", "default", "c2", "short"])
def test_sliding_block_fir_matches_stream_process(tag):
    g = load_golden(f"cb_stream_process_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    orc = rc.SlidingBlockFir(g["taps"], B, U)
    up = rc.PolyphaseUpsampler(g["up_taps"], U)
    for i in range(g["pre_dither"].shape[0]):
        blk = g["x"][i * B:(i + 1) * B]
        u = up.step(blk) if U > 1 else blk.copy()
        np.testing.assert_array_equal(u, g["upsampled"][i])       # stub interpolator == oracle's
        y = orc.step(u, B)
        np.testing.assert_allclose(y, g["pre_dither"][i], rtol=0, atol=TIGHT)
        # dither amplitude: +-2**-24 RPDF around the pre-dither block, then float32 (F6)
        d = g["outdata"][i].astype(np.float64) - g["pre_dither"][i]
        assert np.abs(d).max() <= 2.0 ** -24 + 2.0 ** -25


@pytest.mark.parametrize("tag", ["c1", "c2", "short"])
def test_block_loop_equals_streaming_fir(tag):
    """the sliding-buffer loop is lfilter(h,1,u)[::U] with zero initial state."""
    g = load_golden(f"cb_stream_process_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    u = g["upsampled"].reshape(-1)
    y = rc.streaming_fir(g["taps"], u, U)
    assert_parity(y, g["pre_dither"].reshape(-1), "block loop vs lfilter")
    assert np.abs(y - g["pre_dither"].reshape(-1)).max() < 1e-7


@pytest.mark.parametrize("tag", ["n101_u2", "n301_u4"])
def test_sliding_block_eq_matches_stream_process_eq(tag):
    g = load_golden(f"cb_stream_process_eq_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    orc = rc.SlidingBlockEq(g["taps"], list(zip(g["eq_taps"], g["eq_gains"])), B, U)
    for i in range(g["pre_dither"].shape[0]):
        y = orc.step(g["upsampled"][i], B)
        np.testing.assert_allclose(y, g["pre_dither"][i], rtol=0, atol=1e-9)


def test_sliding_block_eq_c3_size():
    """BASELINE.json configs[2] at full size: 10 bands x 501 taps + 501-tap main, 1024-frame blocks, both channels
    of the stereo pair (fixture: the reference's stream_process_EQ.audio_callback, run per channel)."""
    g = load_golden("cb_stream_process_eq_n501_10band")
    N, B, U, fs = (int(v) for v in g["params"])
    assert (N, B, U, fs) == (501, 1024, 1, 48000) and g["eq_taps"].shape == (10, 501)
    for ch in range(2):
        orc = rc.SlidingBlockEq(g["taps"], list(zip(g["eq_taps"], g["eq_gains"])), B, U)
        for i in range(g["pre_dither"].shape[1]):
            y = orc.step(g["x"][ch, i * B:(i + 1) * B], B)
            np.testing.assert_allclose(y, g["pre_dither"][ch, i], rtol=0, atol=1e-9)


@pytest.mark.parametrize("tag", ["n129_l2048", "n1001_l512", "n51_l256"])
def test_overlap_add_matches_presets_class(tag):
    g = load_golden(f"ola_{tag}")
    N, L = (int(v) for v in g["params"])
    flt = rc.OverlapAddFilter(g["taps"], L)
    assert flt.filter_length == N and flt.block_size == L and flt.overlap_length == N - 1
    for i in range(g["y"].shape[0]):
        y = flt.process_block(g["x"][i * L:(i + 1) * L])
        assert y.dtype == np.float32
        np.testing.assert_allclose(y, g["y"][i], rtol=0, atol=1e-9)
    flt2 = rc.OverlapAddFilter(g["taps"], L)
    np.testing.assert_allclose(flt2.process_block(g["x"][: L // 2]), g["y_short"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(flt2.process_block(g["x"][: L + 17]), g["y_long"], rtol=0, atol=1e-9)


def test_overlap_add_regimes():
    """F8: equals lfilter(float32(h)) when N-1 <= L, and is wrong (by design of the reference)
    when N-1 > L."""
    g = load_golden("ola_n129_l2048")
    ref = signal.lfilter(g["taps"].astype(np.float32).astype(np.float64), 1.0,
                         g["x"].astype(np.float64))
    assert np.abs(g["y"].reshape(-1) - ref).max() < 1e-6
    g = load_golden("ola_n1001_l512")
    ref = signal.lfilter(g["taps"].astype(np.float32).astype(np.float64), 1.0,
                         g["x"].astype(np.float64))
    assert np.abs(g["y"].reshape(-1) - ref).max() > 1e-3


def test_overlap_add_identity():
    g = load_golden("ola_identity")
    flt = rc.OverlapAddFilter([], 64)
    np.testing.assert_array_equal(flt.coeffs, g["coeffs"])
    np.testing.assert_allclose(flt.process_block(g["x"]), g["y"], rtol=0, atol=1e-9)


def test_process_audio_block_chain():
    g = load_golden("ola_chain_9band")
    N, L = (int(v) for v in g["params"])
    eq = [rc.OverlapAddFilter(c, L) for c in g["eq_taps"]]
    main = rc.OverlapAddFilter(g["taps"], L)
    for i in range(g["y"].shape[0]):
        y = rc.process_audio_block(eq, list(g["eq_gains"]), main, g["x"][i * L:(i + 1) * L])
        np.testing.assert_allclose(y, g["y"][i], rtol=0, atol=1e-8)


def test_waterfall_callback():
    g = load_golden("cb_waterfall_9band")
    N, B, U, fs = (int(v) for v in g["params"])
    orc = rc.WaterfallBlockEq(g["taps"], g["eq_taps"], g["eq_gains"], B, U)
    for i in range(g["pre_dither"].shape[0]):
        y = orc.step(g["x"][i * B:(i + 1) * B], B)
        np.testing.assert_allclose(y, g["pre_dither"][i], rtol=0, atol=1e-9)


def test_stateless_eq_worker():
    g = load_golden("eqgui_worker_stateless")
    N, B, U, fs = (int(v) for v in g["params"])
    for i in range(g["y"].shape[0]):
        y = rc.stateless_eq_block(g["x"][i * B:(i + 1) * B], list(g["eq_taps"]),
                                  list(g["eq_gains"]), g["taps"])
        np.testing.assert_allclose(y, g["y"][i], rtol=0, atol=TIGHT)


def test_fused_sum_identity():
    """sum_b g_b (h_b * x) == (sum_b g_b h_b) * x  (stream_process_EQ_GUI.py:86-90 vs
    stream_process_EQ.py:127-129), to rounding."""
    g = load_golden("cb_waterfall_9band")
    x = g["x"][:4096].astype(np.float64)
    per_band = sum(gn * signal.lfilter(c, 1.0, x) for c, gn in zip(g["eq_taps"], g["eq_gains"]))
    comb = signal.lfilter((g["eq_gains"][:, None] * g["eq_taps"]).sum(0), 1.0, x)
    assert np.abs(per_band - comb).max() < 1e-13


def test_polyphase_upsampler_block_vs_whole():
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal(4096)).astype(np.float32)
    from audio_filtering_pipeline_b200.design import design_upsampler
    for U in (2, 3, 4):
        h = design_upsampler(U, 44100)
        whole = rc.polyphase_upsample_whole(h, x, U)
        up = rc.PolyphaseUpsampler(h, U)
        blocks = np.concatenate([up.step(x[i:i + 512]) for i in range(0, 4096, 512)])
        assert np.abs(blocks - whole).max() < 1e-7


def test_reference_chain_collapses_to_one_base_rate_fir():
    """The identity behind afp_abi.cu::rebuild_fused, checked on the CPU with the oracle alone:
    up-sample by U (zero-stuff + h_up) -> FIR h at fs*U -> [::U]  ==  base-rate FIR with taps (h_up (*) h)[::U]."""
    from oracle import reference_chain as rc
    rng = np.random.default_rng(3)
    for U, L, N in [(4, 129, 251), (2, 33, 51), (3, 20, 7), (4, 5, 1001)]:
        h_up = rng.standard_normal(L) / L
        h = rng.standard_normal(N) / N
        x = rng.standard_normal(600)
        staged = rc.streaming_fir(h, rc.polyphase_upsample_whole(h_up, x, U))[::U]
        g0 = np.convolve(h_up, h)[::U]
        fused = rc.streaming_fir(g0, x)
        np.testing.assert_allclose(fused[: len(staged)], staged, rtol=0, atol=1e-12)


def _agc_cases():
    import json
    g = load_golden("agc_blocks")
    return g, json.loads(str(g["meta"]))


def test_agc_oracle_matches_apply_agc():
    """oracle/agc.py against the reference's own apply_agc (stream_process_AGC.py:43-89) on every
    golden block, the error path (block returned unchanged) included."""
    from oracle import agc
    g, meta = _agc_cases()
    assert sum(m["errored"] for m in meta) >= 4 and sum(not m["errored"] for m in meta) >= 20
    for m in meta:
        x, y_ref = g["x_" + m["block"]], g["y_" + m["key"]]
        y = agc.apply_agc(x, target_level=m["target_level"], window_size=m["window_size"],
                          max_gain=m["max_gain"], attack=m["attack"], release=m["release"])
        if m["errored"]:
            assert y.dtype == np.float32 and np.array_equal(y, x) and np.array_equal(y_ref, x), m["key"]
        else:
            assert y.dtype == np.float64
            np.testing.assert_allclose(y, y_ref, rtol=1e-13, atol=0, err_msg=m["key"])


def test_agc_stage_is_call_site_window_and_clip():
    from oracle import agc
    g, meta = _agc_cases()
    for m in meta:
        x = g["x_" + m["block"]]
        if m["window_size"] != min(512, len(x) // 2) or m["target_level"] != 0.1:
            continue
        y = agc.agc_stage(x)
        np.testing.assert_allclose(y, np.clip(g["y_" + m["key"]], -0.99, 0.99), rtol=1e-13, atol=0)
        assert np.abs(y).max() <= 0.99


def test_spectrum_oracle_matches_waterfall_rows():
    from oracle import spectrum
    g = load_golden("waterfall_spectrum")
    for blk, row in zip(g["outdata"], g["spectrum_db"]):
        np.testing.assert_allclose(spectrum.spectrum_db(blk), row, rtol=0, atol=1e-9)
