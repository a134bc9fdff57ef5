"""GPU parity of the EQ modes, the operator-object API and the callback shims against fixtures made
by the reference's own code (tests/golden) and the oracle restatements."""
import numpy as np
import pytest

from conftest import load_golden, assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import compat, design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402


def run_host(eng, x, B):
    return np.concatenate([eng.process_host(x[:, i:i + B]) for i in range(0, x.shape[1], B)], axis=1)


@pytest.mark.parametrize("tag", ["n101_u2", "n301_u4"])
@pytest.mark.parametrize("mode", ["buffer_same_per_band", "buffer_same"])
def test_eq_buffer_same_matches_stream_process_eq(tag, mode):
    """stream_process_EQ.audio_callback (golden): per-band 'same' over the sliding buffer, 'valid'
    main FIR, [::U].  Both the per-band and the fused arithmetic must hold the tolerance."""
    g = load_golden(f"cb_stream_process_eq_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    nblk = g["pre_dither"].shape[0]
    x = g["x"][: nblk * B].reshape(1, -1)
    with BatchFirEngine(g["taps"], n_streams=1, max_frames=B, upsample=U, chain="reference",
                        upsampler_taps=g["up_taps"], eq_bands=list(g["eq_taps"]),
                        eq_gains=list(g["eq_gains"]), eq_mode=mode) as eng:
        y = run_host(eng, x, B)
    assert_parity(y[0], g["pre_dither"].reshape(-1).astype(np.float32), f"{tag} {mode}")


def test_eq_buffer_same_batched_streams_and_state():
    """Same mode, several streams at once; equals the oracle class stream by stream, and the state
    blob restores the sliding buffer."""
    g = load_golden("cb_stream_process_eq_n101_u2")
    N, B, U, fs = (int(v) for v in g["params"])
    S, nblk = 5, 4
    rng = np.random.default_rng(7)
    up = (0.1 * rng.standard_normal((S, nblk * B * U))).astype(np.float32)
    with BatchFirEngine(g["taps"], n_streams=S, max_frames=B * U, eq_bands=list(g["eq_taps"]),
                        eq_gains=list(g["eq_gains"]), eq_mode="buffer_same_per_band") as eng:
        y = run_host(eng, up, B * U)
        for s in range(S):
            orc = rc.SlidingBlockEq(g["taps"], list(zip(g["eq_taps"], g["eq_gains"])), B * U, 1)
            ref = np.concatenate([orc.step(up[s, i * B * U:(i + 1) * B * U], B * U) for i in range(nblk)])
            assert_parity(y[s], ref.astype(np.float32), f"stream {s}")
        eng.reset()
        y1 = run_host(eng, up[:, : 2 * B * U], B * U)
        blob = eng.get_state()
        y2 = run_host(eng, up[:, 2 * B * U:], B * U)
        eng.reset()
        eng.set_state(blob)
        np.testing.assert_array_equal(run_host(eng, up[:, 2 * B * U:], B * U), y2)
        np.testing.assert_array_equal(np.concatenate([y1, y2], axis=1), y)


@pytest.mark.parametrize("mode", ["buffer_same_per_band", "buffer_same"])
def test_waterfall_fixed_buffer_callback(mode):
    """stream_process_GUI_Waterfall callback (golden): fixed 4*B+512 sliding buffer, per-band 'same',
    'valid' main FIR, the OLDEST frames are read out (:132,431-458)."""
    g = load_golden("cb_waterfall_9band")
    N, B, U, fs = (int(v) for v in g["params"])
    assert U == 1
    nblk = g["pre_dither"].shape[0]
    x = g["x"][: nblk * B].reshape(1, -1)
    with BatchFirEngine(g["taps"], n_streams=1, max_frames=B, eq_bands=list(g["eq_taps"]),
                        eq_gains=list(g["eq_gains"]), eq_mode=mode, same_buffer_len=4 * B + 512) as eng:
        y = run_host(eng, x, B)
    assert_parity(y[0], g["pre_dither"].reshape(-1).astype(np.float32), f"waterfall {mode}")
    # cross-check with the oracle class on other data, 3 streams
    rng = np.random.default_rng(11)
    xs = (0.1 * rng.standard_normal((3, 5 * B))).astype(np.float32)
    with BatchFirEngine(g["taps"], n_streams=3, max_frames=B, eq_bands=list(g["eq_taps"]),
                        eq_gains=list(g["eq_gains"]), eq_mode=mode, same_buffer_len=4 * B + 512) as eng:
        ys = run_host(eng, xs, B)
    for s in range(3):
        orc = rc.WaterfallBlockEq(g["taps"], g["eq_taps"], g["eq_gains"], B, 1)
        ref = np.concatenate([orc.step(xs[s, i * B:(i + 1) * B], B) for i in range(5)])
        assert_parity(ys[s], ref.astype(np.float32), f"waterfall oracle stream {s}")


def test_stateless_same_matches_eq_gui_worker():
    """stream_process_EQ_GUI.process_audio (golden): fused EQ 'same' + main 'same', per block."""
    g = load_golden("eqgui_worker_stateless")
    N, B, U, fs = (int(v) for v in g["params"])
    nblk = g["y"].shape[0]
    L = g["y"].shape[1]
    with BatchFirEngine(g["taps"], n_streams=1, max_frames=L, eq_bands=list(g["eq_taps"]),
                        eq_gains=list(g["eq_gains"]), eq_mode="stateless_same") as eng:
        for i in range(nblk):
            blk = g["x"][i * L:(i + 1) * L].reshape(1, -1) if len(g["x"]) >= (i + 1) * L else None
            if blk is None:
                break
            ref = rc.stateless_eq_block(blk[0], list(g["eq_taps"]), list(g["eq_gains"]), g["taps"])
            y = eng.process_host(blk)[0]
            assert_parity(y, ref.astype(np.float32), f"block {i}")


@pytest.mark.parametrize("mode", ["per_band", "fused"])
def test_process_audio_block_chain_matches_presets(mode):
    """process_audio_block over 9 OverlapAddFilter EQ objects + main (golden, reference code) ==
    streaming EQ (per band or fused) + main FIR with carried state (N-1 <= L regime)."""
    g = load_golden("ola_chain_9band")
    N, L = (int(v) for v in g["params"])
    x = g["x"][: g["y"].shape[0] * L].reshape(1, -1)
    with BatchFirEngine(g["taps"].astype(np.float32), n_streams=1, max_frames=L,
                        eq_bands=[b.astype(np.float32) for b in g["eq_taps"]], eq_gains=list(g["eq_gains"]),
                        eq_mode=mode) as eng:
        y = run_host(eng, x, L)
    assert_parity(y[0], g["y"].reshape(-1), f"chain {mode}")


@pytest.mark.parametrize("tag", ["n129_l2048", "n51_l256"])
def test_overlap_add_filter_class(tag):
    """compat.OverlapAddFilter against the reference class's outputs (golden), incl. short/long
    input handling and the attributes the GUI reads."""
    g = load_golden(f"ola_{tag}")
    N, L = (int(v) for v in g["params"])
    flt = compat.OverlapAddFilter(g["taps"], L)
    assert (flt.filter_length, flt.block_size, flt.overlap_length) == (N, L, N - 1)
    assert flt.coeffs.dtype == np.float32
    ref_obj = rc.OverlapAddFilter(g["taps"], L)
    for i in range(g["y"].shape[0]):
        blk = g["x"][i * L:(i + 1) * L]
        y = flt.process_block(blk)
        ref_obj.process_block(blk)
        assert y.dtype == np.float32 and y.shape == (L,)
        assert_parity(y, g["y"][i], f"{tag} block {i}")
    np.testing.assert_allclose(flt.overlap_buffer, ref_obj.overlap_buffer, rtol=0, atol=2e-6)
    flt2 = compat.OverlapAddFilter(g["taps"], L)
    assert_parity(flt2.process_block(g["x"][: L // 2]), g["y_short"], "short block")
    flt3 = compat.OverlapAddFilter(g["taps"], L)
    assert_parity(flt3.process_block(g["x"][: L + 17]), g["y_long"], "long block")
    flt.close(); flt2.close(); flt3.close()


def test_overlap_add_filter_long_filter_is_correct():
    """N-1 > L: the reference class drops overlap; ours equals lfilter."""
    from scipy.signal import lfilter
    g = load_golden("ola_n1001_l512")
    N, L = (int(v) for v in g["params"])
    flt = compat.OverlapAddFilter(g["taps"], L)
    y = np.concatenate([flt.process_block(g["x"][i * L:(i + 1) * L]) for i in range(6)])
    ref = lfilter(g["taps"].astype(np.float32).astype(np.float64), 1.0, g["x"].astype(np.float64))
    assert_parity(y, ref.astype(np.float32), "long filter")
    assert np.abs(y - g["y"].reshape(-1)).max() > 1e-3      # and therefore differs from the reference's defect
    ident = compat.OverlapAddFilter([], 64)
    gi = load_golden("ola_identity")
    np.testing.assert_array_equal(ident.coeffs, gi["coeffs"])
    assert_parity(ident.process_block(gi["x"]), gi["y"], "identity")


def test_process_audio_block_with_filter_objects():
    g = load_golden("ola_chain_9band")
    N, L = (int(v) for v in g["params"])

    class Var:
        def __init__(self, v): self.v = v
        def get(self): return self.v

    class Gui:
        pass
    gui = Gui()
    gui.eq_filter_objects = [compat.OverlapAddFilter(c, L) for c in g["eq_taps"]]
    gui.eq_gains = [Var(float(v)) for v in g["eq_gains"]]
    gui.fir_filter_object = compat.OverlapAddFilter(g["taps"], L)
    for i in range(2):
        y = compat.process_audio_block(gui, g["x"][i * L:(i + 1) * L])
        assert_parity(y, g["y"][i], f"block {i}")


@pytest.mark.parametrize("tag", ["c1", "short"])
def test_audio_callback_shape_and_parity(tag):
    """make_audio_callback: sounddevice signature, in-place outdata, pre-dither parity with the
    reference callback (golden) and dither within +-1/2 LSB of it."""
    g = load_golden(f"cb_stream_process_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    chain = compat.StreamChain(samplerate=fs, upsample_factor=U, blocksize=B, fir_coeff=g["taps"],
                               upsampler_taps=g["up_taps"], dither=True, bit_depth=24)
    cb = compat.make_audio_callback(chain, verbose=False)
    for i in range(g["pre_dither"].shape[0]):
        indata = g["x"][i * B:(i + 1) * B].reshape(B, 1).copy()
        outdata = np.full((B, 1), np.nan, dtype=np.float32)
        assert cb(indata, outdata, B, None, None) is None
        ref = g["pre_dither"][i]
        d = outdata[:, 0].astype(np.float64) - ref
        assert np.abs(d).max() <= 2.0 ** -24 + 1e-5 * np.abs(ref).max() + 2e-7
    chain.close()


def test_audio_callback_never_raises():
    chain = compat.StreamChain(samplerate=44100, upsample_factor=2, blocksize=256, numtaps=51, cutoff=9000)
    cb = compat.make_audio_callback(chain, verbose=False)
    out = np.ones((256, 1), dtype=np.float32)
    cb(np.zeros((256, 1), dtype=np.float32), out, 256, None, None)
    good = out.copy()
    # a block larger than max_frames makes the engine refuse: last good block is replayed
    out2 = np.full((256, 1), 7.0, dtype=np.float32)
    cb(np.zeros((4096, 1), dtype=np.float32), out2, 256, None, None)
    np.testing.assert_array_equal(out2, good)
    chain.close()
    chain2 = compat.StreamChain(samplerate=44100, upsample_factor=1, blocksize=64, numtaps=11, cutoff=9000)
    cb2 = compat.make_audio_callback(chain2, verbose=False)
    out3 = np.full((64, 1), 7.0, dtype=np.float32)
    cb2(np.zeros((512, 1), dtype=np.float32), out3, 64, None, None)      # no good block yet -> silence
    assert not out3.any()
    chain2.close()


def test_helper_shims():
    x = (0.1 * np.random.default_rng(3).standard_normal(1024)).astype(np.float32)
    y = compat.apply_dither(x, 24)
    assert y.dtype == np.float32 and y.shape == x.shape
    d = (y.astype(np.float64) - x) * 2 ** 23
    assert np.abs(d).max() <= 0.5 + 0.07 and np.abs(d).max() > 0.3          # +-1/2 LSB (+ float32 rounding)
    y16 = compat.apply_dither(x, 16)
    assert 0.3 < np.abs((y16.astype(np.float64) - x) * 2 ** 15).max() <= 0.51
    up = compat.safe_upsample(x.reshape(-1, 1), 44100, 176400)
    assert up.shape == (4096,) and up.dtype == np.float32
    ref = rc.polyphase_upsample_whole(design.design_upsampler(4, 44100), x, 4)
    assert_parity(up, ref.astype(np.float32), "safe_upsample")
    assert np.array_equal(compat.safe_upsample(x, 44100, 44100), x)
    down = compat.safe_downsample(up, 176400, 44100, 1024)
    assert down.shape == (1024,)
    from scipy.signal import lfilter
    refd = lfilter(design.design_upsampler(4, 44100) / 4, 1.0, up.astype(np.float64))[::4]
    assert_parity(down, refd.astype(np.float32), "safe_downsample")


@pytest.mark.parametrize("collapse", [True, False])
def test_decimate_option_and_gain_update(collapse, monkeypatch):
    """decimate keeps every n-th output; set_gains swaps the EQ table without touching the state.  With the EQ
    folded into the main FIR (the default for the streaming fused form at U = 1) the carried state is the INPUT,
    so after a gain change the new gains also apply to the carried samples - what the reference's sliding
    input_buffer does (stream_process_EQ.py:119-129); the staged form (AFP_NO_COLLAPSE=1) carries EQ outputs
    computed with the old gains through the main filter's history."""
    from scipy.signal import lfilter
    if not collapse:
        monkeypatch.setenv("AFP_NO_COLLAPSE", "1")
    h = design.create_fir_filter(cutoff=5000, numtaps=151, samplerate=48000)
    x = (0.1 * np.random.default_rng(5).standard_normal((2, 4096))).astype(np.float32)
    with BatchFirEngine(h, n_streams=2, max_frames=1024, decimate=4) as e:
        y = run_host(e, x, 1024)
    for s in range(2):
        assert_parity(y[s], lfilter(h, 1.0, x[s].astype(np.float64))[::4].astype(np.float32), "decimate")
    bank = design.design_eq_bank(65, 48000)
    g1, g2 = [1.0] * 9, [0.5, 1.0, 2.0, 1.0, 0.3, 1.0, 1.5, 1.0, 0.7]
    with BatchFirEngine(h, n_streams=2, max_frames=1024, eq_bands=bank, eq_gains=g1, eq_mode="fused") as e:
        ya = run_host(e, x[:, :2048], 1024)
        e.set_gains(g2)
        yb = run_host(e, x[:, 2048:], 1024)
        assert len(e.stage_info()) == (1 if collapse else 2)
    for s in range(2):
        xs = x[s].astype(np.float64)
        e1 = lfilter(design.combine_eq_taps(bank, g1), 1.0, xs)
        e2 = lfilter(design.combine_eq_taps(bank, g2), 1.0, xs)
        if collapse:
            ref = np.concatenate([lfilter(h, 1.0, e1)[:2048], lfilter(h, 1.0, e2)[2048:]])
        else:
            eq = np.concatenate([e1[:2048], e2[2048:]]).astype(np.float32).astype(np.float64)
            ref = lfilter(h, 1.0, eq)
        assert_parity(np.concatenate([ya[s], yb[s]]), ref.astype(np.float32), "gain swap")


def test_helper_shims_are_stateless_unless_given_a_state():
    """The reference's safe_upsample is a one-shot soxr call per block (stream_process.py:66-77): nothing is carried.  Two
    streams with the same block length through the shim must not see each other's samples; streaming use takes an explicit
    per-stream compat.Resampler, which then equals the whole-signal interpolation."""
    r = np.random.default_rng(17)
    a, b = (0.1 * r.standard_normal((2, 2048))).astype(np.float32)
    h = design.design_upsampler(2, 44100)
    a0 = compat.safe_upsample(a[:1024], 44100, 88200)
    compat.safe_upsample(b[:1024], 44100, 88200)                       # another stream, same shape
    a0_again = compat.safe_upsample(a[:1024], 44100, 88200)
    assert np.array_equal(a0, a0_again)
    assert_parity(a0, rc.polyphase_upsample_whole(h, a[:1024], 2).astype(np.float32), "one-shot")
    st_a, st_b = compat.Resampler(2, 44100, 1024), compat.Resampler(2, 44100, 1024)
    ya, yb = [], []
    for i in range(2):                                                # interleaved streams, each with its own state
        ya.append(compat.safe_upsample(a[i * 1024:(i + 1) * 1024], 44100, 88200, state=st_a))
        yb.append(compat.safe_upsample(b[i * 1024:(i + 1) * 1024], 44100, 88200, state=st_b))
    assert_parity(np.concatenate(ya), rc.polyphase_upsample_whole(h, a, 2).astype(np.float32), "streaming a")
    assert_parity(np.concatenate(yb), rc.polyphase_upsample_whole(h, b, 2).astype(np.float32), "streaming b")
    dn = compat.Resampler(2, 44100, 2048, up=False)
    d = compat.safe_downsample(np.concatenate(ya)[:2048], 88200, 44100, 1024, state=dn)
    assert d.shape == (1024,)
    for s in (st_a, st_b, dn):
        s.close()
    # dither: successive calls continue the Philox stream, different lengths use different seeds
    z = np.zeros(256, dtype=np.float32)
    d1, d2 = compat.apply_dither(z), compat.apply_dither(z)
    assert not np.array_equal(d1, d2)
    assert not np.array_equal(compat.apply_dither(np.zeros(255, dtype=np.float32)), d1[:255])
    # caller-supplied host output buffers are validated before they reach C
    with BatchFirEngine([1.0], n_streams=2, max_frames=64) as e:
        with pytest.raises(ValueError):
            e.process_host(np.zeros((2, 64), np.float32), out=np.zeros((2, 32), np.float32))
        with pytest.raises(ValueError):
            e.process_host_blocks(np.zeros((3, 2, 64), np.float32), out=np.zeros((3, 2, 64), np.float64))


@pytest.mark.parametrize("channels", [1, 2])
def test_callback_allocates_no_block_buffers(channels):
    """the per-block callback runs on PortAudio's audio thread - pinned staging, no allocation in the call.
    After the first blocks, 50 callbacks must not allocate anything of a block's size (tracemalloc sees numpy's buffers);
    what remains are the small transient Python objects of the call itself."""
    import tracemalloc
    B = 4096                                                         # the reference's own blocksize (stream_process.py:23)
    chain = compat.StreamChain(samplerate=44100, upsample_factor=4, blocksize=B, numtaps=251, cutoff=[500, 8000],
                               window_type="kaiser", filter_type="bandpass", channels=channels)
    cb = compat.make_audio_callback(chain, verbose=False)
    x = (0.1 * np.random.default_rng(4).standard_normal((60, B, channels))).astype(np.float32)
    out = np.zeros((B, channels), dtype=np.float32)
    for i in range(10):
        cb(x[i], out, B, None, None)
    tracemalloc.start()
    base, _ = tracemalloc.get_traced_memory()
    tracemalloc.reset_peak()
    for i in range(10, 60):
        cb(x[i], out, B, None, None)
    cur, peak = tracemalloc.get_traced_memory()
    tracemalloc.stop()
    block_bytes = B * channels * 4
    assert peak - base < block_bytes // 2, f"the callback allocated {peak - base} bytes at its peak (a block is {block_bytes})"
    assert cur - base < 4096
    assert np.abs(out).max() > 0 and len(chain.audio_buffer) == 4
    assert np.array_equal(chain.audio_buffer[-1], out)
    chain.close()


def test_interleaved_entry_point_pipelines_session_chunks():
    """Large batches go through afp_process_host_interleaved in chunks of sessions on three streams; the result is the planar
    path's, whatever the chunking (4096 stereo sessions = 8 chunks)."""
    S, B, ch = 8192, 512, 2
    h = design.create_fir_filter(cutoff=5000, numtaps=101, samplerate=48000)
    x = (0.1 * np.random.default_rng(6).standard_normal((S // ch, B, ch))).astype(np.float32)
    with BatchFirEngine(h, n_streams=S, max_frames=B) as e:
        y = e.process_host_interleaved(x)
        e.reset()
        planar = np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(S, B)
        yp = e.process_host(planar)
    assert np.array_equal(y.transpose(0, 2, 1).reshape(S, B), yp)
    for s in (0, 1, 4097, 8191):
        assert_parity(yp[s], rc.streaming_fir(h, planar[s]), f"stream {s}")


def test_reference_chain_with_the_soxr_grade_interpolator():
    """C1's chain with the 1013-tap 'vhq' prototype (libsoxr's VHQ recipe, design.design_upsampler): collapsed to one
    (1013 + 251 - 1) / 4 = 316-tap base-rate filter; oracle = up-sample the whole signal, FIR at 4 fs, keep every 4th."""
    S, B, nblk, U, fs = 4, 1024, 4, 4, 44100
    up = design.design_upsampler(U, fs, quality="vhq")
    h = design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass", samplerate=fs * U)
    x = (0.1 * np.random.default_rng(31).standard_normal((S, nblk * B))).astype(np.float32)
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=up) as e:
        y = run_host(e, x, B)
        info = e.stage_info()
    assert len(info) == 1 and info[0]["n_taps"] == (len(up) + len(h) - 1 + U - 1) // U
    for s in range(S):
        u = rc.polyphase_upsample_whole(up, x[s], U)
        assert_parity(y[s], rc.streaming_fir(h, u, U), f"stream {s}")
