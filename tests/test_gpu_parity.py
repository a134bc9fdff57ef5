"""GPU parity: the CUDA path (through the C-ABI) against the oracle and the reference-generated
golden fixtures.  Float tolerance: 1e-5 relative or 1e-5 absolute (-100 dBFS), SURVEY.md 8(c);
integer output bit-exact."""
import numpy as np
import pytest

from conftest import load_golden, assert_parity

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402
from oracle import dither_quant as dq  # noqa: E402


def synth(S, n, kind="noise", seed=1234):
    """SURVEY.md 8(d) inputs: 0.1-amplitude noise (seed 1234+s) or a 20 Hz-20 kHz log sweep."""
    out = np.empty((S, n), dtype=np.float32)
    for s in range(S):
        if kind == "noise":
            out[s] = np.clip(0.1 * np.random.default_rng(seed + s).standard_normal(n), -1, 1)
        else:
            from scipy.signal import chirp
            t = np.arange(n) / 44100.0
            out[s] = 0.1 * chirp(t, 20, n / 44100.0, 20000, method="logarithmic", phi=90 * (s % 4))
    return out


def run_blocks(eng, x, B):
    """Feed x[S, n] block by block through the device path; returns the concatenated output."""
    S, n = x.shape
    xd = torch.from_numpy(x).cuda()
    outs = []
    for i in range(0, n, B):
        outs.append(eng.process(xd[:, i:i + B]).cpu().numpy())
    return np.concatenate(outs, axis=1)


TAPS = {
    51: dict(cutoff=8000, numtaps=51, window_type="hamming", filter_type="lowpass", samplerate=44100),
    251: dict(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass", samplerate=176400),
    1001: dict(method="remez", cutoff=20000, numtaps=1001, filter_type="lowpass", samplerate=176400),
}


@pytest.mark.parametrize("N,B,S", [(51, 1024, 5), (251, 1024, 4), (1001, 512, 3), (1001, 4096, 2),
                                   (251, 100, 3), (51, 37, 2), (1001, 64, 2), (2, 256, 1), (1, 128, 2)])
@pytest.mark.parametrize("kind", ["noise", "sweep"])
def test_streaming_fir_matches_lfilter(N, B, S, kind):
    """Kernel (1): block FIR with carried state == lfilter over the whole stream."""
    h = design.create_fir_filter(**TAPS[N]) if N in TAPS else np.array([0.75, -0.25][:N])
    n = B * 6
    x = synth(S, n, kind)
    with BatchFirEngine(h, n_streams=S, max_frames=B) as eng:
        y = run_blocks(eng, x, B)
    for s in range(S):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"N={N} B={B} s={s}")


@pytest.mark.parametrize("tag", ["c2", "c1", "default", "short"])
def test_fir_stage_matches_reference_callback(tag):
    """FIR + decimate stage against stream_process.audio_callback outputs (golden, reference code):
    feed the up-sampled blocks the reference consumed, keep every U-th output."""
    g = load_golden(f"cb_stream_process_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    up = g["upsampled"]                                  # [blocks, B*U]
    # stage-level check: plain engine at the processing rate, then [::U] on the host
    with BatchFirEngine(g["taps"], n_streams=1, max_frames=B * U) as eng:
        y = np.stack([eng.process(torch.from_numpy(up[i:i + 1]).cuda()).cpu().numpy()[0] for i in range(len(up))])
    assert_parity(y[:, ::U], g["pre_dither"].astype(np.float32), f"{tag} FIR stage")


@pytest.mark.parametrize("collapse", [True, False])
@pytest.mark.parametrize("tag", ["c1", "default", "short"])
def test_reference_chain_end_to_end(tag, collapse, monkeypatch):
    """interp -> FIR -> [::U] on device (chain='reference') against the reference callback's
    pre-dither block (the golden run used the same streaming interpolator): as ONE
    base-rate FIR with taps (h_up (*) h)[::U] (the default) and as the staged two-kernel form."""
    if not collapse:
        monkeypatch.setenv("AFP_NO_COLLAPSE", "1")
    g = load_golden(f"cb_stream_process_{tag}")
    N, B, U, fs = (int(v) for v in g["params"])
    nblk = g["pre_dither"].shape[0]
    x = g["x"][: nblk * B].reshape(1, -1)
    with BatchFirEngine(g["taps"], n_streams=1, max_frames=B, upsample=U, chain="reference",
                        upsampler_taps=g["up_taps"]) as eng:
        y = run_blocks(eng, x, B)
        stages = eng.stage_info()
    assert len(stages) == (1 if collapse and U > 1 else (2 if U > 1 else 1))
    assert_parity(y[0], g["pre_dither"].reshape(-1).astype(np.float32), f"{tag} reference chain")


def test_collapsed_reference_chain_batched_state_and_dither(monkeypatch):
    """The collapsed chain on a batch: equals the staged chain within tolerance, the state blob restores
    it, the dither counter advances per kept (base-rate) sample, ragged block sizes are fine."""
    U, S, B = 4, 9, 1024
    h = design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass",
                                 samplerate=44100 * U)
    h_up = U * design.create_fir_filter(cutoff=20000, numtaps=129, window_type="kaiser", samplerate=44100 * U)
    sizes = [1024, 1000, 37, 512, 1024]
    x = synth(S, sum(sizes), "noise")
    xd = torch.from_numpy(x).cuda()

    def go(**kw):
        with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=h_up, **kw) as e:
            outs, pos = [], 0
            for i, n in enumerate(sizes):
                if i == 2:                                   # save / reset / restore in mid-stream
                    blob = e.get_state()
                    e.reset()
                    e.set_state(blob)
                outs.append(e.process(xd[:, pos:pos + n].contiguous()).cpu().numpy())
                pos += n
            return np.concatenate(outs, axis=1), len(e.stage_info())

    y1, n1 = go()
    monkeypatch.setenv("AFP_NO_COLLAPSE", "1")
    y2, n2 = go()
    monkeypatch.delenv("AFP_NO_COLLAPSE")
    assert (n1, n2) == (1, 2) and y1.shape == (S, sum(sizes))
    for s in range(S):
        assert_parity(y1[s], y2[s], f"collapsed vs staged, stream {s}")
        ref = rc.streaming_fir(h, rc.polyphase_upsample_whole(h_up, x[s], U).astype(np.float32))[::U]
        assert_parity(y1[s], ref.astype(np.float32), f"collapsed vs oracle, stream {s}")
    q1, _ = go(out_format="i24", dither="tpdf", seed=5)
    ref_q = dq.epilogue(y1, dq.FMT_I24_IN_I32, dq.DITHER_TPDF, 5, np.arange(S), 0)
    np.testing.assert_array_equal(q1, ref_q)


def test_reference_chain_with_fused_eq_collapses_too(monkeypatch):
    """up-sample -> fused EQ -> FIR -> [::U]: one base-rate filter (h_up (*) eq (*) h)[::U]; equals the staged
    three-launch chain within tolerance."""
    U, S, B = 4, 5, 1024
    fs = 44100 * U
    h = design.create_fir_filter(cutoff=16000, numtaps=201, samplerate=fs)
    h_up = U * design.create_fir_filter(cutoff=20000, numtaps=129, window_type="kaiser", samplerate=fs)
    bank = design.design_eq_bank(129, fs)
    gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3]
    x = synth(S, B * 4, "noise")

    def go():
        with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=h_up,
                            eq_bands=bank, eq_gains=gains, eq_mode="fused") as e:
            return run_blocks(e, x, B), len(e.stage_info())

    y1, n1 = go()
    monkeypatch.setenv("AFP_NO_COLLAPSE", "1")
    y2, n2 = go()
    assert (n1, n2) == (1, 3) and y1.shape == (S, B * 4)
    for s in range(S):
        assert_parity(y1[s], y2[s], f"collapsed vs staged with EQ, stream {s}")


@pytest.mark.parametrize("U,N,B", [(2, 251, 1024), (4, 251, 1024), (3, 129, 512), (4, 1001, 512),
                                   (2, 51, 256), (4, 51, 100), (2, 501, 4096)])
def test_dac_chain_polyphase(U, N, B):
    """Kernel (2): polyphase interpolator (filter designed at fs*U is the interpolation filter)
    == zero-stuff + lfilter (scipy upfirdn), streaming across blocks."""
    h = U * design.create_fir_filter(cutoff=18000, numtaps=N, window_type="kaiser", samplerate=44100 * U)
    S = 3
    x = synth(S, B * 5, "noise")
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac") as eng:
        y = run_blocks(eng, x, B)
    assert y.shape == (S, B * 5 * U)
    for s in range(S):
        assert_parity(y[s], rc.polyphase_upsample_whole(h, x[s], U).astype(np.float32), f"U={U} N={N}")


def test_min_phase_variant_c4():
    """Kernel (4): gated minimum-phase design (501 -> 251 taps) + 2x polyphase."""
    h = design.create_fir_filter(cutoff=1000, numtaps=501, window_type="boxcar", filter_type="highpass",
                                 samplerate=88200)
    hm = design.apply_min_phase(h, 88200)
    assert len(hm) == 251
    S, B = 6, 1024
    x = synth(S, B * 4, "sweep")
    with BatchFirEngine(2 * hm, n_streams=S, max_frames=B, upsample=2, chain="dac") as eng:
        y = run_blocks(eng, x, B)
    for s in range(S):
        assert_parity(y[s], rc.polyphase_upsample_whole(2 * hm, x[s], 2).astype(np.float32), "min-phase")


@pytest.mark.parametrize("collapse", [True, False])
def test_eq_fused_streaming(collapse, monkeypatch):
    """Kernel (3), fused form: (sum_b g_b h_b) * x followed by the main FIR, state carried - as one filter
    eq (*) main (default at U = 1) and as two launches."""
    from scipy.signal import lfilter
    if not collapse:
        monkeypatch.setenv("AFP_NO_COLLAPSE", "1")
    fs = 48000
    bank = design.design_eq_bank(129, fs, design.EQ_BANDS_10)
    gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3, 0.9]
    main = design.create_fir_filter(cutoff=16000, numtaps=201, samplerate=fs)
    S, B = 2, 1024
    x = synth(S, B * 5, "noise")
    with BatchFirEngine(main, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=gains, eq_mode="fused") as eng:
        y = run_blocks(eng, x, B)
        assert len(eng.stage_info()) == (1 if collapse else 2)
    comb = design.combine_eq_taps(bank, gains)
    for s in range(S):
        ref = lfilter(main, 1.0, lfilter(comb, 1.0, x[s].astype(np.float64)).astype(np.float32).astype(np.float64))
        assert_parity(y[s], ref.astype(np.float32), "fused EQ")


@pytest.mark.parametrize("fmt,bits", [("i16", 16), ("i24", 24), ("i24_packed", 24)])
@pytest.mark.parametrize("U", [1, 4])
def test_quantise_bit_exact_no_dither(fmt, bits, U):
    """Integer output is bit-exact against the numpy quantiser applied to the kernel's own float
    output (SURVEY.md 8c quantisation definition); dither disabled."""
    h = design.create_fir_filter(cutoff=9000, numtaps=101, samplerate=44100 * U) * U
    S, B = 3, 512
    x = synth(S, B * 3, "noise") * 9.0      # reach the clip region too
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U) as ef:
        yf = run_blocks(ef, x, B)
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, out_format=fmt) as eq:
        yq = run_blocks(eq, x, B)
    ref = dq.quantise(yf, bits)
    assert (np.abs(yf) > 1.0).any()
    if fmt == "i24_packed":
        nb = B * U * 3
        yq = np.concatenate([dq.unpack_i24(yq[:, i:i + nb]) for i in range(0, yq.shape[1], nb)], axis=1)
    np.testing.assert_array_equal(yq.astype(np.int32), ref)


@pytest.mark.parametrize("mode,fmt", [("tpdf", "i24"), ("tpdf", "i16"), ("rpdf", "i24"), ("tpdf", "f32"),
                                      ("tpdf", "i24_packed")])
def test_dither_matches_philox_oracle(mode, fmt, monkeypatch):
    """Kernel (5): Philox4x32-10 TPDF/RPDF dither is bit-exact against the numpy statement for a
    fixed seed, and independent of the block size.  (Bit-exactness needs ONE arithmetic for the float values: the
    4096-frame blocks would otherwise run on the split-FP16 tensor kernel, whose results differ from the FFMA2 kernels'
    in the last bits; its epilogue is checked the same way in tests/test_gpu_tensor2.py.)"""
    monkeypatch.setenv("AFP_NO_TC2", "1")
    h = design.create_fir_filter(cutoff=9000, numtaps=101, samplerate=44100)
    S, n, seed, sid0 = 4, 4096, 2024, 7
    x = synth(S, n, "sweep")
    with BatchFirEngine(h, n_streams=S, max_frames=512) as ef:
        yf = run_blocks(ef, x, 512)
    outs = {}
    for B in (512, 4096, 200):
        with BatchFirEngine(h, n_streams=S, max_frames=B, out_format=fmt, dither=mode, seed=seed,
                            first_stream_id=sid0) as e:
            nn = (n // B) * B
            outs[B] = run_blocks(e, x[:, :nn], B)
    fmt_code = {"f32": dq.FMT_F32, "i16": dq.FMT_I16, "i24": dq.FMT_I24_IN_I32, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    ref = dq.epilogue(yf, fmt_code, {"tpdf": dq.DITHER_TPDF, "rpdf": dq.DITHER_RPDF}[mode], seed,
                      np.arange(S) + sid0, 0)
    k = 3 if fmt == "i24_packed" else 1
    for B, y in outs.items():
        np.testing.assert_array_equal(y, ref[:, : y.shape[1]], err_msg=f"block size {B}")
    assert outs[512].shape[1] == n * k


def test_tpdf_statistics():
    """TPDF dither with a fixed seed: zero mean, variance LSB^2/6, range (-1,1) LSB, triangular
    histogram, white."""
    S, n = 8, 1 << 16
    x = np.zeros((S, n), dtype=np.float32)
    with BatchFirEngine(np.array([1.0]), n_streams=S, max_frames=n, out_format="f32", dither="tpdf", seed=99) as e:
        y = e.process(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
    d = (y * 2.0 ** 23).ravel()                       # in LSB of the 24-bit grid
    assert abs(d.mean()) < 3e-3
    assert abs(d.var() - 1.0 / 6.0) < 0.01 / 6.0 * 3
    assert d.min() > -1.0 and d.max() < 1.0
    hist, edges = np.histogram(d, bins=40, range=(-1, 1))
    centres = 0.5 * (edges[1:] + edges[:-1])
    expect = (1 - np.abs(centres)) * (edges[1] - edges[0]) * d.size
    chi2 = ((hist - expect) ** 2 / expect).sum()
    assert chi2 < 90                                   # 39 dof, p ~ 1e-5
    d0 = d[: n] - d[: n].mean()
    assert abs(np.dot(d0[1:], d0[:-1]) / np.dot(d0, d0)) < 0.02
    # streams are independent
    assert abs(np.corrcoef(y[0], y[1])[0, 1]) < 0.02


def test_state_save_restore_and_reset():
    h = design.create_fir_filter(**TAPS[251])
    S, B = 3, 512
    x = synth(S, B * 4, "noise")
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format="i24", dither="tpdf", seed=5) as e:
        y_all = run_blocks(e, x, B)
        e.reset()
        y1 = run_blocks(e, x[:, : 2 * B], B)
        blob = e.get_state()
        y2 = run_blocks(e, x[:, 2 * B:], B)
        np.testing.assert_array_equal(np.concatenate([y1, y2], axis=1), y_all)
        e.reset()
        e.set_state(blob)
        y2b = run_blocks(e, x[:, 2 * B:], B)
        np.testing.assert_array_equal(y2b, y2)


def test_host_path_equals_device_path():
    h = design.create_fir_filter(**TAPS[251])
    S, B = 4, 1024
    x = synth(S, B * 3, "noise")
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=4) as a, \
            BatchFirEngine(h, n_streams=S, max_frames=B, upsample=4) as b:
        yd = run_blocks(a, x, B)
        yh = np.concatenate([b.process_host(x[:, i:i + B]) for i in range(0, x.shape[1], B)], axis=1)
    np.testing.assert_array_equal(yd, yh)


def test_large_batch_properties():
    """BASELINE-sized batch (4096 streams): linearity and stream independence, checked without
    the oracle; 64 streams checked against it."""
    h = design.create_fir_filter(**TAPS[251])
    S, B = 4096, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = 0.1 * torch.randn(S, B * 2, device="cuda", generator=g)
    xb = 0.1 * torch.randn(S, B * 2, device="cuda", generator=g)

    def run(x):
        with BatchFirEngine(h, n_streams=S, max_frames=B) as e:
            return torch.cat([e.process(x[:, :B]), e.process(x[:, B:])], dim=1)
    ya, yb, yab = run(xa), run(xb), run(xa + 2 * xb)
    assert torch.allclose(yab, ya + 2 * yb, atol=2e-6, rtol=0)
    xs = xa.cpu().numpy()
    for s in (0, 1, 63, 2047, 4095):
        assert_parity(ya[s].cpu().numpy(), rc.streaming_fir(h, xs[s]).astype(np.float32), f"stream {s}")


def test_errors_are_loud():
    from audio_filtering_pipeline_b200._lib import AfpError
    with pytest.raises(AfpError):
        BatchFirEngine(np.ones(5), n_streams=0)
    with pytest.raises(AfpError):
        BatchFirEngine(np.ones(5), upsample=7)
    with pytest.raises(AfpError):
        BatchFirEngine(np.array([np.nan]))
    with BatchFirEngine(np.ones(5), n_streams=2, max_frames=64) as e:
        with pytest.raises(AfpError):
            e.process(torch.zeros(2, 128, device="cuda"))
        with pytest.raises(ValueError):
            e.process(torch.zeros(3, 64, device="cuda"))
