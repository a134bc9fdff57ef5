"""Split-FP16 tensor kernel (csrc/afp_fir_tc2.cuh, kernel id 3) and afp_process_blocks: parity against the float64
oracle of the reference's block loop (stream_process.py:96-107 as restated in oracle/reference_chain.py) at the
filter lengths / block sizes that select its two window layouts, multi-block launches bit-identical to
block-by-block feeding, small-signal accuracy of the hi + lo*2^-11 split, and the FP16 range flag."""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200._lib import AfpError  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402
from oracle import dither_quant as dq  # noqa: E402


def taps_for(N):
    if N == 1001:
        return design.create_fir_filter(method="remez", cutoff=20000, numtaps=1001, filter_type="lowpass", samplerate=176400)
    return design.create_fir_filter(cutoff=9000, numtaps=N, window_type="kaiser", samplerate=96000)


def noise(S, n, seed=1234, amp=0.1):
    return np.stack([np.clip(amp * np.random.default_rng(seed + s).standard_normal(n), -1, 1)
                     for s in range(S)]).astype(np.float32)


def blocks_of(x, B):
    S, n = x.shape
    return np.ascontiguousarray(x.reshape(S, n // B, B).transpose(1, 0, 2))


def unblock(y):
    return np.concatenate(list(y), axis=1)


# every chunk count D + 1 = 7..17 region: < 8 chunks (no early set), exactly 8, 9 and 17; long blocks (several steps per block)
@pytest.mark.parametrize("N,B,S,nb", [(1001, 512, 37, 5), (1025, 512, 16, 4), (513, 512, 3, 4), (512, 512, 17, 4),
                                      (401, 512, 5, 3), (449, 1024, 9, 3), (641, 4096, 2, 2), (900, 1536, 1, 3),
                                      (1001, 512, 160 * 16 + 3, 2)])
def test_block_by_block_matches_oracle(N, B, S, nb):
    h = taps_for(N)
    x = noise(S, B * nb, amp=0.3)
    xd = torch.from_numpy(x).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B) as eng:
        ys = []
        for i in range(nb):
            ys.append(eng.process(xd[:, i * B:(i + 1) * B]).cpu().numpy())
            assert eng.stage_info()[-1]["kernel"] == 3
        assert not eng.range_exceeded()
    y = np.concatenate(ys, axis=1)
    for s in sorted({0, min(1, S - 1), S // 2, S - 1}):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"N={N} B={B} s={s}")


@pytest.mark.parametrize("N,B,S,nb", [(1001, 512, 37, 16), (1001, 512, 16, 1), (1001, 512, 20, 2), (1001, 512, 20, 3),
                                      (577, 1024, 33, 5), (449, 512, 18, 7)])
@pytest.mark.parametrize("fmt,dither", [("f32", "none"), ("i24", "tpdf")])
def test_process_blocks_is_bit_identical_to_block_by_block(N, B, S, nb, fmt, dither):
    h = taps_for(N)
    x = noise(S, B * nb * 2, amp=0.2)
    xb = torch.from_numpy(blocks_of(x, B)).cuda()            # [2 nb, S, B]
    kw = dict(n_streams=S, max_frames=B, out_format=fmt, dither=dither, seed=5, first_stream_id=3)
    with BatchFirEngine(h, **kw) as eng:
        y_one = torch.stack([eng.process(xb[i]) for i in range(2 * nb)]).cpu().numpy()
    with BatchFirEngine(h, **kw) as eng:
        y_a = eng.process_blocks(xb[:nb])                    # two calls: the carried state crosses a call boundary
        y_b = eng.process_blocks(xb[nb:])
        assert eng.stage_info()[-1]["kernel"] == 3
        y_multi = torch.cat([y_a, y_b]).cpu().numpy()
        assert eng.kernel_launches == 2
    np.testing.assert_array_equal(y_multi, y_one)
    if fmt == "f32":
        y = unblock(y_multi)
        for s in (0, S - 1):
            assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"s={s}")


def test_process_blocks_strided_views_and_int24():
    """Non-contiguous block / stream strides (a [n, S, B] window of a larger buffer) and TPDF -> int24 against the
    numpy statement of the Philox epilogue."""
    h = taps_for(1001)
    S, B, nb = 19, 512, 4
    x = noise(S, B * nb, amp=0.4)
    big = torch.zeros((nb, S + 5, B + 64), dtype=torch.float32, device="cuda")
    xin = big[:, 2:2 + S, 32:32 + B]
    xin.copy_(torch.from_numpy(blocks_of(x, B)))
    obig = torch.zeros((nb, S + 3, B + 32), dtype=torch.int32, device="cuda")
    with BatchFirEngine(h, n_streams=S, max_frames=B) as ef:
        yf = ef.process_blocks(xin).cpu().numpy()
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format="i24", dither="tpdf", seed=2024) as eq:
        eq.process_blocks(xin, obig[:, 1:1 + S, 16:16 + B])
    yq = obig[:, 1:1 + S, 16:16 + B].cpu().numpy()
    ref = dq.epilogue(unblock(yf), dq.FMT_I24_IN_I32, dq.DITHER_TPDF, 2024, np.arange(S), 0)
    np.testing.assert_array_equal(unblock(yq), ref)
    assert int(obig[:, 0].abs().max()) == 0 and int(obig[:, :, :16].abs().max()) == 0   # nothing outside the window


def test_process_blocks_falls_back_block_by_block_for_other_chains():
    """A call the tensor kernel does not take (here blocks of 1000 frames: it wants multiples of 512) runs block by block on
    the FFMA2 kernels behind the same entry point."""
    h = 4 * design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass",
                                     samplerate=176400)
    S, B, nb = 6, 1000, 3
    x = noise(S, B * nb)
    xb = torch.from_numpy(blocks_of(x, B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=4, chain="dac") as eng:
        y_one = torch.stack([eng.process(xb[i]) for i in range(nb)]).cpu().numpy()
        eng.reset()
        y_multi = eng.process_blocks(xb).cpu().numpy()
        assert eng.stage_info()[-1]["kernel"] == 2
    np.testing.assert_array_equal(y_multi, y_one)
    assert_parity(unblock(y_multi)[0], rc.polyphase_upsample_whole(h, x[0], 4).astype(np.float32), "4x DAC chain, 1000-frame blocks")


def test_small_and_full_scale_signals():
    """The split keeps 22 significant bits per operand down to FP16's subnormal range (an absolute floor of ~2^-35).
    Worst case is a coherent full-scale signal: three error terms of 2^-22 sum_k |h_k x_k| (the two representations and
    the dropped lo*lo product) - measured 2.9e-6 of the peak for a 0.999 sine, against 1e-8 for noise; -120 dBFS noise
    stays within 1e-9 absolute (-180 dBFS).  All inside the 1e-5 parity tolerance."""
    h = taps_for(1001)
    n = 512 * 6
    t = np.arange(n) / 176400.0
    x = np.stack([0.999 * np.sin(2 * np.pi * 997.0 * t), 0.999 * np.sin(2 * np.pi * 15000.0 * t),
                  1e-6 * np.random.default_rng(1).standard_normal(n), 3e-4 * np.random.default_rng(2).standard_normal(n),
                  0.999 * np.sign(np.sin(2 * np.pi * 50.0 * t))]).astype(np.float32)
    with BatchFirEngine(h, n_streams=5, max_frames=512) as eng:
        y = unblock(eng.process_blocks(torch.from_numpy(blocks_of(x, 512)).cuda()).cpu().numpy())
        assert eng.stage_info()[-1]["kernel"] == 3
    for s in range(5):
        ref = rc.streaming_fir(h, x[s])
        assert_parity(y[s], ref.astype(np.float32), f"signal {s}")
        assert np.abs(y[s] - ref).max() <= 4e-6 * np.abs(ref).max() + 1e-9, f"signal {s}: split precision lost"


@pytest.mark.parametrize("U", [2, 3, 4])
def test_polyphase_full_scale_signals(U):
    """The same worst case through the polyphase form: full-scale sines and a square wave into a 1001-tap U-times interpolator
    (gain U), plus -120 dBFS noise."""
    fs = 44100.0
    h = U * design.create_fir_filter(cutoff=0.45 * fs, numtaps=1001, window_type="kaiser", samplerate=fs * U)
    n = 512 * 5
    t = np.arange(n) / fs
    x = np.stack([0.999 * np.sin(2 * np.pi * 997.0 * t), 0.999 * np.sin(2 * np.pi * 15000.0 * t),
                  1e-6 * np.random.default_rng(1).standard_normal(n), 0.999 * np.sign(np.sin(2 * np.pi * 50.0 * t))]).astype(np.float32)
    with BatchFirEngine(h, n_streams=4, max_frames=512, upsample=U, chain="dac") as eng:
        y = unblock(eng.process_blocks(torch.from_numpy(blocks_of(x, 512)).cuda()).cpu().numpy())
        assert eng.stage_info()[-1]["kernel"] == 3
    for s in range(4):
        ref = rc.polyphase_upsample_whole(h, x[s], U)
        assert_parity(y[s], ref.astype(np.float32), f"U={U} signal {s}")
        assert np.abs(y[s] - ref).max() <= 4e-6 * np.abs(ref).max() + 1e-9, f"U={U} signal {s}: split precision lost"


def test_state_roundtrip_across_process_blocks():
    h = taps_for(1001)
    S, B = 24, 512
    x = noise(S, B * 8)
    xb = torch.from_numpy(blocks_of(x, B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format="i24", dither="tpdf", seed=9) as e:
        y_all = e.process_blocks(xb).cpu().numpy()
        e.reset()
        y1 = e.process_blocks(xb[:3]).cpu().numpy()
        blob = e.get_state()
        e.reset()
        e.set_state(blob)
        y2 = e.process_blocks(xb[3:]).cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([y1, y2]), y_all)


def test_tensor2_and_other_kernels_share_history():
    """512-multiples on kernel 3, 256 frames on kernel 1, other aligned sizes on the single-row FFMA2 kernel, odd sizes
    on the row-pair kernel - all on one carried history."""
    h = taps_for(1001)
    sizes = [512, 100, 256, 1024, 37, 512, 2048]
    x = noise(4, sum(sizes))
    xd = torch.from_numpy(x).cuda()
    outs, kernels, pos = [], [], 0
    with BatchFirEngine(h, n_streams=4, max_frames=2048) as eng:
        for n in sizes:
            outs.append(eng.process(xd[:, pos:pos + n].contiguous()).cpu().numpy())
            kernels.append(eng.stage_info()[-1]["kernel"])
            pos += n
    assert kernels == [3, 2, 1, 3, 0, 3, 3]
    y = np.concatenate(outs, axis=1)
    for s in range(4):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"stream {s}")


def test_range_flag():
    """A sample beyond FP16's range: the host-buffer call reports AFP_ERR_RANGE, the device call sets the flag."""
    h = taps_for(1001)
    x = noise(16, 512)
    x[3, 100] = 1.0e6
    with BatchFirEngine(h, n_streams=16, max_frames=512) as eng:
        with pytest.raises(AfpError) as ei:
            eng.process_host(x)
        assert ei.value.code == 5
        eng.reset()
        eng.process(torch.from_numpy(x).cuda())
        assert eng.range_exceeded()
        assert not eng.range_exceeded()        # cleared by the query
        x[3, 100] = 0.5
        eng.reset()
        eng.process_host(x)                    # in range again: no error


# Ring-buffered window: every step count modulo the ring (1..9 steps per group, 16, 33), every number of history batches a
# group start loads (D = 1: one, D = 2..8: two, D >= 9: three), several groups per CTA (more than 148 x 16 streams) with a
# ragged last group, output formats on the transposed-store path (f32 / int24 in int32 / int16) and off it (packed int24).
@pytest.mark.parametrize("N,B,nb", [(51, 512, 1), (51, 512, 2), (64, 512, 3), (65, 1024, 2), (129, 512, 5), (251, 512, 7),
                                    (251, 1024, 4), (385, 512, 9), (513, 512, 6), (577, 1536, 3), (640, 512, 16),
                                    (1001, 512, 33)])
@pytest.mark.parametrize("S", [21, 148 * 16 * 2 + 5])
def test_ring_window_all_step_counts(N, B, nb, S, monkeypatch):
    monkeypatch.setenv("AFP_TC_MIN_TAPS", "32")              # short filters and single steps on the tensor kernel too
    if S > 1000 and (nb > 7 or N > 600):
        pytest.skip("large batch x long run: covered by the small batch")
    h = taps_for(N)
    x = noise(min(S, 64), B * nb * 2, amp=0.25, seed=N + nb)
    x = np.tile(x, ((S + x.shape[0] - 1) // x.shape[0], 1))[:S] * np.linspace(0.5, 1.0, S, dtype=np.float32)[:, None]
    xb = torch.from_numpy(blocks_of(np.ascontiguousarray(x), B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B) as eng:
        y = torch.cat([eng.process_blocks(xb[:nb]), eng.process_blocks(xb[nb:])]).cpu().numpy()
        assert eng.stage_info()[-1]["kernel"] == 3 and eng.kernel_launches == 2
    y = unblock(y)
    for s in sorted({0, 1, 15, 16, S // 2, S - 6, S - 1}):
        assert_parity(y[s], rc.streaming_fir(h, x[s]).astype(np.float32), f"N={N} B={B} nb={nb} s={s}")


@pytest.mark.parametrize("fmt,dither", [("i24", "tpdf"), ("i16", "tpdf"), ("i24_packed", "tpdf"), ("f32", "rpdf")])
def test_short_filter_epilogues_bit_exact(fmt, dither, monkeypatch):
    """The transposed-store epilogue (4- and 2-byte samples) and the direct one (packed int24) against the numpy statement of
    the Philox / quantiser epilogue, applied to the kernel's own float output."""
    monkeypatch.setenv("AFP_TC_MIN_TAPS", "32")
    N, B, nb, S = 251, 512, 6, 40
    h = taps_for(N)
    x = noise(S, B * nb, amp=0.3, seed=77)
    xb = torch.from_numpy(blocks_of(x, B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B) as eng:
        yf = unblock(eng.process_blocks(xb).cpu().numpy())
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format=fmt, dither=dither, seed=11, first_stream_id=5) as eng:
        yq = unblock(eng.process_blocks(xb).cpu().numpy())
        assert eng.stage_info()[-1]["kernel"] == 3
    code = {"f32": dq.FMT_F32, "i16": dq.FMT_I16, "i24": dq.FMT_I24_IN_I32, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    # the per-block layout of packed samples is 3 bytes per sample: compare block by block through the oracle's own packing
    ref = dq.epilogue(yf, code, {"tpdf": dq.DITHER_TPDF, "rpdf": dq.DITHER_RPDF}[dither], 11, np.arange(S) + 5, 0)
    np.testing.assert_array_equal(yq, ref)


# Polyphase interpolators (U = 2 / 3 / 4) on the same kernel: wider Toeplitz rows (64 U output columns per 64-sample row), cores
# C_q[jj][e] = g[8 q + jj + U e].  Phase lengths below / at / above a chunk boundary, one and several steps per group, a
# ragged last stream group, block-by-block vs one multi-block call, and the quantised formats through the generic epilogue.
@pytest.mark.parametrize("U,N,B,nb,S", [(2, 96, 512, 2, 5), (2, 251, 1024, 3, 37), (2, 501, 512, 5, 16), (2, 1001, 512, 4, 21),
                                        (2, 129, 1536, 2, 3), (4, 251, 512, 3, 19), (4, 501, 1024, 2, 33), (4, 1001, 512, 5, 7),
                                        (4, 160, 512, 1, 2), (2, 251, 512, 2, 148 * 16 + 9),
                                        (3, 90, 512, 2, 5), (3, 251, 1024, 3, 37), (3, 577, 512, 4, 18), (3, 1001, 1536, 2, 9),
                                        (3, 192, 512, 1, 16)])
def test_polyphase_interpolator_on_tensor_kernel(U, N, B, nb, S):
    h = U * design.create_fir_filter(cutoff=0.45 * 44100, numtaps=N, window_type="kaiser", samplerate=44100 * U)
    x = noise(min(S, 48), B * nb, amp=0.3, seed=7 * N + U)
    x = np.tile(x, ((S + x.shape[0] - 1) // x.shape[0], 1))[:S] * np.linspace(0.6, 1.0, S, dtype=np.float32)[:, None]
    xb = torch.from_numpy(blocks_of(np.ascontiguousarray(x), B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac") as eng:
        y_multi = unblock(eng.process_blocks(xb).cpu().numpy())
        assert eng.stage_info()[-1]["kernel"] == 3 and eng.kernel_launches == 1
        assert y_multi.shape == (S, U * B * nb)
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac") as eng:
        y_one = np.concatenate([eng.process(xb[i]).cpu().numpy() for i in range(nb)], axis=1)
        assert eng.stage_info()[-1]["kernel"] == 3
    np.testing.assert_array_equal(y_multi, y_one)
    for s in sorted({0, 1, S // 2, S - 2, S - 1}):
        assert_parity(y_multi[s], rc.polyphase_upsample_whole(h, x[s], U).astype(np.float32), f"U={U} N={N} B={B} s={s}")


@pytest.mark.parametrize("U", [2, 3, 4])
@pytest.mark.parametrize("fmt,dither", [("i24", "tpdf"), ("i16", "none"), ("i24_packed", "tpdf"), ("f32", "rpdf")])
def test_polyphase_epilogues_bit_exact(U, fmt, dither):
    N, B, nb, S = 251, 512, 3, 20
    h = U * design.create_fir_filter(cutoff=0.45 * 44100, numtaps=N, window_type="kaiser", samplerate=44100 * U)
    x = noise(S, B * nb, amp=0.3, seed=5)
    xb = torch.from_numpy(blocks_of(x, B)).cuda()
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac") as eng:
        yf = unblock(eng.process_blocks(xb).cpu().numpy())
    with BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, chain="dac", out_format=fmt, dither=dither, seed=3,
                        first_stream_id=11) as eng:
        yq = unblock(eng.process_blocks(xb).cpu().numpy())
        assert eng.stage_info()[-1]["kernel"] == 3
    code = {"f32": dq.FMT_F32, "i16": dq.FMT_I16, "i24": dq.FMT_I24_IN_I32, "i24_packed": dq.FMT_I24_PACKED}[fmt]
    ref = dq.epilogue(yf, code, {"tpdf": dq.DITHER_TPDF, "rpdf": dq.DITHER_RPDF, "none": dq.DITHER_NONE}[dither], 3, np.arange(S) + 11, 0)
    np.testing.assert_array_equal(yq, ref)
