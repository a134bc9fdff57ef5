"""Per-stream-group filters (afp_assign_groups / afp_set_gains_group / afp_set_taps_group): sessions with their own EQ
gains inside one batch, as every reference process reads its own sliders per block (stream_process_EQ_GUI.py:86-90)
and swaps its filter under filter_lock (:380-385).  Oracle: one run of the reference arithmetic per group -
``combined = sum_b g_b h_b`` then the main filter, streaming with carried state."""
import numpy as np
import pytest
import scipy.signal

from conftest import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200._lib import AfpError  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402
from oracle import dither_quant as dq  # noqa: E402


def noise(S, n, seed=1, amp=0.05):
    return np.stack([np.clip(amp * np.random.default_rng(seed + s).standard_normal(n), -1, 1) for s in range(S)]).astype(np.float32)


def group_gains(G, n_bands, seed=5):
    return np.round(np.random.default_rng(seed).uniform(0.3, 2.0, (G, n_bands)), 3)


def oracle_fused(x, bank, gains, main):
    comb = design.combine_eq_taps(bank, gains)                       # stream_process_EQ_GUI.py:87-90
    return scipy.signal.lfilter(np.convolve(comb, main), 1.0, x.astype(np.float64))


@pytest.mark.parametrize("n_band_taps,B,use_blocks", [(101, 1024, False), (129, 480, False), (501, 512, True), (501, 512, False)])
def test_64_groups_with_their_own_gains(n_band_taps, B, use_blocks):
    """64 stereo sessions (128 streams), every session its own gain set, one engine, one launch per block: each stream
    equals the oracle run with its session's gains.  101/129-tap banks run on the row-pair FFMA2 kernel with per-item
    tables; the 501+501-tap chain (1001 combined taps) on the tensor cores when 16 blocks come in one call."""
    fs, G = 48000, 64
    bank = design.design_eq_bank(n_band_taps, fs)
    main = design.create_fir_filter(cutoff=16000, numtaps=n_band_taps, samplerate=fs)
    gains = group_gains(G, len(bank))
    S, nblk = 2 * G, 16 if use_blocks else 4
    x = noise(S, nblk * B)
    with BatchFirEngine(main, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=[1.0] * len(bank), eq_mode="fused") as eng:
        eng.assign_groups(np.arange(S) // 2)
        for g in range(G):
            eng.set_group_gains(g, gains[g])
        if use_blocks:
            xb = torch.from_numpy(np.ascontiguousarray(x.reshape(S, nblk, B).transpose(1, 0, 2))).cuda()
            y = np.concatenate(list(eng.process_blocks(xb).cpu().numpy()), axis=1)
            assert eng.kernel_launches == 1 and eng.stage_info()[-1]["kernel"] == 3
        else:
            y = np.concatenate([eng.process_host(x[:, i * B:(i + 1) * B]) for i in range(nblk)], axis=1)
            assert eng.stage_info()[-1]["kernel"] == 0
    for s in range(S):
        ref = oracle_fused(x[s], bank, gains[s // 2], main)
        assert_parity(y[s], ref.astype(np.float32), f"stream {s} (group {s // 2})")


def test_groups_with_own_main_taps_unaligned_pairs_and_dither():
    """Groups that differ in their MAIN filter (no EQ), an assignment in which neighbouring streams do not share a
    group (one row per work item), TPDF -> int16 output keyed by the global stream id."""
    fs, S, B, nblk = 44100, 11, 1000, 3
    cut = [3000, 6000, 9000, 12000]
    filt = [design.create_fir_filter(cutoff=c, numtaps=151, samplerate=fs) for c in cut]
    assign = np.array([0, 1, 2, 3, 3, 2, 1, 0, 0, 2, 1])
    x = noise(S, nblk * B, amp=0.3)
    kw = dict(n_streams=S, max_frames=B)
    with BatchFirEngine(filt[0], **kw) as ef, BatchFirEngine(filt[0], out_format="i16", dither="tpdf", seed=9, first_stream_id=100, **kw) as eq:
        for e in (ef, eq):
            e.assign_groups(assign)
            for g in range(1, 4):
                e.set_group_taps(g, filt[g])
        yf = np.concatenate([ef.process_host(x[:, i * B:(i + 1) * B]) for i in range(nblk)], axis=1)
        yq = np.concatenate([eq.process_host(x[:, i * B:(i + 1) * B]) for i in range(nblk)], axis=1)
    for s in range(S):
        assert_parity(yf[s], rc.streaming_fir(filt[assign[s]], x[s]).astype(np.float32), f"stream {s}")
    np.testing.assert_array_equal(yq, dq.epilogue(yf, dq.FMT_I16, dq.DITHER_TPDF, 9, np.arange(S) + 100, 0))


def test_group_gain_change_between_blocks_keeps_state():
    """The reference's live swap (stream_process_EQ_GUI.py:380-385): new gains for ONE group from the next block on, the
    carried samples untouched - the output is the new filter over the true past input; other groups do not notice."""
    fs, B = 48000, 512
    bank = design.design_eq_bank(129, fs)
    main = design.create_fir_filter(cutoff=15000, numtaps=129, samplerate=fs)
    g0, g1, g1b = [1.0] * 9, [0.5, 1.5, 1.0, 0.7, 1.2, 1.0, 2.0, 0.4, 1.1], [2.0, 0.3, 1.0, 1.7, 0.2, 1.0, 0.5, 1.4, 0.9]
    S, nblk = 4, 6
    x = noise(S, nblk * B, amp=0.1)
    with BatchFirEngine(main, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=g0, eq_mode="fused") as eng:
        eng.assign_groups([0, 0, 1, 1])
        eng.set_group_gains(1, g1)
        ys = []
        for i in range(nblk):
            if i == 3:
                eng.set_group_gains(1, g1b)
            ys.append(eng.process_host(x[:, i * B:(i + 1) * B]))
        y = np.concatenate(ys, axis=1)
    for s in range(S):
        if s < 2:
            ref = oracle_fused(x[s], bank, g0, main)
        else:   # time-varying filter: block i < 3 with g1, later blocks with g1b, both over the whole past input
            ra, rb = oracle_fused(x[s], bank, g1, main), oracle_fused(x[s], bank, g1b, main)
            ref = np.concatenate([ra[: 3 * B], rb[3 * B:]])
        assert_parity(y[s], ref.astype(np.float32), f"stream {s}")


def test_collapsed_reference_chain_with_groups():
    """up-sample x2 -> fused EQ -> FIR -> [::2] with per-session gains: one base-rate filter per group."""
    fs, U, B, G = 44100, 2, 1024, 5
    bank = design.design_eq_bank(101, fs * U)
    main = design.create_fir_filter(cutoff=15000, numtaps=101, samplerate=fs * U)
    h_up = design.design_upsampler(U, fs)
    gains = group_gains(G, len(bank), seed=8)
    S, nblk = 2 * G, 3
    x = noise(S, nblk * B, amp=0.1)
    with BatchFirEngine(main, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=h_up, eq_bands=bank,
                        eq_gains=[1.0] * 9, eq_mode="fused") as eng:
        eng.assign_groups(np.arange(S) // 2)
        for g in range(G):
            eng.set_group_gains(g, gains[g])
        y = np.concatenate([eng.process_host(x[:, i * B:(i + 1) * B]) for i in range(nblk)], axis=1)
        assert len(eng.stage_info()) == 1
    for s in range(S):
        up = rc.polyphase_upsample_whole(h_up, x[s], U)
        ref = oracle_fused(up, bank, gains[s // 2], main)[::U]
        assert_parity(y[s], ref.astype(np.float32), f"stream {s}")


def test_group_errors_and_clear():
    fs = 48000
    bank = design.design_eq_bank(51, fs)
    main = design.create_fir_filter(cutoff=9000, numtaps=51, samplerate=fs)
    x = noise(4, 256)
    with BatchFirEngine(main, n_streams=4, max_frames=256, eq_bands=bank, eq_mode="per_band") as eng:
        eng.assign_groups([0, 0, 1, 1])
        with pytest.raises(AfpError):          # per-band EQ is ten filters per stream, not one
            eng.process_host(x)
    with BatchFirEngine(main, n_streams=4, max_frames=256) as eng:
        with pytest.raises(AfpError):
            eng.assign_groups([0, 1, 2])       # one entry per stream
        with pytest.raises(AfpError):
            eng.set_group_taps(0, main)        # no groups yet
        eng.assign_groups([0, 1, 0, 1])
        with pytest.raises(AfpError):
            eng.set_group_taps(1, main[:-2])   # another length
        eng.set_group_taps(1, 2.0 * main)
        y = eng.process_host(x)
        np.testing.assert_allclose(y[1], 2.0 * y[0] * 0 + rc.streaming_fir(2.0 * main, x[1]), atol=2e-6)
        eng.assign_groups(None)
        eng.reset()
        y2 = eng.process_host(x)
        for s in range(4):
            assert_parity(y2[s], rc.streaming_fir(main, x[s]).astype(np.float32), f"ungrouped {s}")
        assert eng.stage_info()[-1]["kernel"] == 2
