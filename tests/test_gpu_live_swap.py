"""afp_swap_taps: the reference's live filter swap (active_filter_params replaced under filter_lock,
stream_process_EQ_GUI.py:380-385 - the sliding input buffer is left alone): from the next block on the
output is the new filter over the TRUE past input.  Oracle: lfilter of the whole input with each filter,
stitched at the swap block; a filter longer than what was carried sees zeros before that."""
import numpy as np
import pytest
from scipy import signal

from conftest import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402


def rand_taps(n, seed):
    r = np.random.default_rng(seed)
    return r.standard_normal(n) * np.exp(-np.arange(n) / (0.4 * n + 3)) / np.sqrt(n)


def noise(S, n, seed=1):
    return np.clip(0.1 * np.random.default_rng(seed).standard_normal((S, n)), -1, 1).astype(np.float32)


@pytest.mark.parametrize("n1,n2,B", [(129, 129, 1024), (129, 65, 1024), (251, 33, 512), (65, 129, 1024), (33, 251, 512),
                                     (1001, 1001, 512), (1001, 701, 512), (501, 1001, 512), (51, 1001, 512), (1001, 51, 512)])
def test_swap_keeps_the_carried_samples(n1, n2, B):
    S, nblk, k = 4, 6, 3
    h1, h2 = rand_taps(n1, 1), rand_taps(n2, 2)
    x = noise(S, nblk * B, seed=n1 + n2)
    xd = torch.from_numpy(x).cuda()
    with BatchFirEngine(h1, n_streams=S, max_frames=B) as e:
        carried = e.stage_info()[-1]["history"]
        outs = []
        for i in range(nblk):
            if i == k:
                e.swap_taps(h2)
            outs.append(e.process(xd[:, i * B:(i + 1) * B]).cpu().numpy())
    y = np.concatenate(outs, axis=1)
    assert carried >= n1 - 1
    for s in range(S):
        y1 = rc.streaming_fir(h1, x[s])
        xz = x[s].astype(np.float64).copy()
        xz[:max(k * B - carried, 0)] = 0.0                       # what the old history did not hold
        y2 = signal.lfilter(h2, [1.0], xz)
        ref = np.concatenate([y1[:k * B], y2[k * B:]])
        assert_parity(y[s], ref, f"stream {s}")
    if n2 <= carried + 1:
        # the swap is glitch-free: identical to having run the new filter from the start
        for s in range(S):
            assert_parity(y[s, k * B:], rc.streaming_fir(h2, x[s])[k * B:], f"stream {s} (no transient)")


def test_swap_in_the_collapsed_reference_chain():
    """chain='reference', U=4: up-sample -> FIR -> [::4] runs as one collapsed base-rate filter; the swap
    rebuilds it and keeps its history."""
    S, B, nblk, k, U = 3, 1024, 5, 2, 4
    fs = 44100
    up = design.design_upsampler(U, fs)
    h1 = design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass",
                                  samplerate=fs * U)
    h2 = design.create_fir_filter(cutoff=6000, numtaps=251, window_type="hamming", filter_type="lowpass", samplerate=fs * U)
    x = noise(S, nblk * B, seed=9)
    xd = torch.from_numpy(x).cuda()
    with BatchFirEngine(h1, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=up) as e:
        outs = []
        for i in range(nblk):
            if i == k:
                e.swap_taps(h2)
            outs.append(e.process(xd[:, i * B:(i + 1) * B]).cpu().numpy())
    y = np.concatenate(outs, axis=1)
    for s in range(S):
        u = rc.polyphase_upsample_whole(up, x[s], U)
        y1 = rc.streaming_fir(h1, u, U)
        y2 = rc.streaming_fir(h2, u, U)
        assert_parity(y[s], np.concatenate([y1[:k * B], y2[k * B:]]), f"stream {s}")


def test_set_taps_still_resets():
    S, B = 2, 512
    h1, h2 = rand_taps(101, 3), rand_taps(101, 4)
    x = noise(S, 2 * B, seed=5)
    xd = torch.from_numpy(x).cuda()
    with BatchFirEngine(h1, n_streams=S, max_frames=B) as e:
        e.process(xd[:, :B])
        e.set_taps(h2)                                           # OverlapAddFilter.set_coefficients semantics
        y = e.process(xd[:, B:]).cpu().numpy()
    for s in range(S):
        assert_parity(y[s], rc.streaming_fir(h2, x[s, B:]), f"stream {s}")


@pytest.mark.parametrize("n1,n2", [(129, 65), (65, 129), (101, 101)])
def test_swap_migrates_the_sliding_buffer_of_the_buffer_same_modes(n1, n2):
    """AFP_EQ_BUFFER_SAME*: the carried state is the reference's input_buffer minus the new block (the last N_main-1
    samples, stream_process_EQ.py:44-45,119-121).  A swap to another filter length keeps those samples right-aligned;
    what a longer filter needs beyond them is zero.  Oracle: SlidingBlockEq of the new filter with its buffer preset to
    exactly that."""
    S, B, nblk, k = 3, 512, 5, 2
    h1, h2 = rand_taps(n1, 11), rand_taps(n2, 12)
    bank = design.design_eq_bank(33, 48000)[:3]
    gains = [1.0, 0.7, 1.3]
    x = noise(S, nblk * B, seed=n1 * 7 + n2)
    with BatchFirEngine(h1, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=gains, eq_mode="buffer_same_per_band") as e:
        outs = []
        for i in range(nblk):
            if i == k:
                e.swap_taps(h2)
            outs.append(e.process_host(x[:, i * B:(i + 1) * B]))
    y = np.concatenate(outs, axis=1)
    for s in range(S):
        o1 = rc.SlidingBlockEq(h1, list(zip(bank, gains)), B, 1)
        ref = [o1.step(x[s, i * B:(i + 1) * B]) for i in range(k)]
        o2 = rc.SlidingBlockEq(h2, list(zip(bank, gains)), B, 1)
        keep = min(n1 - 1, n2 - 1)                                  # what the old buffer held of the new one's history
        past = np.zeros(n2 - 1, dtype=np.float32)
        if keep:
            past[-keep:] = x[s, k * B - keep:k * B]
        o2.buffer[:] = 0
        o2.buffer[-(n2 - 1):] = past                                 # step() shifts this into the history position
        ref += [o2.step(x[s, i * B:(i + 1) * B]) for i in range(k, nblk)]
        assert_parity(y[s], np.concatenate(ref), f"stream {s}")


def test_swap_across_the_collapsed_and_the_staged_form(monkeypatch):
    """A swap that changes whether the chain runs as ONE collapsed filter (here forced with AFP_NO_COLLAPSE; in the field:
    the combined filter no longer fits one launch).  staged -> collapsed hands the interpolator's carried input samples to
    the collapsed filter: exact.  collapsed -> staged seeds the interpolator from them and restarts the main FIR from zero
    state: from the swap on the output equals the new filter over an up-sampled signal that is zero before the swap block
    (minus what the interpolator still holds)."""
    S, B, nblk, k, U = 2, 1024, 5, 2, 4
    fs = 44100
    up = design.design_upsampler(U, fs)
    h1 = design.create_fir_filter(cutoff=[500, 8000], numtaps=251, window_type="kaiser", filter_type="bandpass", samplerate=fs * U)
    h2 = design.create_fir_filter(cutoff=6000, numtaps=251, window_type="hamming", filter_type="lowpass", samplerate=fs * U)
    x = noise(S, nblk * B, seed=21)
    xd = torch.from_numpy(x).cuda()

    def run(first_staged):
        monkeypatch.setenv("AFP_NO_COLLAPSE", "1" if first_staged else "0")
        with BatchFirEngine(h1, n_streams=S, max_frames=B, upsample=U, chain="reference", upsampler_taps=up) as e:
            outs, forms, hist_up = [], [], None
            for i in range(nblk):
                if i == k:
                    hist_up = e.stage_info()[0]["history"]          # carried input samples of the first stage before the swap
                    monkeypatch.setenv("AFP_NO_COLLAPSE", "0" if first_staged else "1")
                    e.swap_taps(h2)
                outs.append(e.process(xd[:, i * B:(i + 1) * B]).cpu().numpy())
                forms.append(len(e.stage_info()))
        return np.concatenate(outs, axis=1), forms, hist_up

    # staged -> collapsed: exact continuation
    y, forms, n_hist = run(first_staged=True)
    assert forms[:k] == [2] * k and forms[k:] == [1] * (nblk - k)
    for s in range(S):
        u = rc.polyphase_upsample_whole(up, x[s], U)
        y1, y2 = rc.streaming_fir(h1, u, U), rc.streaming_fir(h2, u, U)
        # the collapsed filter (h_up * h2)[::U] needs (129 + 251 - 2) / 4 = 95 past input samples; the interpolator carried n_hist (48)
        xz = x[s].astype(np.float64).copy()
        xz[:k * B - n_hist] = 0.0
        y2z = rc.streaming_fir(h2, rc.polyphase_upsample_whole(up, xz.astype(np.float32), U), U)
        assert_parity(y[s, :k * B], y1[:k * B], f"stream {s} before")
        assert_parity(y[s, k * B:], y2z[k * B:], f"stream {s} after (history = the interpolator's carried samples)")
        assert_parity(y[s, (k + 1) * B:], y2[(k + 1) * B:], f"stream {s} settled")
    # collapsed -> staged: interpolator continues, main FIR restarts
    y, forms, hist_up = run(first_staged=False)
    assert forms[:k] == [1] * k and forms[k:] == [2] * (nblk - k)
    for s in range(S):
        u = rc.polyphase_upsample_whole(up, x[s], U)
        y1, y2 = rc.streaming_fir(h1, u, U), rc.streaming_fir(h2, u, U)
        uz = u.astype(np.float64).copy()
        uz[:k * B * U] = 0.0                                       # the main FIR's history (up-sampled signal) starts empty
        y2z = signal.lfilter(h2, [1.0], uz)[::U]
        assert_parity(y[s, :k * B], y1[:k * B], f"stream {s} before")
        assert_parity(y[s, k * B:], y2z[k * B:], f"stream {s} after (main FIR from zero state)")
        assert_parity(y[s, (k + 1) * B:], y2[(k + 1) * B:], f"stream {s} settled")
