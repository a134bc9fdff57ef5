"""Drop-in shapes of the reference's per-block API, backed by the CUDA engine.

Same names, positional arguments and error behaviour as the reference so that a script such as
stream_process.py keeps its structure and only swaps its imports:

    make_audio_callback      <- audio_callback(indata, outdata, frames, time, status)
                                stream_process.py:79-120, stream_process_GUI_Presets.py:617-686
    OverlapAddFilter         <- stream_process_GUI_Presets.py:35-123
    process_audio_block      <- stream_process_GUI_Presets.py:598-613
    apply_dither             <- stream_process.py:61-64
    safe_upsample            <- stream_process.py:66-77, stream_process_EQ_GUI.py:420-427
    safe_downsample          <- stream_process_EQ_GUI.py:429-436
    apply_agc                <- stream_process_AGC.py:43-89

All arithmetic runs on the GPU through BatchFirEngine; there is no numpy fallback - if the CUDA
library or the device is missing, constructing any of these raises.  What the callback does on a
*runtime* error is the reference's convention (never raise into PortAudio: replay the last good
block, else silence).
"""
from __future__ import annotations

from collections import deque

import numpy as np

from . import design
from .engine import BatchFirEngine, PinnedArray

__all__ = ["make_audio_callback", "OverlapAddFilter", "process_audio_block", "apply_dither",
           "safe_upsample", "safe_downsample", "apply_agc", "StreamChain", "Resampler", "close_helpers"]


class StreamChain:
    """The module-level state of stream_process.py (:19-51) as one object: design constants, the
    engine that replaces ``input_buffer`` + ``oaconvolve``, and the last-good-block deque.

    ``channels`` independent channels are batched as streams (the reference is mono, :22)."""

    def __init__(self, samplerate=44100, upsample_factor=4, blocksize=4096, numtaps=301, cutoff=11000,
                 window_type="hamming", filter_type="lowpass", method="window", channels=1,
                 eq_bands=None, eq_gains=None, eq_numtaps=None, eq_mode="buffer_same_per_band",
                 min_phase=False, dither=True, bit_depth=24, seed=0, device=0, upsampler_taps=None,
                 fir_coeff=None, agc=None, show_waterfall=False, upsampler_quality="fast"):
        self.samplerate, self.upsample_factor, self.blocksize = int(samplerate), int(upsample_factor), int(blocksize)
        self.channels = int(channels)
        up_rate = self.samplerate * self.upsample_factor                     # stream_process.py:20
        if fir_coeff is None:
            fir_coeff = design.create_fir_filter(method=method, cutoff=cutoff, numtaps=numtaps,   # :33-40
                                                 window_type=window_type, filter_type=filter_type,
                                                 samplerate=up_rate)
        if min_phase:
            fir_coeff = design.apply_min_phase(fir_coeff, up_rate)           # stream_process_EQ_GUI.py:377-379
        self.fir_coeff = np.asarray(fir_coeff, dtype=np.float64)
        bands = None
        if eq_bands is not None:
            n_eq = eq_numtaps or len(self.fir_coeff)
            bands = [design.create_fir_filter(method="window", cutoff=list(b), numtaps=n_eq,
                                              window_type="hamming", filter_type="bandpass",
                                              samplerate=up_rate) for b in eq_bands]   # stream_process_EQ.py:40-42
            if eq_gains is None:
                eq_gains = [1.0] * len(bands)
        if upsampler_taps is None and self.upsample_factor > 1:
            upsampler_taps = design.design_upsampler(self.upsample_factor, self.samplerate, quality=upsampler_quality)
        self.engine = BatchFirEngine(
            self.fir_coeff, n_streams=self.channels, max_frames=self.blocksize,
            upsample=self.upsample_factor, chain="reference", out_format="f32",
            dither="rpdf" if dither else "none", dither_bits=bit_depth, seed=seed, device=device,
            upsampler_taps=upsampler_taps, eq_bands=bands, eq_gains=eq_gains,
            eq_mode=eq_mode if bands is not None else "none", agc=agc)
        self.show_waterfall = bool(show_waterfall)                            # stream_process_GUI_Waterfall.py:471
        self.waterfall_data = deque(maxlen=50)                                # :112
        self.audio_buffer = deque(maxlen=4)                                   # stream_process.py:50
        self.silence_block = np.zeros((self.blocksize, self.channels), dtype=np.float32)
        # Everything the callback touches per block exists from here on (pinned staging, no allocation in
        # the call): page-locked in / out blocks in the [frames, channels] wire format, and the last-good blocks as a
        # ring of five buffers of which audio_buffer (maxlen 4) holds views - the slot being written is never in it.
        self._pin_in = PinnedArray((self.blocksize, self.channels), np.float32)
        self._pin_out = PinnedArray((self.engine.out_frames(self.blocksize), self.channels), np.float32)
        self._good = np.zeros((5, self.blocksize, self.channels), dtype=np.float32)
        self._good_i = 0

    def process(self, indata):
        """indata float32 [frames, channels] (or [frames]) -> float32 [frames, channels]: a view of the chain's pinned
        output block, valid until the next call."""
        x = np.asarray(indata)
        if x.ndim == 1:
            x = x[:, None]
        frames = x.shape[0]
        if frames > self.blocksize:
            raise ValueError(f"{frames} frames exceed the chain's blocksize {self.blocksize}")
        # [frames, channels] goes to the device as it is; the (de)interleave runs there
        xin = self._pin_in.a[:frames]
        np.copyto(xin, x[:, :self.channels], casting="same_kind")
        out = self._pin_out.a[:self.engine.out_frames(frames)]
        return self.engine.process_host_interleaved(xin, out=out)[:frames]

    def keep_good(self, y):
        """``audio_buffer.append(processed.copy())`` (stream_process.py:110) without the allocation."""
        slot = self._good[self._good_i][:y.shape[0], :y.shape[1]]
        np.copyto(slot, y)
        self.audio_buffer.append(slot)
        self._good_i = (self._good_i + 1) % 5

    def waterfall_row(self, column):
        """``20*log10(|rfft(column * hanning)| + 1e-12)`` of one output column on the GPU
        (stream_process_GUI_Waterfall.py:473-475); float32 [len(column)//2 + 1]."""
        import torch
        x = torch.from_numpy(np.ascontiguousarray(column, dtype=np.float32)[None, :]).cuda(self.engine.device)
        return self.engine.spectrum_db(x)[0].cpu().numpy()

    def close(self):
        self.engine.close()
        self._pin_in.free()
        self._pin_out.free()


def make_audio_callback(chain, verbose=True):
    """Returns ``audio_callback(indata, outdata, frames, time, status)`` for ``sd.Stream(callback=...)``
    with the reference's contract (stream_process.py:79-120): ``outdata`` is written in place, only
    the first ``chain.channels`` columns are used, status flags are printed, and the function never
    raises - on an error it replays the last good block, else writes silence."""

    def audio_callback(indata, outdata, frames, time, status):
        if status and verbose:
            print(f"Audio status: {status}")                                   # :82-85
        try:
            y = chain.process(indata)                                          # :89-113
            if y.shape[0] != frames:
                y = np.resize(y, (frames, y.shape[1]))                         # :106-107
            chain.keep_good(y)                                                 # :110
            outdata[:, :y.shape[1]] = y
            if chain.show_waterfall:                                           # stream_process_GUI_Waterfall.py:471-478
                try:
                    chain.waterfall_data.append(chain.waterfall_row(outdata[:, 0]))
                except Exception as exc:                                       # noqa: BLE001
                    if verbose:
                        print(f"Waterfall error: {exc}")
        except Exception as exc:                                               # noqa: BLE001 (:115-120)
            if verbose:
                print(f"Processing error: {exc}")
            if len(chain.audio_buffer) > 0:
                last = chain.audio_buffer[-1]
                outdata[:, :last.shape[1]] = np.resize(last, (frames, last.shape[1]))
            else:
                outdata[:] = 0.0

    return audio_callback


class OverlapAddFilter:
    """Stateful block FIR with the constructor, methods and attributes of the reference class
    (stream_process_GUI_Presets.py:35-123).  The carried state is the engine's input history, so
    unlike the reference it is also correct when ``len(coeffs) - 1 > block_size``."""

    def __init__(self, filter_coeffs, block_size, device=0):
        self._device = device
        self._engine = None
        self.set_coefficients(filter_coeffs, block_size)

    def set_coefficients(self, filter_coeffs, block_size):
        if len(filter_coeffs) == 0:                                            # :45-46
            self.coeffs = np.array([1.0], dtype=np.float32)
        else:
            self.coeffs = np.array(filter_coeffs, dtype=np.float32)           # :48 (taps truncated to f32)
        self.block_size = int(block_size)
        self.filter_length = len(self.coeffs)
        self.overlap_length = self.filter_length - 1
        need = self.block_size + self.filter_length - 1
        self.fft_length = int(2 ** np.ceil(np.log2(need)))                     # :56-57 (informational here)
        if self._engine is not None:
            self._engine.close()
        self._engine = BatchFirEngine(self.coeffs.astype(np.float64), n_streams=1,
                                      max_frames=self.block_size, device=self._device)
        self._tail_x = np.zeros(self.overlap_length, dtype=np.float32)

    def process_block(self, input_block):
        L = self.block_size
        x = np.asarray(input_block, dtype=np.float32).ravel()
        if len(x) != L:                                                        # :71-76
            x = np.pad(x, (0, L - len(x)), "constant") if len(x) < L else x[:L]
        y = self._engine.process_host(x[None, :])[0]
        if self.overlap_length:
            self._tail_x = np.concatenate([self._tail_x, x])[-self.overlap_length:]
        return y.astype(np.float32, copy=False)

    @property
    def overlap_buffer(self):
        """The reference's carried tail (:115-121): what the samples seen so far still contribute to
        the next ``N-1`` outputs.  Computed on request from the input history (GUI read-out only)."""
        n = self.overlap_length
        if n == 0:
            return np.zeros(0, dtype=np.float32)
        full = np.convolve(self._tail_x.astype(np.float64), self.coeffs.astype(np.float64))
        return full[n:2 * n].astype(np.float32)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


def process_audio_block(gui, upsampled_block):
    """``fir(sum_b gain_b * eq_b(x))`` over the filter objects a GUI holds
    (stream_process_GUI_Presets.py:598-613): ``gui.eq_filter_objects``, ``gui.eq_gains[i].get()``,
    ``gui.fir_filter_object``."""
    upsampled_block = np.asarray(upsampled_block, dtype=np.float32)
    current_eq_sum = np.zeros_like(upsampled_block)
    for eq_filter, gain_var in zip(gui.eq_filter_objects, gui.eq_gains):
        current_eq_sum += gain_var.get() * eq_filter.process_block(upsampled_block)
    if gui.fir_filter_object:
        return gui.fir_filter_object.process_block(current_eq_sum)
    return current_eq_sum


# --- helper shims ------------------------------------------------------------------------------------
# The reference's helpers are plain functions, called from the audio thread and from GUI worker threads alike, and its
# resampler (soxr.resample per block) keeps nothing between calls.  The engines behind the shims are cached per shape,
# so: one lock serialises their use (a context is single-owner), and a call WITHOUT a `state` object starts from a
# cleared history - two streams, channels or threads that happen to share a block length never see each other's
# samples.  Streaming use (history carried from block to block, what StreamChain does internally) takes an explicit
# per-stream Resampler.
import atexit
import threading

_ENGINES = {}
_LOCK = threading.RLock()


def _cached(key, factory):
    eng = _ENGINES.get(key)
    if eng is None:
        eng = _ENGINES[key] = factory()
    return eng


def close_helpers():
    """Release the engines the helper shims have created (also runs at interpreter exit)."""
    with _LOCK:
        for eng in _ENGINES.values():
            try:
                eng.close()
            except Exception:                                                  # noqa: BLE001
                pass
        _ENGINES.clear()


atexit.register(close_helpers)


class Resampler:
    """Per-stream state for :func:`safe_upsample` / :func:`safe_downsample`: an integer-factor polyphase interpolator
    (``up=True``) or anti-alias decimator with its own carried history.  One object per stream and direction."""

    def __init__(self, factor, sr=44100, blocksize=4096, up=True, device=0):
        self.factor, self.sr, self.blocksize, self.up = int(factor), int(sr), int(blocksize), bool(up)
        taps = design.design_upsampler(self.factor, self.sr)
        if up:
            self.engine = BatchFirEngine(taps, n_streams=1, max_frames=self.blocksize, upsample=self.factor, chain="dac", device=device)
        else:
            self.engine = BatchFirEngine(taps / self.factor, n_streams=1, max_frames=self.blocksize, upsample=1, chain="dac",
                                         decimate=self.factor, device=device)
        self._lock = threading.Lock()

    def process(self, x):
        with self._lock:
            return self.engine.process_host(np.ascontiguousarray(x, dtype=np.float32)[None, :])[0]

    def reset(self):
        with self._lock:
            self.engine.reset()

    def close(self):
        self.engine.close()


def apply_dither(audio, bit_depth=24):
    """``audio + uniform(-0.5, 0.5) LSB`` at LSB = 2/2**bit_depth (stream_process.py:61-64), drawn
    from the engine's Philox stream instead of numpy's global RNG (successive calls continue the stream; every
    block length has its own seed).  Returns float32."""
    a = np.asarray(audio, dtype=np.float32)
    flat = a.reshape(1, -1)
    with _LOCK:
        eng = _cached(("dither", flat.shape[1], int(bit_depth)), lambda: BatchFirEngine(
            [1.0], n_streams=1, max_frames=flat.shape[1], dither="rpdf", dither_bits=int(bit_depth), seed=0x5EED + len(_ENGINES)))
        return eng.process_host(flat).reshape(a.shape)


def apply_agc(signal, target_level=0.1, window_size=512, alpha=0.01, max_gain=10.0, attack=0.01, release=0.1):
    """Moving-RMS automatic gain control of one block (stream_process_AGC.py:43-89): same positional
    arguments (``alpha`` is accepted and unused, as there), same [n, 1] result shape; float32 instead of
    the reference's float64.  The clip to +-0.99 belongs to the caller (:143) and is not applied here.
    Where the reference's loop raises (a zero-sample attack / release constant, a window longer than
    the block) and returns the block unchanged, so does this."""
    import torch
    sig = np.asarray(signal, dtype=np.float32)
    sig = np.atleast_2d(sig).T if sig.ndim == 1 else sig                       # :52-53
    with _LOCK:
        # (outside the try: a missing library or device must raise, not look like the reference's pass-through)
        eng = _cached(("agc",), lambda: BatchFirEngine([1.0], n_streams=1, max_frames=1))
        try:
            x = np.ascontiguousarray(sig[:, 0])[None, :]
            xd = torch.from_numpy(x).cuda(eng.device)
            y = eng.agc_block(xd, target_level=target_level, window_size=window_size, max_gain=max_gain,
                              attack=attack, release=release, clip=0.0)
            return y.cpu().numpy()[0][:, None]
        except Exception as exc:                                               # noqa: BLE001 (:87-89)
            print(f"AGC error: {exc}, shape={sig.shape}")
            return sig


def safe_upsample(data, sr=44100, upsr=176400, state=None):
    """Interpolation by ``upsr/sr`` (1..4) with the polyphase FIR of ``design.design_upsampler``; zeros on failure like
    stream_process.py:75-77.  Without ``state`` every call is on its own, like the reference's one-shot ``soxr.resample``
    per block (the block starts from silence); with a :class:`Resampler` the history is carried from block to block
    (what a streaming interpolator should do, and what StreamChain does)."""
    x = np.asarray(data, dtype=np.float32)
    x = x[:, 0] if x.ndim > 1 else x                                           # :70
    if upsr == sr:
        return x.astype(np.float32)                                            # stream_process_EQ_GUI.py:422-423
    if upsr < sr:
        return safe_downsample(x, sr, upsr, int(round(len(x) * upsr / sr)), state=state)
    factor = int(round(upsr / sr))
    if state is not None:
        try:
            return state.process(x)
        except Exception as exc:                                               # noqa: BLE001
            print(f"Upsample error: {exc}")
            return np.zeros(int(len(data) * upsr / sr), dtype=np.float32)
    with _LOCK:
        # (outside the try: a missing library or device must raise, not turn into the reference's block of zeros)
        eng = _cached(("up", factor, len(x), int(sr)), lambda: BatchFirEngine(
            design.design_upsampler(factor, sr), n_streams=1, max_frames=len(x), upsample=factor, chain="dac"))
        try:
            eng.reset()
            return eng.process_host(x[None, :])[0]
        except Exception as exc:                                               # noqa: BLE001
            print(f"Upsample error: {exc}")
            return np.zeros(int(len(data) * upsr / sr), dtype=np.float32)


def safe_downsample(data, upsr, sr, expected_len, state=None):
    """Anti-alias FIR + keep every ``upsr/sr``-th sample (stream_process_EQ_GUI.py:429-436); ``state`` as in
    :func:`safe_upsample` (a ``Resampler(..., up=False)``)."""
    x = np.asarray(data, dtype=np.float32).ravel()
    if upsr == sr:
        return x
    factor = int(round(upsr / sr))
    if state is not None:
        try:
            y = state.process(x)
            return np.resize(y, expected_len) if len(y) != expected_len else y
        except Exception as exc:                                               # noqa: BLE001
            print(f"Downsample error: {exc}")
            return np.zeros(expected_len, dtype=np.float32)
    with _LOCK:
        eng = _cached(("down", factor, len(x), int(sr)), lambda: BatchFirEngine(
            design.design_upsampler(factor, sr) / factor, n_streams=1, max_frames=len(x), upsample=1,
            chain="dac", decimate=factor))                                     # (raises without library / device)
        try:
            eng.reset()
            y = eng.process_host(x[None, :])[0]
            return np.resize(y, expected_len) if len(y) != expected_len else y
        except Exception as exc:                                               # noqa: BLE001
            print(f"Downsample error: {exc}")
            return np.zeros(expected_len, dtype=np.float32)
