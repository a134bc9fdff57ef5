"""ORACLE - test infrastructure, not product code.

CPU restatement (numpy/scipy, float64 arithmetic on float32 buffers) of the reference's block
loops for the streaming-FIR hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``audio_filtering_pipeline_b200``) never does.

Parity status: PINNED for the FIR / EQ / overlap-add / min-phase / design stages - every class here
is checked in ``tests/test_oracle_golden.py`` against fixtures produced by executing the
reference's *own* code (``tests/golden/make_golden.py`` imports /root/reference with the audio and
GUI modules stubbed).  UNPINNED for (a) the up-sampler: the reference calls libsoxr, which is
neither vendored nor installed (requirements.txt:5); the interpolator below is the builder-defined
polyphase specification; and (b) the dither noise sequence: the reference draws from numpy's
unseeded global MT19937 (stream_process.py:63), so only its scale is restated.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import numpy as np
from scipy import signal

__all__ = [
    "SlidingBlockFir",
    "SlidingBlockEq",
    "WaterfallBlockEq",
    "OverlapAddFilter",
    "process_audio_block",
    "stateless_eq_block",
    "apply_dither_scale",
    "streaming_fir",
    "PolyphaseUpsampler",
    "polyphase_upsample_whole",
    "zero_stuff",
]


# ---------------------------------------------------------------------------------------------
# (a3)/(a2)/(a4): the canonical headless loop, stream_process.py:45-46 and :96-107
# ---------------------------------------------------------------------------------------------
class SlidingBlockFir:
    """``stream_process.py`` block step.  State is the float32 sliding buffer of length
    ``N + B*U - 1`` (:45-46); each step shifts it left by one up-sampled block (:97-98), runs
    ``oaconvolve(buffer, h, mode='valid', axes=0)`` (:102) and keeps every U-th sample (:106-107).
    """

    def __init__(self, taps, block_frames, upsample):
        self.h = np.asarray(taps, dtype=np.float64)
        self.B = int(block_frames)
        self.U = int(upsample)
        self.buffer = np.zeros(len(self.h) + self.B * self.U - 1, dtype=np.float32)

    def step_full(self, upsampled):
        """Returns the float64 'processed' array of :102 (all B*U outputs at the processing rate)."""
        n = self.B * self.U
        upsampled = np.asarray(upsampled)
        if len(upsampled) != n:                      # :92-94
            upsampled = np.resize(upsampled, n)
        self.buffer[:-n] = self.buffer[n:]           # :97
        self.buffer[-n:] = upsampled                 # :98 (cast to float32 on assignment)
        return signal.oaconvolve(self.buffer, self.h, mode="valid", axes=0)   # :102

    def step(self, upsampled, frames=None):
        """Returns the float64 'downsampled' block of :106-107 (length ``frames``)."""
        frames = self.B if frames is None else frames
        processed = self.step_full(upsampled)
        down = processed[:: self.U][:frames]         # :106
        return np.resize(down, frames)               # :107


# ---------------------------------------------------------------------------------------------
# (a7): per-band EQ over the sliding buffer, stream_process_EQ.py:119-148
# ---------------------------------------------------------------------------------------------
class SlidingBlockEq:
    """``stream_process_EQ.py`` block step: per-band ``oaconvolve(buffer, c_b, 'same')`` summed
    with gains into a float32 array (:126-129), then ``'valid'`` main FIR (:134), ``[::U]``
    (:139) and zero-pad/trim to ``frames`` (:143-148)."""

    def __init__(self, taps, eq_filters, block_frames, upsample):
        self.h = np.asarray(taps, dtype=np.float64)
        self.eq = [(np.asarray(c, dtype=np.float64), float(g)) for c, g in eq_filters]
        self.B = int(block_frames)
        self.U = int(upsample)
        self.buffer = np.zeros(len(self.h) + self.B * self.U - 1, dtype=np.float32)

    def eq_buffer(self):
        out = np.zeros_like(self.buffer)             # :126 float32
        for coeffs, gain in self.eq:                 # :127-129
            band = signal.oaconvolve(self.buffer, coeffs, mode="same")
            out += gain * band
        return out

    def step(self, upsampled, frames=None):
        frames = self.B if frames is None else frames
        n = self.B * self.U
        upsampled = np.asarray(upsampled)
        if len(upsampled) < n:                       # :112-117
            pad = np.zeros(n, dtype=np.float32)
            pad[: len(upsampled)] = upsampled
            upsampled = pad
        elif len(upsampled) > n:
            upsampled = upsampled[:n]
        self.buffer[:-n] = self.buffer[n:]           # :120
        self.buffer[-n:] = upsampled                 # :121
        eq_out = self.eq_buffer()
        final = signal.oaconvolve(eq_out, self.h, mode="valid", axes=0)      # :134
        down = final[:: self.U]                      # :139
        if len(down) < frames:                       # :143-148
            pad = np.zeros(frames, dtype=np.float32)
            pad[: len(down)] = down
            return pad
        return down[:frames]


class WaterfallBlockEq:
    """``stream_process_GUI_Waterfall.py`` callback body (:431-458): fixed ``4*B + 512`` buffer
    (:132), per-band ``fftconvolve(buffer, c_b, 'same')`` (:439-445), ``fftconvolve(eq, h,
    'valid')`` (:450) and the *oldest* ``frames`` decimated samples (:453)."""

    def __init__(self, taps, eq_taps, gains, block_frames, upsample):
        self.h = np.asarray(taps, dtype=np.float64)
        self.eq = [np.asarray(c, dtype=np.float64) for c in eq_taps]
        self.gains = [float(g) for g in gains]
        self.B = int(block_frames)
        self.U = int(upsample)
        self.buffer = np.zeros(self.B * 4 + 512, dtype=np.float32)       # :132
        self.eq_output = np.zeros_like(self.buffer)

    def step(self, upsampled, frames=None):
        frames = self.B if frames is None else frames
        upsampled = np.asarray(upsampled)
        n = len(upsampled)
        self.buffer[:-n] = self.buffer[n:]           # :435
        self.buffer[-n:] = upsampled                 # :436
        self.eq_output.fill(0)                       # :439
        for coeffs, gain in zip(self.eq, self.gains):
            band = signal.fftconvolve(self.buffer, coeffs, mode="same")   # :444
            self.eq_output += gain * band
        final = signal.fftconvolve(self.eq_output, self.h, mode="valid")  # :450
        down = final[:: self.U][:frames]             # :453
        if len(down) < frames:                       # :457-460
            down = np.pad(down, (0, frames - len(down)))
        return down


# ---------------------------------------------------------------------------------------------
# (a9)/(a10): the stateful operator object, stream_process_GUI_Presets.py:35-123 and :598-613
# ---------------------------------------------------------------------------------------------
class OverlapAddFilter:
    """Overlap-add FIR with a carried tail of ``N-1`` samples.  Only the first ``min(L, N-1)``
    samples of the tail are ever added back (:101-109), so for ``N-1 > L`` the output is wrong
     - restated faithfully, defect included."""

    def __init__(self, filter_coeffs, block_size):
        self.set_coefficients(filter_coeffs, block_size)

    def set_coefficients(self, filter_coeffs, block_size):
        if len(filter_coeffs) == 0:                                      # :45-46
            self.coeffs = np.array([1.0], dtype=np.float32)
        else:
            self.coeffs = np.array(filter_coeffs, dtype=np.float32)     # :48 taps truncated to f32
        self.block_size = block_size
        self.filter_length = len(self.coeffs)
        need = self.block_size + self.filter_length - 1
        self.fft_length = int(2 ** np.ceil(np.log2(need)))              # :56-57
        self.overlap_length = self.filter_length - 1
        self.filter_fft = np.fft.rfft(self.coeffs, self.fft_length)     # :63
        self.overlap_buffer = np.zeros(self.overlap_length, dtype=np.float32)

    def process_block(self, input_block):
        L = self.block_size
        if len(input_block) != L:                                        # :71-76
            if len(input_block) < L:
                input_block = np.pad(input_block, (0, L - len(input_block)), "constant")
            else:
                input_block = input_block[:L]
        padded = np.pad(input_block, (0, self.fft_length - L), "constant", constant_values=0)
        conv = np.fft.irfft(np.fft.rfft(padded) * self.filter_fft, self.fft_length)  # :83-90
        out = np.zeros(L, dtype=np.float32)
        k = min(L, self.overlap_length)                                  # :101
        if k > 0:
            out[:k] = conv[:k] + self.overlap_buffer[:k]                 # :105
        if L > k:
            out[k:] = conv[k:L]                                          # :109
        tail = conv[L: L + self.overlap_length]                          # :115
        if len(tail) < self.overlap_length:
            tail = np.pad(tail, (0, self.overlap_length - len(tail)), "constant")
        self.overlap_buffer = tail                                       # :121 (stays float64)
        return out


def process_audio_block(eq_filter_objects, gains, fir_filter_object, upsampled_block):
    """``process_audio_block(gui, upsampled_block)`` (stream_process_GUI_Presets.py:598-613) with
    the GUI object unpacked: ``fir(sum_b g_b * eq_b(x))``."""
    total = np.zeros_like(upsampled_block)
    for flt, g in zip(eq_filter_objects, gains):
        total += float(g) * flt.process_block(upsampled_block)   # python-float gain: stays float32
    if fir_filter_object:
        return fir_filter_object.process_block(total)
    return total


# ---------------------------------------------------------------------------------------------
# (a8): the stateless fused-sum worker, stream_process_EQ_GUI.py:86-95
# ---------------------------------------------------------------------------------------------
def stateless_eq_block(upsampled, eq_filters, eq_gains, main_taps=None):
    """``combined = sum_b g_b*h_b`` (:87-89), ``oaconvolve(x, combined, 'same')`` (:90), then the
    main filter, also ``'same'`` (:92-95).  No state crosses blocks."""
    combined = np.zeros_like(eq_filters[0])
    for coeffs, gain in zip(eq_filters, eq_gains):
        combined += gain * coeffs
    eq_out = signal.oaconvolve(upsampled, combined, mode="same")
    if main_taps is not None:
        return signal.oaconvolve(eq_out, main_taps, mode="same")
    return eq_out


# ---------------------------------------------------------------------------------------------
# (a6): dither scale, stream_process.py:61-64
# ---------------------------------------------------------------------------------------------
def apply_dither_scale(audio, uniform01, bit_depth=24):
    """``audio + (u - 0.5) * (2 / 2**bit_depth)`` with the uniform draw supplied by the caller
    (the reference uses the unseeded global RNG; the noise sequence itself is unpinned)."""
    return audio + (np.asarray(uniform01) - 0.5) * (2 / (2 ** bit_depth))


# ---------------------------------------------------------------------------------------------
# Whole-signal statements used for size-independent checks
# ---------------------------------------------------------------------------------------------
def streaming_fir(taps, x, decimate=1):
    """``lfilter(h, 1, x)[::D]`` in float64 - identical to the sliding-buffer loop with zero
    initial state (max deviation ~2e-8 from FFT rounding)."""
    y = signal.lfilter(np.asarray(taps, dtype=np.float64), 1.0, np.asarray(x, dtype=np.float64))
    return y[::decimate]


def zero_stuff(x, factor):
    x = np.asarray(x)
    out = np.zeros(len(x) * factor, dtype=np.float64)
    out[::factor] = x
    return out


def polyphase_upsample_whole(taps, x, factor):
    """Builder-defined interpolator over a whole signal: zero-stuff by ``factor`` then FIR with
    ``taps`` at the high rate, truncated to ``len(x)*factor`` (== ``upfirdn(taps, x, up=factor)``
    truncated).  The device computes only the non-zero products (polyphase form)."""
    return signal.upfirdn(np.asarray(taps, dtype=np.float64), np.asarray(x, dtype=np.float64),
                          up=factor)[: len(x) * factor]


class PolyphaseUpsampler:
    """Block-at-a-time version of :func:`polyphase_upsample_whole` with carried state, float32
    output (what the next stage's float32 sliding buffer would hold, stream_process.py:98)."""

    def __init__(self, taps, factor):
        self.h = np.asarray(taps, dtype=np.float64)
        self.U = int(factor)
        self.zi = np.zeros(max(len(self.h) - 1, 0))

    def step(self, block):
        u = zero_stuff(np.asarray(block, dtype=np.float32), self.U)
        if len(self.h) == 1:
            return (u * self.h[0]).astype(np.float32)
        y, self.zi = signal.lfilter(self.h, 1.0, u, zi=self.zi)
        return y.astype(np.float32)
