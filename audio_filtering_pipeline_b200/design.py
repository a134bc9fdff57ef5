"""Host-side FIR design front end (numpy/scipy) - the part of the path that stays on the CPU.

Drop-in for the reference's design API: same names, argument order, defaults, return dtype
(float64[numtaps]) and error types.

    create_fir_filter   <- /root/reference/fir_filter.py:5-55
    design_fir_filter   <- /root/reference/filter_methods.py:6-99
    get_window          <- /root/reference/window_types.py:5-53
    is_symmetric        <- /root/reference/stream_process_EQ_GUI.py:57-58
    normalize_filter    <- /root/reference/stream_process_EQ_GUI.py:60-63
    minimum-phase gate  <- /root/reference/stream_process_EQ_GUI.py:377-379

Two reference quirks are kept on purpose because they change the taps that reach
the device: windows are *periodic* (scipy's ``fftbins=True`` default), so windowed designs are
slightly asymmetric; ``minimum_phase(..., 'hilbert')`` halves the length.  One reference defect is
repaired: ``method='remez'`` raises NameError upstream (``filter_methods.py:96`` never imports
``remez``); here it calls ``scipy.signal.remez`` with the reference's own band-edge rules.

Nothing in this module touches the GPU: the engine converts the returned float64 taps to the
float32 tables the kernels read (see ``engine.py``).
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _sig

__all__ = [
    "get_window",
    "design_fir_filter",
    "create_fir_filter",
    "is_symmetric",
    "normalize_filter",
    "apply_min_phase",
    "design_upsampler",
    "EQ_BANDS_9",
    "EQ_BANDS_10",
    "design_eq_bank",
    "combine_eq_taps",
]

_KAISER_BETA = 8.6  # window_types.py:39

# The nine GUI bands (stream_process_EQ_GUI.py:162-172) and the builder-defined tenth band that
# BASELINE.json's "10-band" config needs.
EQ_BANDS_9 = [
    (20, 60), (60, 120), (120, 250), (250, 500), (500, 1000),
    (1000, 2000), (2000, 4000), (4000, 8000), (8000, 16000),
]
EQ_BANDS_10 = EQ_BANDS_9 + [(16000, 20000)]


def get_window(name, numtaps):
    """Window values for ``name`` (periodic form, as the reference: window_types.py:33-45).

    'kaiser' (any case) means beta = 8.6.  Every failure - unknown name, a tuple passed as the
    name, bad length - surfaces as ValueError, like window_types.py:47-53.
    """
    try:
        key = name.lower()  # a non-str name fails here, inside the try, like the reference
        spec = ("kaiser", _KAISER_BETA) if key == "kaiser" else name
        return _sig.get_window(spec, numtaps)
    except Exception as exc:  # noqa: BLE001 - the reference folds everything into ValueError
        raise ValueError(
            f"Window type '{name}' not recognized or invalid parameters. "
            f"Supported types: boxcar, hamming, hann, blackman, kaiser, "
            f"bartlett, flattop, nuttall, etc. Original error: {exc}"
        )


def _ideal_response(kind, fc, t):
    """Ideal (unwindowed) impulse response on the centred grid ``t``; ``fc`` is Nyquist-normalised
    (filter_methods.py:35-50)."""
    if kind == "lowpass":
        return fc * np.sinc(fc * t)
    if kind == "highpass":
        return np.sinc(t) - fc * np.sinc(fc * t)
    if kind == "bandpass":
        return fc[1] * np.sinc(fc[1] * t) - fc[0] * np.sinc(fc[0] * t)
    if kind == "bandstop":
        return np.sinc(t) - (fc[1] * np.sinc(fc[1] * t) - fc[0] * np.sinc(fc[0] * t))
    raise ValueError("Invalid filter type")


def _remez_bands(kind, cutoff, fs):
    """Band edges / desired gains for the equiripple branch (filter_methods.py:78-95)."""
    nyq = fs / 2
    if kind == "lowpass":
        width = min(1000, (nyq - cutoff) / 2)
        upper = min(nyq, cutoff + width)
        if cutoff >= upper:
            raise ValueError("Cutoff frequency too close to Nyquist for given transition width.")
        return [0, cutoff, upper, nyq], [1, 0]
    if kind == "highpass":
        width = min(1000, cutoff / 2)
        lower = max(0, cutoff - width)
        return [0, lower, cutoff, nyq], [0, 1]
    if kind == "bandpass":
        return [0, cutoff[0], cutoff[1], nyq], [0, 1, 0]
    if kind == "bandstop":
        return [0, cutoff[0], cutoff[1], nyq], [1, 0, 1]
    # the reference falls through to an UnboundLocalError here; ValueError is the closest sane type
    raise ValueError("Invalid filter type")


def design_fir_filter(method="window", cutoff=None, numtaps=101, window="hamming",
                      filter_type="lowpass", samplerate=44100):
    """Core designer (filter_methods.py:6-99).  ``cutoff`` is in Hz: a scalar for low/high-pass,
    ``[lo, hi]`` for band filters.  A non-string ``window`` means "no window" (``:30-33``)."""
    if method == "window":
        if cutoff is None:
            raise ValueError("Cutoff frequency must be provided.")
        # Centred time grid.  The reference nudges t==0 to 1e-20 before calling np.sinc; sinc(0)
        # and cos(0) give bit-identical values, so no nudge is needed here.
        t = np.arange(numtaps) - (numtaps - 1) / 2
        cutoff = np.asarray(cutoff)
        fc = cutoff / (samplerate / 2)
        win = get_window(window, numtaps) if isinstance(window, str) else np.ones(numtaps)

        h = _ideal_response(filter_type, fc, t) * win

        if filter_type == "lowpass":          # unity gain at DC (:58-59)
            h = h / np.sum(h)
        elif filter_type == "highpass":       # unity gain at Nyquist (:60-62)
            h = h / np.sum(h * np.cos(2 * np.pi * 0.5 * t))
        elif filter_type == "bandpass":       # unity gain at the geometric-mean centre (:63-68)
            w, resp = _sig.freqz(h, worN=8000, fs=samplerate)
            centre = np.sqrt(cutoff[0] * cutoff[1])
            h = h / np.abs(resp[np.argmin(np.abs(w - centre))])
        else:                                 # bandstop: mean gain at the two edges (:69-74)
            w, resp = _sig.freqz(h, worN=8000, fs=samplerate)
            g_lo = np.abs(resp[np.argmin(np.abs(w - cutoff[0]))])
            g_hi = np.abs(resp[np.argmin(np.abs(w - cutoff[1]))])
            h = h / ((g_lo + g_hi) / 2)
        return h

    if method == "remez":
        bands, desired = _remez_bands(filter_type, cutoff, samplerate)
        return _sig.remez(numtaps, bands, desired, fs=samplerate)

    raise NotImplementedError(f"Method '{method}' not implemented.")


def create_fir_filter(method="window", cutoff=0.3, numtaps=101, window_type="hamming",
                      filter_type="lowpass", samplerate=44100):
    """Design front end with the reference's signature and defaults (fir_filter.py:5-7).

    Like the reference (``:38-43``) the window is evaluated up front for ``method='window'`` - its
    only effect is to raise ValueError early for a bad window name - and the window *name* is what
    reaches the designer (``:46-53``).  Callers pass ``samplerate = fs * upsample_factor``
    (stream_process.py:39).
    """
    if method == "window":
        get_window(window_type, numtaps)
    return design_fir_filter(method=method, cutoff=cutoff, numtaps=numtaps, window=window_type,
                             filter_type=filter_type, samplerate=samplerate)


def is_symmetric(h, tol=1e-8):
    """Linear-phase test used to gate the minimum-phase transform (stream_process_EQ_GUI.py:57-58)."""
    h = np.asarray(h)
    return bool(np.allclose(h, h[::-1], atol=tol))


def normalize_filter(h, fs):
    """Scale to max|H| = 1 on the 8000-point grid (stream_process_EQ_GUI.py:60-63)."""
    _, resp = _sig.freqz(h, worN=8000, fs=fs)
    peak = np.max(np.abs(resp))
    return h / peak if peak != 0 else h


def apply_min_phase(h, fs, min_phase=True, force=False, normalize=True):
    """The GUI "Minimum Phase" switch (stream_process_EQ_GUI.py:377-379; EQ-band variant without
    the normalisation: ``:217-218`` -> ``normalize=False``).

    With ``force=False`` this is the reference gate: the transform only happens when the taps pass
    ``is_symmetric`` - which, because windows are periodic, is true only for 'boxcar' and Remez
    designs.  ``force=True`` skips the gate (builder extension).  The result has
    ``(len(h)+1)//2`` taps with the energy at index 0; on the device it runs through the ordinary
    causal FIR kernel (there is no linear-phase shortcut to lose).
    """
    h = np.asarray(h, dtype=np.float64)
    if not min_phase:
        return h
    if not (force or is_symmetric(h)):
        return h
    h = _sig.minimum_phase(h, method="hilbert")
    return normalize_filter(h, fs) if normalize else h


# libsoxr's quality recipes (soxr.h / soxr_quality_spec of libsoxr 0.1.3, the library behind the reference's
# soxr.resample(..., 'VHQ'), stream_process.py:66-77 - restated from its published source, not runnable here):
# 'HQ' = 20-bit, 'VHQ' = 28-bit precision; stop-band rejection = bits * 6.02 dB; the stop band begins AT the Nyquist
# frequency of the lower rate; the pass band ends at 1 - 0.05 / TO_3dB(rejection) of it, TO_3dB(a) = (1.6e-6 a - 7.5e-4) a + 0.646.
_SOXR_BITS = {"hq": 20, "vhq": 28}


def upsampler_spec(quality):
    """(stop-band rejection in dB, pass-band end as a fraction of the base-rate Nyquist frequency) of a soxr recipe."""
    bits = _SOXR_BITS[quality]
    rej = bits * 20.0 * np.log10(2.0)
    return rej, 1.0 - 0.05 / ((1.6e-6 * rej - 7.5e-4) * rej + 0.646)


def design_upsampler(factor, samplerate, numtaps=None, window_type="kaiser", quality="fast"):
    """Anti-imaging prototype for the polyphase interpolator that replaces the per-block
    ``soxr.resample`` call (stream_process.py:66-77), scaled by ``factor`` so the passband gain after
    zero-stuffing is 1.  ``factor == 1`` gives the identity ``[1.0]``.  Builder-defined (soxr parity is
    unpinned); what each choice guarantees is tested in tests/test_design.py:

    ``quality='fast'`` (default): 32*factor+1 taps, Kaiser windowed-sinc low-pass at ``samplerate/2`` from the
        reference's own designer: flat to 0.001 dB up to 0.8 Nyquist, -0.5 dB at 0.91 Nyquist, -6 dB at Nyquist,
        images beyond 1.2 Nyquist below -88 dB.  Short on purpose (a live callback's latency: 16 base-rate samples).
    ``quality='hq'`` / ``'vhq'``: Kaiser designs to libsoxr's HQ / VHQ recipes (``upsampler_spec``): pass band flat
        (< 1e-4 dB) to 0.913 Nyquist, stop band from Nyquist on at -120 dB (365 / 727 taps for 2x / 4x) or -160 dB
        (507 / 1013 taps; -150 dB with the float32 taps the device holds).  In the collapsed reference chain the extra
        taps cost a quarter of their number at 4x: (1013 + N - 1) / 4 base-rate taps.
    """
    factor = int(factor)
    if factor <= 1:
        return np.ones(1)
    if quality != "fast":
        from scipy import signal
        rej, pass_end = upsampler_spec(quality)
        nyq, rate = samplerate / 2.0, float(samplerate * factor)
        n, beta = signal.kaiserord(rej, (1.0 - pass_end) * nyq / (0.5 * rate))
        n = int(numtaps) if numtaps is not None else (n | 1)
        return factor * signal.firwin(n, 0.5 * (pass_end + 1.0) * nyq, window=("kaiser", beta), fs=rate)
    if numtaps is None:
        numtaps = 32 * factor + 1
    h = create_fir_filter(method="window", cutoff=samplerate / 2, numtaps=numtaps,
                          window_type=window_type, filter_type="lowpass",
                          samplerate=samplerate * factor)
    return h * factor


def design_eq_bank(numtaps, samplerate, bands=None, min_phase=False):
    """The EQ band filters (precompute_eq_filters, stream_process_EQ_GUI.py:204-221): Hamming
    band-pass per band at the processing rate, optional gated minimum phase (no normalisation)."""
    bands = EQ_BANDS_9 if bands is None else bands
    out = []
    for lo_hi in bands:
        c = create_fir_filter(method="window", cutoff=lo_hi, numtaps=numtaps,
                              window_type="hamming", filter_type="bandpass", samplerate=samplerate)
        if min_phase:
            c = apply_min_phase(c, samplerate, normalize=False)
        out.append(c)
    return out


def combine_eq_taps(band_taps, gains):
    """Taps-domain EQ sum ``sum_b g_b * h_b`` (stream_process_EQ_GUI.py:86-90).  Bands of unequal
    length are aligned on their 'same' centres ``(N_b-1)//2`` - where scipy centres each band's output
    in stream_process_EQ.py:128 - which is also how the engine lays out its EQ tables."""
    cmax = max((len(c) - 1) // 2 for c in band_taps)
    n = max(cmax - (len(c) - 1) // 2 + len(c) for c in band_taps)
    acc = np.zeros(n)
    for c, g in zip(band_taps, gains):
        c = np.asarray(c, dtype=np.float64)
        off = cmax - (len(c) - 1) // 2
        acc[off:off + len(c)] += float(g) * c
    return acc
