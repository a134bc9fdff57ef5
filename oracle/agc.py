"""ORACLE - test infrastructure, not product code (see reference_chain.py header for who may import it).

CPU restatement of the reference's automatic gain control, ``apply_agc``
(stream_process_AGC.py:43-89) and of its call site in the worker loop (:131-143): per block and
stateless - moving-window RMS ('same' convolution of the squared block with a boxcar), desired gain
``target / rms`` limited to ``max_gain``, a first-order smoother whose coefficient switches between an
attack and a release constant on the sign of (desired - smoothed), a final [0.1, max_gain] limiter,
and at the call site a clip of the result to +-0.99.

Parity status: PINNED - ``tests/test_oracle_golden.py`` checks this file against
``tests/golden/agc_*.npz``, produced by executing the reference's own ``apply_agc``
(``tests/golden/make_golden.py::gen_agc``), including its error path (an attack or release constant of
zero samples divides by zero inside the loop; the reference catches the exception and returns the block
unchanged, :87-89).

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["apply_agc", "agc_stage", "smoothing_constants"]


def smoothing_constants(window_size, attack=0.01, release=0.1):
    """(attack_samples, release_samples, alpha_attack, alpha_release); an alpha is None when its
    sample count is 0 (stream_process_AGC.py:56-57,73-78: ``1 - exp(-1/0)`` raises)."""
    a_n = int(attack * window_size)                                   # :56
    r_n = int(release * window_size)                                  # :57
    a = 1.0 - math.exp(-1.0 / a_n) if a_n != 0 else None              # :73-75
    r = 1.0 - math.exp(-1.0 / r_n) if r_n != 0 else None              # :76-78
    return a_n, r_n, a, r


def apply_agc(signal, target_level=0.1, window_size=512, alpha=0.01, max_gain=10.0, attack=0.01,
              release=0.1):
    """One block, 1-D float32 in -> float64 out of the same length (the reference returns [n, 1]);
    on the reference's error path the input comes back unchanged (float32), as at :87-89.
    ``alpha`` is accepted and ignored, as in the reference."""
    x = np.asarray(signal)
    x = x.reshape(-1)
    n = len(x)
    a_n, r_n, a_att, a_rel = smoothing_constants(window_size, attack, release)
    squared = np.square(x)                                            # :60 (stays float32)
    window = np.ones(window_size) / window_size                       # :61
    rms = np.sqrt(np.convolve(squared, window, mode="same"))          # :62 (float64)
    if len(rms) != n:                                                 # window longer than the block:
        return x                                                      # the product at :85 cannot broadcast -> :87-89
    desired = np.clip(target_level / (rms + 1e-10), 0, max_gain)      # :65
    smoothed = np.zeros_like(desired)
    smoothed[0] = desired[0]                                          # :69
    for i in range(1, n):                                             # :71-80
        if desired[i] > smoothed[i - 1]:
            if a_att is None:
                return x                                              # ZeroDivisionError -> :87-89
            a = a_att
        else:
            if a_rel is None:
                return x
            a = a_rel
        smoothed[i] = a * desired[i] + (1 - a) * smoothed[i - 1]
    smoothed = np.clip(smoothed, 0.1, max_gain)                       # :83
    return x * smoothed                                               # :86 (float32 * float64 -> float64)


def agc_stage(block, target_level=0.1, blocksize=None, max_gain=10.0, attack=0.01, release=0.1,
              clip=0.99):
    """The call site (stream_process_AGC.py:131-143): window ``min(512, blocksize // 2)``, then the
    clip to +-0.99.  ``block`` 1-D; returns float64 (float32 on the pass-through path)."""
    n = len(block) if blocksize is None else blocksize
    y = apply_agc(block, target_level=target_level, window_size=min(512, n // 2), max_gain=max_gain,
                  attack=attack, release=release)
    return np.clip(y, -clip, clip)                                    # :143
