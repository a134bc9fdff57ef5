"""ORACLE - test infrastructure, not product code (see reference_chain.py header for who may import it).

CPU restatement of the waterfall row the reference computes from every output block when its display is on
(stream_process_GUI_Waterfall.py:471-476): Hann window, rfft, 20*log10(|X| + 1e-12).

Parity status: PINNED - tests/test_oracle_golden.py checks it against tests/golden/waterfall_spectrum.npz,
rows produced by the reference's own callback with show_waterfall on (make_golden.py::gen_waterfall_spectrum).
"""
from __future__ import annotations

import numpy as np

__all__ = ["spectrum_db"]


def spectrum_db(block):
    """block: 1-D float32 (the outdata column) -> float64 [len(block)//2 + 1]."""
    block = np.asarray(block)
    windowed = block * np.hanning(len(block))                        # :473
    spectrum = np.fft.rfft(windowed)                                 # :474
    return 20 * np.log10(np.abs(spectrum) + 1e-12)                   # :475
