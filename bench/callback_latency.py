#!/usr/bin/env python
"""Per-block latency of the drop-in sounddevice callback (compat.make_audio_callback) for C1: one mono
stream, 251-tap Kaiser band-pass at 4x (176.4 kHz), 1024-frame blocks, host numpy buffers in and out -
the reference's own call shape (stream_process.py:79-120).  Prints median / p99 microseconds per block and
the real-time budget (frames / fs).  The CPU leg times the oracle's statement of the reference loop."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_filtering_pipeline_b200.compat import StreamChain, make_audio_callback  # noqa: E402

fs, U, B, n_blocks = 44100, 4, 1024, 400
chain = StreamChain(samplerate=fs, upsample_factor=U, blocksize=B, numtaps=251, cutoff=[500, 8000],
                    window_type="kaiser", filter_type="bandpass", dither=True)
cb = make_audio_callback(chain, verbose=False)
rng = np.random.default_rng(0)
x = (0.1 * rng.standard_normal((n_blocks, B, 1))).astype(np.float32)
out = np.zeros((B, 1), dtype=np.float32)
for i in range(20):
    cb(x[i], out, B, None, None)
lat = []
for i in range(n_blocks):
    t0 = time.perf_counter()
    cb(x[i], out, B, None, None)
    lat.append((time.perf_counter() - t0) * 1e6)
lat = np.array(lat)
res = {"config": "C1: 251-tap Kaiser BP, 44.1 kHz mono, 4x, 1024-frame blocks, sounddevice-shaped callback, host buffers",
       "median_us": round(float(np.median(lat)), 1), "p99_us": round(float(np.percentile(lat, 99)), 1),
       "budget_us": round(B / fs * 1e6, 1), "stages": [(s["kernel"], s["n_taps"]) for s in chain.engine.stage_info()]}
if "--cpu" in sys.argv:
    from oracle.reference_chain import SlidingBlockFir, PolyphaseUpsampler
    from audio_filtering_pipeline_b200 import design
    up = PolyphaseUpsampler(design.design_upsampler(U, fs), U)
    fir = SlidingBlockFir(chain.fir_coeff, B, U)
    t = []
    for i in range(100):
        t0 = time.perf_counter()
        fir.step(up.step(x[i, :, 0]), B)
        t.append((time.perf_counter() - t0) * 1e6)
    res["cpu_reference_port_median_us"] = round(float(np.median(t)), 1)
print(json.dumps(res))
