#!/usr/bin/env python
"""End-to-end (host buffers) throughput of the C2 workload through BatchFirEngine.process_host_blocks: pinned host
input -> H2D -> kernels -> D2H, wall clock.

    python bench/e2e_probe.py [--steps 8] [--fmt i24|i24_packed|i16] [--wc]

--fmt  output sample format: int24 in an int32 container (4 B per sample back over PCIe), packed 3-byte int24, int16.
--wc   input staging in write-combined pinned memory.
AFP_HOST_CHUNK_MB selects the pipeline-unit size (tuning experiments).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine, PinnedArray  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--fmt", default="i24")
ap.add_argument("--wc", action="store_true")
ap.add_argument("--blocks", type=int, default=B.BLOCKS_PER_STEP)
a = ap.parse_args()
S, NB, F = B.DEFAULT_STREAMS, a.blocks, B.FRAMES
eng = BatchFirEngine(B.c2_taps(), n_streams=S, max_frames=F, out_format=a.fmt, dither="tpdf", seed=2024)
out_shape, out_dtype = {"i24": ((NB, S, F), np.int32), "i24_packed": ((NB, S, 3 * F), np.uint8), "i16": ((NB, S, F), np.int16)}[a.fmt]
hx, hy = PinnedArray((NB, S, F), np.float32, write_combined=a.wc), PinnedArray(out_shape, out_dtype)
hx.a[...] = 0.1 * np.random.default_rng(0).standard_normal((NB, S, F)).astype(np.float32)
for _ in range(2):
    eng.process_host_blocks(hx.a, hy.a)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    eng.process_host_blocks(hx.a, hy.a)
torch.cuda.synchronize()
t = time.perf_counter() - t0
rate = S * F * NB * a.steps / t
print(json.dumps({"fmt": a.fmt, "write_combined_input": a.wc, "blocks": NB, "chunk_mb": os.environ.get("AFP_HOST_CHUNK_MB", "default"),
                  "e2e_msamples_s": round(rate / 1e6, 1), "h2d_gbs": round(rate * 4 / 1e9, 1),
                  "d2h_gbs": round(rate * hy.a.itemsize * hy.a.shape[2] / F / 1e9, 1), "ms_per_step": round(t / a.steps * 1e3, 3),
                  "checksum": int(np.bitwise_xor.reduce(hy.a.ravel()[::4099].astype(np.int64)))}))
