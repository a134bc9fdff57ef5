#!/usr/bin/env python
"""End-to-end (host buffers) throughput of the C2 workload through BatchFirEngine.process_host_blocks for
the pipeline-unit size in AFP_HOST_CHUNK_MB: pinned host input -> H2D -> kernels -> D2H, wall clock."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine, PinnedArray  # noqa: E402

S, NB, F = B.DEFAULT_STREAMS, B.BLOCKS_PER_STEP, B.FRAMES
eng = BatchFirEngine(B.c2_taps(), n_streams=S, max_frames=F, out_format="i24", dither="tpdf", seed=2024)
hx, hy = PinnedArray((NB, S, F), np.float32), PinnedArray((NB, S, F), np.int32)
hx.a[...] = 0.1 * np.random.default_rng(0).standard_normal((NB, S, F)).astype(np.float32)
for _ in range(2):
    eng.process_host_blocks(hx.a, hy.a)
torch.cuda.synchronize()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.perf_counter()
for _ in range(steps):
    eng.process_host_blocks(hx.a, hy.a)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print(json.dumps({"chunk_mb": os.environ.get("AFP_HOST_CHUNK_MB", "default"), "e2e_msamples_s": round(S * F * NB * steps / t / 1e6, 1),
                  "ms_per_step": round(t / steps * 1e3, 3), "checksum": int(np.bitwise_xor.reduce(hy.a.ravel()[::4099]))}))
