#!/usr/bin/env python
"""Kernel-only throughput of the AGC kernel (csrc/afp_agc.cuh): device-resident blocks, CUDA events,
distinct input buffers larger than L2 in rotation.  HBM-bound by design: 4 bytes in + 4 bytes out
per sample.

    python bench/agc_perf.py --streams 16384 --frames 1024 --window 512
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402

HBM_GBS = 6531.9   # MEASURED_PEAKS.json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--window", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nbuf", type=int, default=4)
    a = ap.parse_args()
    S, n = a.streams, a.frames
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [(10.0 ** (-3.0 * torch.rand((S, 1), device="cuda", generator=g))) *
          torch.randn((S, n), device="cuda", generator=g) for _ in range(a.nbuf)]
    out = torch.empty_like(xs[0])
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as eng:
        for i in range(3):
            eng.agc_block(xs[i % a.nbuf], out=out, window_size=a.window)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            eng.agc_block(xs[i % a.nbuf], out=out, window_size=a.window)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    gbs = S * n * 8 / ms / 1e6
    print(json.dumps({"kernel": "agc", "streams": S, "frames": n, "window": a.window, "ms": round(ms, 4),
                      "gsamples_s": round(S * n / ms / 1e6, 1), "alg_gbs": round(gbs, 1),
                      "hbm_frac": round(gbs / HBM_GBS, 3)}))


if __name__ == "__main__":
    main()
