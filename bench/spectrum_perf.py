#!/usr/bin/env python
"""Kernel-only throughput of afp_spectrum_db (csrc/afp_spectrum.cuh): input samples per second, CUDA events,
rotating input buffers beyond L2.  Algorithmic bytes: 4 in + 2 out per input sample."""
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
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    S, n = a.streams, a.frames
    xs = [0.1 * torch.randn(S, n, device="cuda") for _ in range(4)]
    out = torch.empty(S, n // 2 + 1, device="cuda")
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as eng:
        for i in range(3):
            eng.spectrum_db(xs[i], out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            eng.spectrum_db(xs[i % 4], out=out)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    gbs = S * (n * 4 + (n // 2 + 1) * 4) / ms / 1e6
    print(json.dumps({"kernel": "spectrum_db", "streams": S, "frames": n, "ms": round(ms, 4),
                      "gsamples_s": round(S * n / ms / 1e6, 1), "alg_gbs": round(gbs, 1), "hbm_frac": round(gbs / HBM_GBS, 3),
                      "tflops_5nlog2n": round(5.0 * n * (n.bit_length() - 1) * S / ms / 1e9, 2)}))


if __name__ == "__main__":
    main()
