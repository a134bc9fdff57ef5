#!/usr/bin/env python
"""Kernel-only throughput probe for one configuration (device-resident I/O, CUDA events).

    python bench/quick_perf.py --taps 1001 --up 1 --frames 512 --streams 16384 --fmt i24 --dither tpdf
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_filtering_pipeline_b200.engine import BatchFirEngine, measure_fma_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=1001)
    ap.add_argument("--up", type=int, default=1)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--fmt", default="f32")
    ap.add_argument("--dither", default="none")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--peak", action="store_true")
    ap.add_argument("--chain", default="dac", choices=["dac", "reference"],
                    help="reference: interpolate (129-tap Kaiser prototype) -> FIR -> keep every U-th sample")
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    h = rng.standard_normal(a.taps) / a.taps
    S, B, U = a.streams, a.frames, a.up
    kw = {}
    if a.chain == "reference" and U > 1:
        from audio_filtering_pipeline_b200 import design
        kw = dict(chain="reference",
                  upsampler_taps=U * design.create_fir_filter(cutoff=20000, numtaps=129, window_type="kaiser",
                                                              samplerate=44100 * U))
    eng = BatchFirEngine(h, n_streams=S, max_frames=B, upsample=U, out_format=a.fmt, dither=a.dither, seed=1, **kw)
    xs = [0.1 * torch.randn(S, B, device="cuda") for _ in range(a.nbuf)]
    out = eng.process(xs[0])
    for i in range(3):
        eng.process(xs[i % a.nbuf], out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        eng.process(xs[i % a.nbuf], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    n_out = S * B * (1 if a.chain == "reference" else U)
    eb = {"f32": 4, "i16": 2, "i24": 4, "i24_packed": 3}[a.fmt]
    res = {
        "taps": a.taps, "up": U, "frames": B, "streams": S, "fmt": a.fmt, "dither": a.dither,
        "ms": round(ms, 4), "gsamples_s": round(n_out / ms / 1e6, 2),
        "tflops": round(2.0 * np.ceil(a.taps / U) * n_out / ms / 1e9, 2),
        "alg_gbs": round((S * B * 4 + n_out * eb) / ms / 1e6, 1),
        "stage": eng.stage_info(),
    }
    if a.peak:
        res["fma_peak_tflops"], res["sm_mhz"] = (round(v, 2) for v in measure_fma_peak(0))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
