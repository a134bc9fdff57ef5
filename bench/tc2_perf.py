"""C2 kernel-only timing of the split-FP16 tensor kernel: block-by-block launches vs one afp_process_blocks launch.
usage: python bench/tc2_perf.py [streams] [blocks] [taps] [frames]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from audio_filtering_pipeline_b200 import design
from audio_filtering_pipeline_b200.engine import BatchFirEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
B = int(sys.argv[4]) if len(sys.argv) > 4 else 512
h = design.create_fir_filter(method="remez", cutoff=20000, numtaps=N, filter_type="lowpass", samplerate=176400) if N == 1001 else \
    design.create_fir_filter(cutoff=9000, numtaps=N, window_type="kaiser", samplerate=96000)
dev = torch.device("cuda", 0)
x = torch.empty((NB, S, B), device=dev).normal_(0, 0.1).clamp_(-1, 1)
for fmt, dither in (("i24", "tpdf"), ("f32", "none")):
    y = torch.empty((NB, S, B), device=dev, dtype=torch.int32 if fmt == "i24" else torch.float32)
    with BatchFirEngine(h, n_streams=S, max_frames=B, out_format=fmt, dither=dither, seed=1) as eng:
        res = {"streams": S, "blocks": NB, "taps": N, "frames": B, "fmt": fmt}
        for mode in ("per_block", "multi"):
            def step():
                if mode == "multi":
                    eng.process_blocks(x, y)
                else:
                    for b in range(NB):
                        eng.process(x[b], y[b])
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            res[mode + "_ms"] = round(ms, 4)
            res[mode + "_gsamples_s"] = round(S * B * NB / ms / 1e6, 1)
        res["kernel"] = eng.stage_info()[-1]["kernel"]
        print(json.dumps(res), flush=True)
