#!/usr/bin/env python
"""C3 (BASELINE.json configs[2]): 10-band EQ bank (fused form) + main low-pass, 48 kHz, batched stereo
sessions, 1024-frame blocks; kernel-only, device-resident.  AFP_NO_COLLAPSE=1 runs EQ and main FIR as two
launches instead of one filter eq (*) main."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_filtering_pipeline_b200 import design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402

n_eq, n_main, S, B = (int(v) for v in (sys.argv[1:5] + ["501", "301", "8192", "1024"][len(sys.argv) - 1:]))
fs = 48000
bank = design.design_eq_bank(n_eq, fs, design.EQ_BANDS_10)
gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3, 0.9]
main = design.create_fir_filter(cutoff=16000, numtaps=n_main, samplerate=fs)
eng = BatchFirEngine(main, n_streams=S, max_frames=B, eq_bands=bank, eq_gains=gains, eq_mode="fused")
xs = [0.05 * torch.randn(S, B, device="cuda") for _ in range(3)]
out = eng.process(xs[0])
for i in range(3):
    eng.process(xs[i], out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    eng.process(xs[i % 3], out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"n_eq": n_eq, "n_main": n_main, "streams": S, "frames": B, "collapse": not os.environ.get("AFP_NO_COLLAPSE"),
                  "gsamples_s": round(S * B / ms / 1e6, 2), "ms": round(ms, 4),
                  "stages": [(s["kernel"], s["n_taps"]) for s in eng.stage_info()]}))
