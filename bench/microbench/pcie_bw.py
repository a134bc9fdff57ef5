#!/usr/bin/env python
"""Host<->device copy ceilings with pinned memory: H2D alone, D2H alone, both at once."""
import json
import time
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


res = {"bytes": n, "h2d_gbs": n / timed(h2d) / 1e9, "d2h_gbs": n / timed(d2h) / 1e9}
t = timed(both)
res["bidir_each_gbs"] = n / t / 1e9
for chunk in (1 << 20, 4 << 20, 32 << 20):
    def chunked():
        with torch.cuda.stream(s1):
            for o in range(0, n, chunk):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
    res[f"h2d_chunk_{chunk >> 20}MiB_gbs"] = n / timed(chunked, 3) / 1e9
print(json.dumps(res))
