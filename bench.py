#!/usr/bin/env python
"""Headline benchmark: output Msamples/s of the batched streaming block-FIR path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path

Workload (BASELINE.json configs[1], "C2"): 1001-tap Remez low-pass at 176.4 kHz, stereo sessions
batched as independent streams, 512-frame blocks, TPDF dither to int24.  One step = one pass of the
hot path over one batch = BLOCKS_PER_STEP consecutive 512-frame blocks for every stream of the rank
(16384 streams per GPU by default).  Streams are partitioned across GPUs by stream id (weak
scaling, no collective on the data path); torch.distributed is used only for the barrier and the
max-over-ranks time.

`value`   : device-resident input and output (HBM), CUDA events around exactly K steps.
`e2e`     : same metric through the host API (engine.process_host_blocks): pinned host input,
            H2D + kernels + D2H inside the timed region.
`roofline`: the kernel is FP32-FMA bound (2*N flop per output sample); the peak is the
            FMA issue rate measured in this run with a packed-FFMA2 loop (MEASURED_PEAKS.json has no
            FP32 entry); HBM numbers are reported beside it.
`cpu_baseline`: the reference's block loop (sliding float32 buffer -> scipy oaconvolve 'valid' ->
            apply_dither, stream_process.py:96-113) as restated in oracle/, on all host cores, on a
            bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

# the CPU legs run one process per core; keep BLAS/OpenMP pools from oversubscribing them
for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
# stdout carries exactly one JSON line: keep NCCL's version / debug banner on stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "output Msamples/s (batched streams)"
N_TAPS, FS, FRAMES = 1001, 176400, 512
DEFAULT_STREAMS = 16384          # per GPU: 8192 stereo sessions
BLOCKS_PER_STEP = 16


def c2_taps():
    from audio_filtering_pipeline_b200 import design
    return design.create_fir_filter(method="remez", cutoff=20000, numtaps=N_TAPS, filter_type="lowpass",
                                    samplerate=FS)


def workload_config(streams, n_gpus, blocks):
    return {
        "workload": "C2: 1001-tap Remez lowpass @176.4kHz, stereo sessions as independent streams, "
                    "512-frame blocks, TPDF dither -> int24",
        "taps": N_TAPS, "upsample": 1, "frames_per_block": FRAMES, "blocks_per_step": blocks,
        "streams_per_gpu": streams, "streams_total": streams * n_gpus, "out_format": "int24 in int32",
        "dither": "tpdf philox4x32-10", "parallelism": f"stream-sharded x{n_gpus}, no collective",
        "l2_policy": "inputs larger than L2: each step streams blocks_per_step distinct input blocks "
                     f"({streams * FRAMES * 4 * blocks / 2**20:.0f} MiB in, same out) through HBM",
    }


# --------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """Reference block loop for `n_streams` streams x `n_blocks` blocks on one core."""
    taps, n_streams, n_blocks, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.reference_chain import SlidingBlockFir, apply_dither_scale
    rng = np.random.default_rng(seed)
    x = (0.1 * rng.standard_normal((n_blocks, FRAMES))).astype(np.float32)
    filt = [SlidingBlockFir(taps, FRAMES, 1) for _ in range(n_streams)]
    out = np.empty(FRAMES, dtype=np.float32)
    t0 = time.perf_counter()
    for b in range(n_blocks):
        for f in filt:
            y = f.step(x[b], FRAMES)                                        # stream_process.py:96-107
            out[:] = apply_dither_scale(y, np.random.random(FRAMES), 24)    # :61-64, :113
    return time.perf_counter() - t0


def cpu_reference(taps, seconds_target=12.0, pool=None, cores=None):
    """Times the oracle port of the reference loop on all host cores.  Returns (Msamples/s, cores,
    sample description)."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("fork").Pool(cores)
    try:
        # calibrate: 4 streams x 8 blocks per core
        t = max(pool.map(_cpu_worker, [(taps, 4, 8, i) for i in range(cores)]))
        per_sb = t / 32.0
        n_streams = 8
        n_blocks = int(max(8, min(4096, seconds_target / (per_sb * n_streams))))
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(taps, n_streams, n_blocks, 100 + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        if own_pool:
            pool.close()
            pool.join()
    samples = cores * n_streams * n_blocks * FRAMES
    return samples / wall / 1e6, cores, f"{cores * n_streams} streams x {n_blocks} blocks of {FRAMES} frames"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    reference's callback modules need sounddevice/soxr/matplotlib and do not import headless)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    taps = c2_taps()
    cores = len(os.sched_getaffinity(0))
    pool = mp.get_context("fork").Pool(cores)
    # per core per step: a bounded sample of the workload, long enough (~0.2 s) that the pool dispatch and the
    # per-call set-up inside the worker stay a few per cent of the step - the reference is not undersold
    n_streams, n_blocks = 8, 256
    job = [(taps, n_streams, n_blocks, i) for i in range(cores)]
    for _ in range(args.warmup):
        pool.map(_cpu_worker, job)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.map(_cpu_worker, job)
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    samples = args.steps * cores * n_streams * n_blocks * FRAMES
    value = samples / wall / 1e6
    sample = f"{cores * n_streams} streams x {n_blocks} blocks of {FRAMES} frames per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(wall / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.streams, args.gpus, args.blocks),
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        mhz = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": mhz, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs NVML reports as local to GPU `index`, so that the pinned
    staging buffers it allocates next are first-touched on that GPU's NUMA node (multi-GPU e2e runs are bound
    by host memory / PCIe root traffic).  Returns (previous affinity, description) - the caller restores the
    affinity before the CPU-baseline leg."""
    prev = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (max(prev) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & prev
        if cpus and cpus != prev:
            os.sched_setaffinity(0, cpus)
            return prev, f"{len(cpus)} of {len(prev)} cpus local to gpu {index}"
        return prev, "gpu-local cpu set == process cpu set"
    except Exception as exc:          # noqa: BLE001 - best effort only
        return prev, f"unavailable ({type(exc).__name__})"


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from audio_filtering_pipeline_b200.engine import BatchFirEngine, PinnedArray, measure_fma_peak
    from audio_filtering_pipeline_b200.sharding import max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    prev_affinity, numa_note = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    S, NB, K, W = args.streams, args.blocks, args.steps, args.warmup
    taps = c2_taps()
    eng = BatchFirEngine(taps, n_streams=S, max_frames=FRAMES, upsample=1, chain="dac", out_format="i24",
                         dither="tpdf", seed=2024, device=local, first_stream_id=rank * S)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic input, resident in HBM: NB distinct blocks [NB][S][FRAMES], sigma = 0.1 noise
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.empty((NB, S, FRAMES), device=dev, dtype=torch.float32)
    x.normal_(0.0, 0.1, generator=g).clamp_(-1.0, 1.0)
    y = torch.empty((NB, S, FRAMES), device=dev, dtype=torch.int32)

    def step():
        for b in range(NB):
            eng.process(x[b], y[b])

    fma_peak_tflops, fma_mhz = measure_fma_peak(local)
    for _ in range(max(W, 3)):
        step()
    barrier()
    l0 = eng.kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.kernel_launches - l0
    ms_total = e0.elapsed_time(e1)
    ms_total_max = max_over_ranks(ms_total, dev)
    samples_rank = S * FRAMES * NB * K
    value = samples_rank * world / (ms_total_max * 1e-3) / 1e6

    # roofline of the dominant (only) kernel, from this rank's own launches
    ms_launch = ms_total / launches
    flops_launch = 2.0 * N_TAPS * S * FRAMES
    bytes_launch = S * FRAMES * (4 + 4)
    achieved_tf = flops_launch / (ms_launch * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("c2_dram_bytes_per_launch")
        except Exception:
            traffic = None
    info = eng.stage_info()[-1]
    tensor = None
    if info.get("kernel") == 1:
        # tcgen05 path: MMA work actually issued per launch (3xTF32: x_hi*[W_hi|W_lo] as one N=128 MMA +
        # x_lo*W_hi as one N=64 MMA per K=8 step, 64-sample Toeplitz chunks), against the nominal dense TF32 peak
        R = 64
        n_chunks = (N_TAPS + R - 2) // R + 1
        items = (S + 15) // 16
        mma_flops = 2.0 * items * n_chunks * (R // 8) * (FRAMES // (8 * R)) * (128 * 2 * R * 8 + 128 * R * 8)
        tensor = {"kind": "tcgen05.mma kind::tf32, FP32 accumulators in TMEM, 3xTF32 split",
                  "issued_tflops": round(mma_flops / (ms_launch * 1e-3) / 1e12, 1), "nominal_tf32_dense_peak_tflops": 1125.0,
                  "note": "roofline.frac stays relative to the FP32-FMA roof of the north star (2N flop per sample); "
                          "it can exceed 1 because the work runs on the tensor pipe"}
        kernel_name = "afp::fir_tc_kernel<EPI_I24_TPDF, 1, 64> (tcgen05 / TMEM Toeplitz FIR)"
    else:
        kernel_name = ("afp::fir_row_kernel<1,16,256> (packed FFMA2, single row)" if info.get("kernel") == 2
                       else "afp::fir_stream_kernel<1,16,256> (packed FFMA2, row pair)")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp32_fma", "achieved": round(achieved_tf, 2), "peak": round(fma_peak_tflops, 2),
        "unit": "TFLOP/s", "frac": round(achieved_tf / fma_peak_tflops, 4), "traffic": traffic,
        "peak_source": f"packed-FFMA2 issue loop measured in this run at {fma_mhz:.0f} MHz "
                       "(MEASURED_PEAKS.json has no FP32 entry; nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4)",
        "kernel": kernel_name, "flop_per_sample": 2 * N_TAPS,
        "ms_per_launch": round(ms_launch, 5),
        "tensor": tensor,
        "hbm": {"achieved_gbs": round(bytes_launch / (ms_launch * 1e-3) / 1e9, 1), "peak_gbs": hbm_peak,
                "frac": round(bytes_launch / (ms_launch * 1e-3) / 1e9 / hbm_peak, 4),
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }

    # end to end through the host API: pinned host buffers, H2D + kernels + D2H in the timed region
    hx = PinnedArray((NB, S, FRAMES), np.float32)
    hy = PinnedArray((NB, S, FRAMES), np.int32)
    hx.a[...] = x.cpu().numpy()
    eng.reset()
    Ke = max(2, min(K, args.e2e_steps))
    for _ in range(2):
        eng.process_host_blocks(hx.a, hy.a)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        eng.process_host_blocks(hx.a, hy.a)
    torch.cuda.synchronize()
    te = max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = S * FRAMES * NB * Ke * world / te / 1e6
    e2e = {"value": round(e2e_value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": S * FRAMES * NB * 4,
           "d2h_bytes_per_step": S * FRAMES * NB * 4, "steps": Ke,
           "api": "BatchFirEngine.process_host_blocks (afp_process_host_blocks), pinned host memory"}
    checksum = int(np.bitwise_xor.reduce(hy.a.ravel()[:: 4099]))

    line = None
    os.sched_setaffinity(0, prev_affinity)      # the CPU legs use every core the process was given
    e2e["host_affinity"] = numa_note
    if rank == 0:
        sweep = agc = None
        if world == 1 and not args.no_sweep:
            del x, y
            torch.cuda.empty_cache()
            sweep = c5_sweep(local, hbm_peak, fma_peak_tflops)
            agc = agc_points(local, hbm_peak)
        cpu_val, cores, sample = cpu_reference(taps, seconds_target=args.cpu_seconds)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Msamples/s", "n_gpus": world,
            "steps": K, "warmup": max(W, 3), "ms_per_step": round(ms_total_max / K, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(S, world, NB),
            "roofline": roofline,
            "cpu_baseline": {"value": round(cpu_val, 3), "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "stage": eng.stage_info(), "checksum": checksum,
            "c5_sweep_kernel_only": sweep, "agc_kernel_only": agc,
        }
        print(json.dumps(line))
    hx.free()
    hy.free()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def c5_sweep(local, hbm_peak_gbs, fma_peak_tflops, iters=10):
    """Kernel-only points of BASELINE.json configs[4] (batch sweep: {51,101,251,1001} taps x {1,2,4}x up-sampling,
    f32 out, 1024-frame blocks - 512 for the 1001-tap U=1 case, the tensor kernel's block), device-resident I/O rotated over 3 buffer pairs (0.8-2 GB, far beyond L2),
    CUDA events.  Work per output sample: 2*ceil(N/U) flop, 4/U + 4 bytes; the relevant roof
    is the lower of the two."""
    import torch
    from audio_filtering_pipeline_b200.engine import BatchFirEngine
    dev = torch.device("cuda", local)
    out = []
    for n_taps, up, streams, B in [(51, 1, 65536, 1024), (51, 2, 32768, 1024), (51, 4, 32768, 1024),
                                   (101, 1, 32768, 1024), (101, 2, 32768, 1024), (101, 4, 32768, 1024),
                                   (251, 1, 16384, 1024), (251, 2, 16384, 1024), (251, 4, 16384, 1024),
                                   (1001, 1, 16384, 512), (1001, 2, 8192, 1024), (1001, 4, 8192, 1024)]:
        rng = np.random.default_rng(n_taps * 8 + up)
        h = up * rng.standard_normal(n_taps) / n_taps
        eng = BatchFirEngine(h, n_streams=streams, max_frames=B, upsample=up, device=local)
        xs = [0.1 * torch.randn(streams, B, device=dev) for _ in range(3)]
        ys = [torch.empty(streams, B * up, device=dev) for _ in range(3)]
        for i in range(3):
            eng.process(xs[i], ys[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            eng.process(xs[i % 3], ys[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        n_out = streams * B * up
        gs = n_out / ms / 1e6
        flop, byts = 2.0 * -(-n_taps // up), 4.0 / up + 4.0
        roof_fma, roof_hbm = fma_peak_tflops * 1e3 / flop, hbm_peak_gbs / byts     # Gsamples/s
        bound = "hbm" if roof_hbm < roof_fma else "fp32_fma"
        info = eng.stage_info()[-1]
        out.append({"taps": n_taps, "upsample": up, "streams": streams, "frames": B, "gsamples_s": round(gs, 1),
                    "tflops": round(gs * flop / 1e3, 1), "alg_gbs": round(gs * byts, 0), "bound": bound,
                    "frac": round(gs / min(roof_fma, roof_hbm), 3), "kernel": info["kernel"], "tile": info["tile_taps"]})
        eng.close()
        del xs, ys
    return out


def agc_points(local, hbm_peak_gbs, iters=10):
    """Kernel-only points of the AGC kernel (csrc/afp_agc.cuh): HBM-bound by design,
    4 bytes in + 4 bytes out per sample; 3 rotating input buffers (0.2-0.8 GB, beyond L2), CUDA events."""
    import torch
    from audio_filtering_pipeline_b200.engine import BatchFirEngine
    dev = torch.device("cuda", local)
    out = []
    with BatchFirEngine([1.0], n_streams=1, max_frames=1, device=local) as eng:
        for streams, frames, window in [(16384, 1024, 512), (65536, 1024, 512), (16384, 512, 256)]:
            xs = [0.05 * torch.randn(streams, frames, device=dev) for _ in range(3)]
            y = torch.empty_like(xs[0])
            for i in range(3):
                eng.agc_block(xs[i], out=y, window_size=window)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                eng.agc_block(xs[i % 3], out=y, window_size=window)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            gs = streams * frames / ms / 1e6
            out.append({"streams": streams, "frames": frames, "window": window, "gsamples_s": round(gs, 1),
                        "alg_gbs": round(gs * 8, 0), "bound": "hbm", "frac": round(gs * 8 / hbm_peak_gbs, 3)})
            del xs, y
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=DEFAULT_STREAMS, help="streams per GPU")
    ap.add_argument("--blocks", type=int, default=BLOCKS_PER_STEP, help="512-frame blocks per step")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-sweep", action="store_true", help="skip the short-filter kernel-only sweep (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
