"""Builds the CUDA shared library in-tree with nvcc for sm_100a (B200 only, no other arch).

    python -m audio_filtering_pipeline_b200._build            # build if stale
    python -m audio_filtering_pipeline_b200._build --force

The .so lands next to this file (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libafp_b200.so")
SOURCES = ["afp_abi.cu", "afp_kernels_u1.cu", "afp_kernels_u2.cu", "afp_kernels_u3.cu", "afp_kernels_u4.cu", "afp_kernels_tc.cu", "afp_kernels_tc2.cu",
           "afp_kernels_row_u1.cu", "afp_kernels_row_u2.cu", "afp_kernels_row_u3.cu", "afp_kernels_row_u4.cu",
           "afp_kernels_agc.cu", "afp_kernels_peak.cu", "afp_kernels_grouped.cu"]
DEPS = SOURCES + ["afp_fir_kernel.cuh", "afp_kernel_table.cuh", "afp_fir_tc.cuh", "afp_fir_tc2.cuh", "afp_philox.cuh", "afp_fir_row.cuh", "afp_kernels_row.cuh", "afp_agc.cuh", "afp_spectrum.cuh",
                  os.path.join("..", "..", "include", "afp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile libafp_b200.so if missing or older than its sources.  Returns the library path.
    The translation units (one per up-sampling factor + the ABI) are compiled in parallel."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
