"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/afp.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from audio_filtering_pipeline_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "afp.h")).read()
    declared = set(re.findall(r"\b(afp_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libafp_b200.so"


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.Config) == 56          # 8 x int32, uint64, 4 x int32 (the last: same_buffer_len)
    assert _lib.Config.seed.offset == 32
    assert C.sizeof(_lib.StageInfo) == 44


def test_version_string(lib):
    assert b"sm_100a" in lib.afp_version()


def test_no_cpu_fallback(lib):
    """On a box without a CUDA device afp_create must fail with a message, not compute on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    ctx = C.c_void_p()
    cfg = _lib.Config(device=0, n_streams=1, max_frames=64, upsample=1)
    rc = lib.afp_create(C.byref(cfg), C.byref(ctx))
    assert rc != 0 and not ctx
    assert lib.afp_last_error(None)
    from audio_filtering_pipeline_b200.engine import BatchFirEngine
    with pytest.raises(_lib.AfpError):
        BatchFirEngine([1.0])
    # the reference-shaped helpers keep the reference's "zeros / unchanged block on a processing error"
    # convention, but a missing device is not a processing error: it raises
    import numpy as np
    from audio_filtering_pipeline_b200 import compat
    blk = np.zeros(64, dtype=np.float32)
    for call in (lambda: compat.apply_agc(blk), lambda: compat.safe_upsample(blk, 44100, 88200),
                 lambda: compat.safe_downsample(blk, 88200, 44100, 32), lambda: compat.apply_dither(blk)):
        with pytest.raises(_lib.AfpError):
            call()


def test_argument_validation_needs_no_gpu(lib):
    ctx = C.c_void_p()
    for bad in (dict(n_streams=0), dict(upsample=5), dict(out_format=9), dict(chain=3), dict(dither=4)):
        kw = dict(device=0, n_streams=1, max_frames=64, upsample=1)
        kw.update(bad)
        cfg = _lib.Config(**kw)
        assert lib.afp_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_INVALID
    assert lib.afp_process(None, None, 0, None, 0, 0, None) == _lib.ERR_INVALID


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may import it."""
    pkg = os.path.join(ROOT, "audio_filtering_pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def _build_c_host(tmp_path):
    """examples/c_host.c: a host in plain C99 against include/afp.h (what a cgo / JNI / FFI binding calls)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    _build.build()
    libdir = os.path.join(ROOT, "audio_filtering_pipeline_b200")
    exe = str(tmp_path / "c_host")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "examples", "c_host.c"), "-L", libdir, "-lafp_b200", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_plain_c_host_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    import subprocess
    exe = _build_c_host(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu-marked run")
    except ImportError:
        pass
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode != 0 and "afp_create" in res.stderr and "failed" in res.stderr


@pytest.mark.gpu
def test_plain_c_host_runs_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_c_host(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "-> ok" in res.stdout


def test_agc_struct_layout_and_signature():
    """afp_agc_params: five doubles, window_size, reserved (include/afp.h); compat.apply_agc keeps the
    reference's positional arguments and defaults (stream_process_AGC.py:43)."""
    import inspect
    assert C.sizeof(_lib.AgcParams) == 48 and _lib.AgcParams.window_size.offset == 40
    p = _lib.AgcParams.make()
    assert (p.target_level, p.window_size, p.max_gain, p.attack, p.release, p.clip) == (0.1, 512, 10.0, 0.01, 0.1, 0.99)
    src = open(os.path.join(ROOT, "audio_filtering_pipeline_b200", "compat.py")).read()
    assert ("def apply_agc(signal, target_level=0.1, window_size=512, alpha=0.01, max_gain=10.0, attack=0.01, "
            "release=0.1):") in src


def test_new_entry_points_reject_null_without_a_gpu(lib):
    assert lib.afp_set_agc(None, None) == _lib.ERR_INVALID
    assert lib.afp_swap_taps(None, None, 0) == _lib.ERR_INVALID
    assert lib.afp_agc_block(None, None, None, 0, None, 0, 0, 0, None) == _lib.ERR_INVALID
    assert lib.afp_process_host_interleaved(None, None, None, 0, 1) == _lib.ERR_INVALID
