"""ctypes binding of the C-ABI in include/afp.h (libafp_b200.so, built by _build.py).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libafp_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_RANGE = 0, 1, 2, 3, 4, 5
CHAIN = {"dac": 0, "reference": 1}
FMT = {"f32": 0, "i16": 1, "i24": 2, "i24_in_i32": 2, "i24_packed": 3}
DITHER = {"none": 0, None: 0, False: 0, "tpdf": 1, True: 1, "rpdf": 2}
EQ_MODE = {"none": 0, None: 0, "fused": 1, "per_band": 2, "buffer_same": 3, "buffer_same_per_band": 4,
           "stateless_same": 5}

# every symbol include/afp.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "afp_version", "afp_create", "afp_destroy", "afp_last_error", "afp_set_taps", "afp_set_upsampler",
    "afp_set_eq", "afp_set_gains", "afp_process", "afp_process_host", "afp_process_host_blocks", "afp_out_frames",
    "afp_out_elem_bytes", "afp_reset", "afp_state_size", "afp_get_state", "afp_set_state",
    "afp_get_stage_info", "afp_kernel_launches", "afp_host_alloc", "afp_host_free",
    "afp_measure_fma_peak", "afp_set_agc", "afp_agc_block", "afp_process_host_interleaved", "afp_swap_taps", "afp_spectrum_db",
    "afp_process_blocks", "afp_range_status", "afp_measure_tensor_peak",
    "afp_assign_groups", "afp_set_gains_group", "afp_set_taps_group", "afp_host_alloc_wc",
]


class AfpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"afp error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("n_streams", C.c_int32), ("max_frames", C.c_int32),
        ("upsample", C.c_int32), ("chain", C.c_int32), ("out_format", C.c_int32),
        ("dither", C.c_int32), ("eq_mode", C.c_int32), ("seed", C.c_uint64),
        ("first_stream_id", C.c_int32), ("decimate", C.c_int32), ("dither_bits", C.c_int32),
        ("same_buffer_len", C.c_int32),
    ]


class StageInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "upsample", "n_taps", "taps_per_phase_padded", "tile_taps", "threads",
        "threads_per_segment", "smem_bytes", "grid", "history", "kernel")]


class AgcParams(C.Structure):
    """afp_agc_params; the defaults are apply_agc's (stream_process_AGC.py:43) and its call site's clip (:143)."""
    _fields_ = [("target_level", C.c_double), ("max_gain", C.c_double), ("attack", C.c_double),
                ("release", C.c_double), ("clip", C.c_double), ("window_size", C.c_int32),
                ("reserved", C.c_int32)]

    @classmethod
    def make(cls, target_level=0.1, window_size=512, max_gain=10.0, attack=0.01, release=0.1, clip=0.99):
        return cls(float(target_level), float(max_gain), float(attack), float(release), float(clip),
                   int(window_size), 0)


_lib = None


def load():
    """Load libafp_b200.so and declare the signatures.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m audio_filtering_pipeline_b200._build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the block-FIR path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    dp = C.POINTER(C.c_double)
    lib.afp_version.restype = C.c_char_p
    lib.afp_version.argtypes = []
    lib.afp_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.afp_destroy.argtypes = [vp]
    lib.afp_last_error.restype = C.c_char_p
    lib.afp_last_error.argtypes = [vp]
    lib.afp_set_taps.argtypes = [vp, dp, i32]
    lib.afp_set_upsampler.argtypes = [vp, dp, i32]
    lib.afp_swap_taps.argtypes = [vp, dp, i32]
    lib.afp_set_eq.argtypes = [vp, dp, C.POINTER(i32), i32, dp]
    lib.afp_set_gains.argtypes = [vp, dp, i32]
    lib.afp_assign_groups.argtypes = [vp, C.POINTER(i32), i32]
    lib.afp_set_gains_group.argtypes = [vp, i32, dp, i32]
    lib.afp_set_taps_group.argtypes = [vp, i32, dp, i32]
    lib.afp_process.argtypes = [vp, vp, i64, vp, i64, i32, vp]
    lib.afp_process_blocks.argtypes = [vp, vp, i64, i64, vp, i64, i64, i32, i32, vp]
    lib.afp_range_status.argtypes = [vp, C.POINTER(i32)]
    lib.afp_process_host.argtypes = [vp, vp, vp, i32]
    lib.afp_process_host_blocks.argtypes = [vp, vp, vp, i32, i32]
    lib.afp_process_host_interleaved.argtypes = [vp, vp, vp, i32, i32]
    lib.afp_out_frames.argtypes = [vp, i32]
    lib.afp_out_frames.restype = i32
    lib.afp_out_elem_bytes.argtypes = [vp]
    lib.afp_out_elem_bytes.restype = i32
    lib.afp_reset.argtypes = [vp]
    lib.afp_state_size.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.afp_get_state.argtypes = [vp, vp, C.c_size_t]
    lib.afp_set_state.argtypes = [vp, vp, C.c_size_t]
    lib.afp_get_stage_info.argtypes = [vp, i32, C.POINTER(StageInfo)]
    lib.afp_kernel_launches.argtypes = [vp]
    lib.afp_kernel_launches.restype = i64
    lib.afp_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.afp_host_alloc_wc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.afp_host_free.argtypes = [vp]
    lib.afp_measure_fma_peak.argtypes = [i32, dp, dp]
    lib.afp_measure_tensor_peak.argtypes = [i32, i32, dp, dp, dp]
    lib.afp_spectrum_db.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp]
    lib.afp_set_agc.argtypes = [vp, C.POINTER(AgcParams)]
    lib.afp_agc_block.argtypes = [vp, C.POINTER(AgcParams), vp, i64, vp, i64, i32, i32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:   # default: status code
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != OK:
        msg = load().afp_last_error(ctx)
        raise AfpError(rc, msg.decode() if msg else "unknown error")
