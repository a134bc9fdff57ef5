"""BatchFirEngine - the host-side operator over the C-ABI (one engine = one GPU = one chain).

It batches S independent streams (channel x session) through the reference's per-block chain

    [interpolate xU] -> [EQ bank] -> FIR with carried state -> [decimate] -> [dither / quantise]

(stream_process.py:79-120, stream_process_EQ.py:99-166) with all arithmetic in the sm_100a kernels
of csrc/.  Taps come from ``design.create_fir_filter`` exactly as in the reference
(stream_process.py:33-40: designed at ``samplerate * upsample``).  torch is used only to hold
device buffers and to pick the CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_TORCH_DTYPE = {0: "float32", 1: "int16", 2: "int32", 3: "uint8"}
_NP_DTYPE = {0: np.float32, 1: np.int16, 2: np.int32, 3: np.uint8}


def _dbl(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class BatchFirEngine:
    """Streaming block-FIR over ``n_streams`` independent streams on one B200.

    chain='dac'        out[S, frames*U]: the designed filter *is* the interpolation filter
                       (polyphase, only non-zero products are computed).
    chain='reference'  out[S, frames]: interpolate with ``upsampler_taps`` -> [EQ] -> FIR -> keep
                       every U-th sample, the structure of stream_process.py:89-107.
    out_format         'f32' | 'i16' | 'i24' (int32 container) | 'i24_packed' (3-byte little endian)
    dither             'none' | 'tpdf' | 'rpdf'  (Philox4x32-10 keyed by seed/stream/sample index)
    eq_mode            'none' | 'fused' (taps-domain sum, stream_process_EQ_GUI.py:86-90) | 'per_band'
                       (every band convolved and accumulated, stream_process_EQ.py:127-129), both
                       with carried state; parity modes 'buffer_same' / 'buffer_same_per_band'
                       (the sliding-buffer 'same' + 'valid' arithmetic of stream_process_EQ.py:119-134)
                       and 'stateless_same' (stream_process_EQ_GUI.py:86-95, nothing carried).
    decimate           keep every n-th output sample (default: U for chain='reference', else 1)
    dither_bits        bit depth of the float-output dither (apply_dither's ``bit_depth``; default 24)
    same_buffer_len    'buffer_same*' only: fixed sliding-buffer length; 4*blocksize+512 gives the
                       Waterfall GUI's callback (stream_process_GUI_Waterfall.py:132,431-458)
    agc                None | dict(target_level=0.1, window_size=512, max_gain=10.0, attack=0.01,
                       release=0.1, clip=0.99): automatic gain control of every input block ahead of the
                       chain, per block and stateless (stream_process_AGC.py:43-89,131-143)
    """

    def __init__(self, taps, n_streams=1, max_frames=1024, upsample=1, chain="dac", out_format="f32",
                 dither="none", seed=0, device=0, upsampler_taps=None, eq_bands=None, eq_gains=None,
                 eq_mode="none", first_stream_id=0, decimate=0, dither_bits=0, same_buffer_len=0, agc=None):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        if eq_bands is not None and eq_mode in ("none", None):
            eq_mode = "fused"
        cfg = _lib.Config(device=int(device), n_streams=int(n_streams), max_frames=int(max_frames),
                          upsample=int(upsample), chain=_lib.CHAIN[chain],
                          out_format=_lib.FMT[out_format], dither=_lib.DITHER[dither],
                          eq_mode=_lib.EQ_MODE[eq_mode], seed=int(seed) & (2 ** 64 - 1),
                          first_stream_id=int(first_stream_id), decimate=int(decimate),
                          dither_bits=int(dither_bits), same_buffer_len=int(same_buffer_len))
        _lib.check(self._lib.afp_create(C.byref(cfg), C.byref(self._ctx)))
        self.cfg = cfg
        self.n_streams, self.max_frames, self.upsample = int(n_streams), int(max_frames), int(upsample)
        self.device = int(device)
        self.fmt = cfg.out_format
        self.taps = None
        try:
            self.set_taps(taps)
            if upsampler_taps is not None:
                self.set_upsampler(upsampler_taps)
            if eq_bands is not None:
                self.set_eq(eq_bands, eq_gains)
            if agc is not None:
                self.set_agc(**agc)
        except Exception:
            self.close()
            raise

    # -- filter tables -----------------------------------------------------------------------
    def set_taps(self, taps):
        a, p = _dbl(taps)
        _lib.check(self._lib.afp_set_taps(self._ctx, p, len(a)), self._ctx)
        self.taps = a

    def swap_taps(self, taps):
        """New main filter from the next block on, carried samples kept (the reference's live filter
        swap, stream_process_EQ_GUI.py:380-385): no restart transient."""
        a, p = _dbl(taps)
        _lib.check(self._lib.afp_swap_taps(self._ctx, p, len(a)), self._ctx)
        self.taps = a

    def set_upsampler(self, taps):
        a, p = _dbl(taps)
        _lib.check(self._lib.afp_set_upsampler(self._ctx, p, len(a)), self._ctx)

    def set_eq(self, bands, gains=None):
        bands = [np.asarray(b, dtype=np.float64).ravel() for b in bands]
        lens = (C.c_int32 * len(bands))(*[len(b) for b in bands])
        flat, fp = _dbl(np.concatenate(bands) if bands else np.zeros(0))
        gp = None
        if gains is not None:
            g, gp = _dbl(gains)
        _lib.check(self._lib.afp_set_eq(self._ctx, fp, lens, len(bands), gp), self._ctx)

    def set_gains(self, gains):
        g, gp = _dbl(gains)
        _lib.check(self._lib.afp_set_gains(self._ctx, gp, len(g)), self._ctx)

    # -- per-stream-group filters ---------------------------------------------------------------
    def assign_groups(self, group_of_stream):
        """Give every stream a filter group (``None``: back to one filter set): sessions with their own EQ gains or
        main taps inside one batch, as every reference process reads its own sliders per block
        (stream_process_EQ_GUI.py:86-90).  The chain must be one filter per stream (eq_mode 'none' or 'fused')."""
        if group_of_stream is None:
            _lib.check(self._lib.afp_assign_groups(self._ctx, None, 0), self._ctx)
            return
        g = np.ascontiguousarray(group_of_stream, dtype=np.int32).ravel()
        _lib.check(self._lib.afp_assign_groups(self._ctx, g.ctypes.data_as(C.POINTER(C.c_int32)), len(g)), self._ctx)

    def set_group_gains(self, group, gains):
        """EQ gains of one group (``None``: the engine-level gains); takes effect at the next block, state kept."""
        if gains is None:
            _lib.check(self._lib.afp_set_gains_group(self._ctx, int(group), None, 0), self._ctx)
            return
        a, p = _dbl(gains)
        _lib.check(self._lib.afp_set_gains_group(self._ctx, int(group), p, len(a)), self._ctx)

    def set_group_taps(self, group, taps):
        """Main filter of one group, same length as the engine's (``None``: the engine-level taps)."""
        if taps is None:
            _lib.check(self._lib.afp_set_taps_group(self._ctx, int(group), None, 0), self._ctx)
            return
        a, p = _dbl(taps)
        _lib.check(self._lib.afp_set_taps_group(self._ctx, int(group), p, len(a)), self._ctx)

    def set_agc(self, enabled=True, **params):
        """Switch the input AGC on (keyword arguments of ``apply_agc`` plus ``clip``) or off
        (``set_agc(False)``).  Takes effect at the next block."""
        if not enabled:
            _lib.check(self._lib.afp_set_agc(self._ctx, None), self._ctx)
            return
        p = _lib.AgcParams.make(**params)
        _lib.check(self._lib.afp_set_agc(self._ctx, C.byref(p)), self._ctx)

    def agc_block(self, x, out=None, **params):
        """The AGC kernel alone on a float32 CUDA tensor [rows, frames] (any row count)."""
        import torch
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
            raise ValueError("x must be a float32 CUDA tensor of shape [rows, frames]")
        if x.stride(1) != 1:
            x = x.contiguous()
        if out is None:
            out = torch.empty_like(x, memory_format=torch.contiguous_format)
        p = _lib.AgcParams.make(**params)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rows, frames = x.shape
        # (the stride of a one-row tensor is arbitrary)
        xp, op = (x.stride(0), out.stride(0)) if rows > 1 else (max(x.stride(0), frames), max(out.stride(0), frames))
        _lib.check(self._lib.afp_agc_block(self._ctx, C.byref(p), x.data_ptr(), xp, out.data_ptr(), op,
                                           rows, frames, stream), self._ctx)
        return out

    def spectrum_db(self, x, out=None):
        """Waterfall rows (stream_process_GUI_Waterfall.py:471-476) of a float32 CUDA tensor [rows, frames]:
        ``20*log10(|rfft(x*hanning)| + 1e-12)`` -> [rows, frames//2 + 1].  frames: power of two, 64..4096."""
        import torch
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
            raise ValueError("x must be a float32 CUDA tensor of shape [rows, frames]")
        if x.stride(1) != 1:
            x = x.contiguous()
        rows, frames = x.shape
        if out is None:
            out = torch.empty((rows, frames // 2 + 1), dtype=torch.float32, device=x.device)
        xp, op = (x.stride(0), out.stride(0)) if rows > 1 else (max(x.stride(0), frames), max(out.stride(0), frames // 2 + 1))
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._lib.afp_spectrum_db(self._ctx, x.data_ptr(), xp, out.data_ptr(), op, rows, frames, stream),
                   self._ctx)
        return out

    # -- processing --------------------------------------------------------------------------
    def out_frames(self, frames):
        return int(self._lib.afp_out_frames(self._ctx, int(frames)))

    def _out_cols(self, frames):
        n = self.out_frames(frames)
        return n * 3 if self.fmt == 3 else n

    def _check_out(self, out, shape):
        """A caller-supplied host output buffer goes to C as a bare pointer: it must be exactly the array the library writes."""
        if not isinstance(out, np.ndarray) or out.dtype != _NP_DTYPE[self.fmt] or not out.flags.c_contiguous or \
                not out.flags.writeable or out.size != int(np.prod(shape)):
            raise ValueError(f"out must be a writable C-contiguous {np.dtype(_NP_DTYPE[self.fmt]).name} array of {int(np.prod(shape))} "
                             f"elements (shape {tuple(shape)})")

    def process(self, x, out=None):
        """One block for all streams on device tensors: ``x`` float32 cuda [S, frames] (row stride
        arbitrary, unit column stride).  Asynchronous on torch's current stream."""
        import torch
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.shape[0] != self.n_streams:
            raise ValueError("x must be a float32 CUDA tensor of shape [n_streams, frames]")
        if x.stride(1) != 1:
            x = x.contiguous()
        frames = x.shape[1]
        if out is None:
            out = torch.empty((self.n_streams, self._out_cols(frames)),
                              dtype=getattr(torch, _TORCH_DTYPE[self.fmt]), device=x.device)
        elif out.stride(1) != 1 or out.shape[0] != self.n_streams or out.shape[1] < self._out_cols(frames):
            raise ValueError("out has the wrong shape / stride")
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._lib.afp_process(self._ctx, x.data_ptr(), x.stride(0), out.data_ptr(),
                                         out.stride(0), frames, stream), self._ctx)
        return out

    def process_blocks(self, x, out=None):
        """``n`` consecutive blocks on device tensors: ``x`` float32 cuda [n, S, frames] (unit frame stride) ->
        [n, S, out_cols].  The result of ``n`` calls of :meth:`process` (to the bit when both run on the same kernel,
        see include/afp.h); a chain that is one filter - U = 1, the collapsed reference chain, or a 2x / 3x / 4x polyphase
        interpolator - runs as a single tensor-core launch that keeps [history | block] in shared memory from block
        to block (afp_process_blocks).  Asynchronous on torch's current stream."""
        import torch
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3 or x.shape[1] != self.n_streams:
            raise ValueError("x must be a float32 CUDA tensor of shape [n_blocks, n_streams, frames]")
        if x.stride(2) != 1:
            x = x.contiguous()
        nb, _, frames = x.shape
        cols = self._out_cols(frames)
        if out is None:
            out = torch.empty((nb, self.n_streams, cols), dtype=getattr(torch, _TORCH_DTYPE[self.fmt]), device=x.device)
        elif (out.dim() != 3 or out.stride(2) != 1 or out.shape[0] != nb or out.shape[1] != self.n_streams
              or out.shape[2] < cols or out.dtype != getattr(torch, _TORCH_DTYPE[self.fmt])):
            raise ValueError("out has the wrong shape / stride / dtype")
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._lib.afp_process_blocks(self._ctx, x.data_ptr(), x.stride(1), x.stride(0), out.data_ptr(),
                                                out.stride(1), out.stride(0), frames, nb, stream), self._ctx)
        return out

    def range_exceeded(self):
        """True if the split-FP16 tensor path has seen a sample beyond +-65504 since the last check (waits for the
        device; clears the flag).  The host-buffer calls raise ``AfpError`` (code 5) for this on their own."""
        flag = C.c_int32(0)
        _lib.check(self._lib.afp_range_status(self._ctx, C.byref(flag)), self._ctx)
        return bool(flag.value)

    def process_host(self, x, out=None):
        """One block with host arrays (numpy float32 [S, frames]); returns a numpy array."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[0] != self.n_streams:
            raise ValueError("x must have shape [n_streams, frames]")
        frames = x.shape[1]
        if out is None:
            out = np.empty((self.n_streams, self._out_cols(frames)), dtype=_NP_DTYPE[self.fmt])
        else:
            self._check_out(out, (self.n_streams, self._out_cols(frames)))
        _lib.check(self._lib.afp_process_host(self._ctx, x.ctypes.data, out.ctypes.data, frames), self._ctx)
        return out

    def process_host_interleaved(self, x, out=None):
        """One block in the sounddevice wire format: x float32 [frames, channels] (one session) or
        [sessions, frames, channels] -> the same shape with ``out_frames`` frames, in ``out_format``
        ('i24_packed': a trailing axis of 3 bytes).  Stream s = session*channels + channel."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        single = x.ndim == 2
        if single:
            x = x[None]
        if x.ndim != 3 or x.shape[0] * x.shape[2] != self.n_streams:
            raise ValueError("x must have shape [sessions, frames, channels] with sessions*channels == n_streams")
        se, frames, ch = x.shape
        shape = (se, self.out_frames(frames), ch) + ((3,) if self.fmt == 3 else ())
        if out is None:
            out = np.empty(shape, dtype=_NP_DTYPE[self.fmt])
        elif out.size != int(np.prod(shape)) or not out.flags.c_contiguous or out.dtype != _NP_DTYPE[self.fmt]:
            raise ValueError("out has the wrong shape / dtype")
        _lib.check(self._lib.afp_process_host_interleaved(self._ctx, x.ctypes.data, out.ctypes.data, frames, ch),
                   self._ctx)
        return out.reshape(shape[1:]) if single and out.ndim == len(shape) else out

    def process_host_blocks(self, x, out=None):
        """``n`` consecutive blocks with host arrays, block-major: x float32 [n, S, frames] ->
        [n, S, out_cols].  Copies and kernels of successive blocks overlap (three CUDA streams);
        pass pinned arrays (``pinned_empty``) for full PCIe speed."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3 or x.shape[1] != self.n_streams:
            raise ValueError("x must have shape [n_blocks, n_streams, frames]")
        nb, _, frames = x.shape
        if out is None:
            out = np.empty((nb, self.n_streams, self._out_cols(frames)), dtype=_NP_DTYPE[self.fmt])
        else:
            self._check_out(out, (nb, self.n_streams, self._out_cols(frames)))
        _lib.check(self._lib.afp_process_host_blocks(self._ctx, x.ctypes.data, out.ctypes.data, frames, nb),
                   self._ctx)
        return out

    # -- state -------------------------------------------------------------------------------
    def reset(self):
        _lib.check(self._lib.afp_reset(self._ctx), self._ctx)

    def get_state(self):
        n = C.c_size_t()
        _lib.check(self._lib.afp_state_size(self._ctx, C.byref(n)), self._ctx)
        buf = np.empty(n.value, dtype=np.uint8)
        _lib.check(self._lib.afp_get_state(self._ctx, buf.ctypes.data, n.value), self._ctx)
        return buf

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        _lib.check(self._lib.afp_set_state(self._ctx, blob.ctypes.data, blob.size), self._ctx)

    # -- introspection -----------------------------------------------------------------------
    def stage_info(self):
        out, i = [], 0
        while True:
            info = _lib.StageInfo()
            if self._lib.afp_get_stage_info(self._ctx, i, C.byref(info)) != 0:
                return out
            out.append({n: getattr(info, n) for n, _ in info._fields_})
            i += 1

    @property
    def kernel_launches(self):
        return int(self._lib.afp_kernel_launches(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.afp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class PinnedArray:
    """Page-locked host memory from afp_host_alloc with a numpy view in ``.a`` (keep the object
    alive while the view is in use; ``free()`` or garbage collection releases it).  ``write_combined``: for
    buffers the host only writes (input staging); reading such memory from the host is very slow."""

    def __init__(self, shape, dtype, write_combined=False):
        self._lib = _lib.load()
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        self._ptr = C.c_void_p()
        alloc = self._lib.afp_host_alloc_wc if write_combined else self._lib.afp_host_alloc
        _lib.check(alloc(C.byref(self._ptr), max(count * dtype.itemsize, 1)))
        buf = (C.c_uint8 * max(count * dtype.itemsize, 1)).from_address(self._ptr.value)
        self.a = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def free(self):
        if self._ptr:
            self.a = None
            self._lib.afp_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def measure_fma_peak(device=0):
    """FP32 FMA roof of the device (TFLOP/s, SM MHz) from the packed-FFMA2 issue loop."""
    lib = _lib.load()
    t, m = C.c_double(), C.c_double()
    _lib.check(lib.afp_measure_fma_peak(int(device), C.byref(t), C.byref(m)))
    return t.value, m.value


def measure_tensor_peak(device=0, n_cols=256):
    """Tensor-pipe roof (TFLOP/s, SM MHz, clocks per MMA) from a bare tcgen05.mma kind::f16 loop with M = 128, K = 16,
    N = ``n_cols``; 64 is the split-FP16 FIR kernel's instruction shape (operand-fetch bound), 256 the pipe's rate."""
    lib = _lib.load()
    t, m, c = C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.afp_measure_tensor_peak(int(device), int(n_cols), C.byref(t), C.byref(m), C.byref(c)))
    return t.value, m.value, c.value
