/* afp.h - C-ABI of the B200 streaming block-FIR engine.
 *
 * This is the boundary a maintainer of MrEstefano/Audio-Filtering-Pipeline binds (ctypes, see
 * INTEGRATION.md) to move the per-block DSP of the sounddevice callback onto a B200.  Plain C:
 * pointers, sizes and integer status codes only; no C++/torch types.  Every entry point returns
 * AFP_OK (0) or a non-zero AFP_ERR_* code and never throws; afp_last_error() gives the message.
 *
 * Reference interfaces replaced (file:line under the reference checkout):
 *   afp_set_taps / afp_process        sliding buffer + oaconvolve(buffer, fir_coeff, 'valid')
 *                                     stream_process.py:45-46,96-107; OverlapAddFilter.process_block
 *                                     stream_process_GUI_Presets.py:69-123
 *   afp_set_upsampler                 safe_upsample -> soxr.resample     stream_process.py:66-77
 *   afp_set_eq / afp_set_gains        EQ bank: per-band 'same' sum       stream_process_EQ.py:124-129
 *                                     fused taps-domain sum              stream_process_EQ_GUI.py:86-90
 *                                     filter swap under filter_lock      stream_process_EQ_GUI.py:380-385
 *   out_format / dither in afp_config apply_dither                       stream_process.py:61-64
 *   chain = AFP_CHAIN_REFERENCE       processed[::UPSAMPLE_FACTOR]       stream_process.py:106
 *   afp_process_host                  the body of audio_callback(indata, outdata, frames, ...)
 *                                     stream_process.py:79-120 (host buffers in, host buffers out)
 *   afp_set_agc / afp_agc_block       apply_agc + np.clip(+-0.99) ahead of the up-sampler
 *                                     stream_process_AGC.py:43-89,131-143
 *   afp_spectrum_db                   waterfall row: hanning, rfft, dB   stream_process_GUI_Waterfall.py:471-476
 *
 * Data layout: planar, stream-major.  in[s*in_pitch + i] is frame i of stream s (float32 at the
 * base rate); out[s*out_pitch + j] is output sample j of stream s in `out_format`.  A "stream" is
 * one channel of one session; stereo is two streams.  Streams share taps and nothing else.
 */
#ifndef AFP_H
#define AFP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFP_OK 0
#define AFP_ERR_INVALID 1   /* bad argument / configuration */
#define AFP_ERR_CUDA 2      /* a CUDA runtime call failed; message has the CUDA error string */
#define AFP_ERR_STATE 3     /* call made in the wrong state (e.g. process before set_taps) */
#define AFP_ERR_NOMEM 4
#define AFP_ERR_RANGE 5     /* an input sample is outside the range of the split-FP16 tensor path (|x| > 65504) */

/* afp_config.chain */
#define AFP_CHAIN_DAC 0        /* x -> [EQ] -> polyphase FIR(h, U) -> out at fs*U (the FIR is the interpolator) */
#define AFP_CHAIN_REFERENCE 1  /* x -> interp(h_up, U) -> [EQ] -> FIR(h) -> [::U] -> out at fs (stream_process.py) */

/* afp_config.out_format */
#define AFP_FMT_F32 0
#define AFP_FMT_I16 1
#define AFP_FMT_I24_IN_I32 2   /* 24-bit code sign-extended in an int32 */
#define AFP_FMT_I24_PACKED 3   /* little-endian 3-byte samples */

/* afp_config.dither */
#define AFP_DITHER_NONE 0
#define AFP_DITHER_TPDF 1      /* triangular, (-1,1) LSB, Philox4x32-10 keyed (seed, stream, sample index) */
#define AFP_DITHER_RPDF 2      /* rectangular +-1/2 LSB: the amplitude of the reference's apply_dither */

/* afp_config.eq_mode */
#define AFP_EQ_NONE 0
#define AFP_EQ_FUSED 1         /* one FIR with taps sum_b g_b*h_b, state carried (stream_process_EQ_GUI.py:86-90) */
#define AFP_EQ_PER_BAND 2      /* every band convolved separately and summed with its gain */
#define AFP_EQ_BUFFER_SAME 3   /* parity mode: 'same' convolution over the [history|block] sliding buffer,
                                  zero-extended at both ends, recomputed every block, then the 'valid'
                                  main FIR (stream_process_EQ.py:119-134); taps-domain sum of the bands */
#define AFP_EQ_BUFFER_SAME_PER_BAND 4  /* the same, each band accumulated separately (:127-129) */
#define AFP_EQ_STATELESS_SAME 5 /* parity mode: no carried state; 'same' EQ then 'same' main filter over
                                  the block alone (worker thread, stream_process_EQ_GUI.py:86-95) */

typedef struct afp_ctx afp_ctx;

typedef struct afp_config {
    int32_t device;           /* CUDA device ordinal */
    int32_t n_streams;        /* S */
    int32_t max_frames;       /* largest `frames` a process call will pass (base rate, per stream) */
    int32_t upsample;         /* U in 1..4 */
    int32_t chain;            /* AFP_CHAIN_* */
    int32_t out_format;       /* AFP_FMT_* */
    int32_t dither;           /* AFP_DITHER_* */
    int32_t eq_mode;          /* AFP_EQ_* */
    uint64_t seed;            /* dither seed */
    int32_t first_stream_id;  /* global id of stream 0 (keeps dither identical under multi-GPU sharding) */
    int32_t decimate;         /* keep every decimate-th output sample; 0 = default (U for AFP_CHAIN_REFERENCE,
                                 1 for AFP_CHAIN_DAC); frames*rate must be a multiple of it */
    int32_t dither_bits;      /* bit depth of the float-output dither (apply_dither's bit_depth); 0 = 24 */
    int32_t same_buffer_len;  /* AFP_EQ_BUFFER_SAME*: total length of the sliding buffer in processing-rate
                                 samples.  0 = N_main-1 + block (stream_process_EQ.py:44-45); the Waterfall GUI
                                 uses a fixed 4*blocksize+512 and reads the OLDEST valid outputs
                                 (stream_process_GUI_Waterfall.py:132,450-453) - a constant extra latency */
} afp_config;

typedef struct afp_stage_info {
    int32_t kind;             /* 0 interp, 1 eq, 2 main */
    int32_t upsample;         /* phases computed by this stage */
    int32_t n_taps;           /* taps as given */
    int32_t taps_per_phase_padded;
    int32_t tile_taps;        /* T: taps per register-blocked chunk = base positions per thread */
    int32_t threads;          /* CTA size chosen for the last launch */
    int32_t threads_per_segment;
    int32_t smem_bytes;
    int32_t grid;
    int32_t history;          /* carried samples per stream */
    int32_t kernel;           /* kernel of the last launch: 0 = packed-FFMA2 row-pair form, 1 = tcgen05 / TMEM
                                 Toeplitz 3xTF32 (long single-table U = 1 filters, 256-frame blocks, AFP_NO_TC2),
                                 2 = packed-FFMA2 single-row form (single-table stages, aligned buffers),
                                 3 = tcgen05 / TMEM split-FP16 form with resident Hankel tap cores (long single-table
                                 U = 1 filters, blocks of a multiple of 512 frames) */
} afp_stage_info;

/* Per-stream-group filters: sessions with their own EQ gains (or their own main filter) inside ONE batch - what every
 * reference process does when it reads its own sliders per block (stream_process_EQ_GUI.py:86-90,
 * stream_process_GUI_Presets.py:603-607) and swaps its filter under filter_lock (:380-385).
 *   afp_assign_groups     group_of_stream[s] in 0..G-1 for every stream (NULL: back to one filter set).  Two streams of
 *                         one session (stereo) normally share a group.
 *   afp_set_gains_group   EQ gains of one group (NULL: the context's gains, afp_set_eq / afp_set_gains)
 *   afp_set_taps_group    main filter of one group, same length as the context's (NULL: the context's taps)
 * The chain must be one filter per stream: eq_mode AFP_EQ_NONE or AFP_EQ_FUSED with U = 1 or the collapsed reference
 * chain (for stream s it is  (h_up * sum_b g_{group(s),b} h_b * h_{group(s)})[::U]).  Tables are rebuilt at the next
 * process call; the carried samples are kept (they are input samples, valid for any taps), so a change of gains
 * between two blocks is the reference's live filter swap.  Per-group tables live in device memory and are read per
 * work item by the packed-FFMA2 row-pair kernel; long U = 1 filters in afp_process_blocks calls of a multiple of
 * 16 x 512 frames run on the tensor cores with per-stream Hankel tap cores. */
int afp_assign_groups(afp_ctx* ctx, const int32_t* group_of_stream, int32_t n_streams);
int afp_set_gains_group(afp_ctx* ctx, int32_t group, const double* gains, int32_t n_bands);
int afp_set_taps_group(afp_ctx* ctx, int32_t group, const double* taps, int32_t n_taps);

/* Automatic gain control of the input block (stream_process_AGC.py:43-89), per block and stateless like
 * the reference.  Defaults of apply_agc / of its call site in comments. */
typedef struct afp_agc_params {
    double target_level;      /* 0.1   target RMS */
    double max_gain;          /* 10.0 */
    double attack;            /* 0.01  fraction of window_size: int(attack*window_size) samples, evaluated in
                                       double exactly as Python does */
    double release;           /* 0.1   likewise.  A constant of 0 samples makes the reference's loop divide by
                                       zero when that branch is first taken; it then returns the block
                                       unchanged (:87-89), and so does the affected stream here */
    double clip;              /* 0.99  np.clip of the result at the call site (:143); <= 0 = none */
    int32_t window_size;      /* 512   moving-RMS window in samples; the call site passes min(512, blocksize/2) */
    int32_t reserved;         /* 0 */
} afp_agc_params;

const char* afp_version(void);

int afp_create(const afp_config* cfg, afp_ctx** out_ctx);
int afp_destroy(afp_ctx* ctx);
const char* afp_last_error(const afp_ctx* ctx);   /* ctx may be NULL: error of the last failed afp_create */

/* Filter tables (host pointers, float64 as create_fir_filter returns them).  Each call resets the
 * carried state of the stage it changes, like OverlapAddFilter.set_coefficients
 * (stream_process_GUI_Presets.py:44-67); afp_set_gains does not. */
int afp_set_taps(afp_ctx* ctx, const double* taps, int32_t n_taps);
/* Replace the main filter between two blocks WITHOUT touching the carried samples - the reference's swap
 * of active_filter_params under filter_lock (stream_process_EQ_GUI.py:380-385; deferred form
 * stream_process_AGC.py:27-29,102-108), which leaves input_buffer alone: the next block is the new filter
 * over the true past input, no restart transient.  A longer filter sees zeros beyond what was carried.  This holds
 * for the stage the call rebuilds, for the sliding buffer of the AFP_EQ_BUFFER_SAME* modes (migrated right-aligned
 * when the filter length changes) and when the chain enters its collapsed one-filter form.  When a swap makes the
 * collapsed form impossible (the combined filter no longer fits one launch), the staged chain restarts from the
 * carried INPUT samples: its first stage continues, the later stages start from zero state. */
int afp_swap_taps(afp_ctx* ctx, const double* taps, int32_t n_taps);
int afp_set_upsampler(afp_ctx* ctx, const double* taps, int32_t n_taps);
int afp_set_eq(afp_ctx* ctx, const double* band_taps, const int32_t* band_len, int32_t n_bands,
               const double* gains);
int afp_set_gains(afp_ctx* ctx, const double* gains, int32_t n_bands);

/* One block for all streams.  Device pointers are caller-owned; the call is asynchronous on
 * `cuda_stream` (a cudaStream_t, NULL = the legacy default stream).  Pitches are in elements
 * (bytes for AFP_FMT_I24_PACKED output).  Output samples per stream: afp_out_frames(). */
int afp_process(afp_ctx* ctx, const float* in_dev, int64_t in_pitch, void* out_dev, int64_t out_pitch,
                int32_t frames, void* cuda_stream);
/* n_blocks consecutive blocks (the buffers of n_blocks consecutive callbacks) in one call: block b of stream s
 * starts at in_dev[b*in_block_stride + s*in_pitch] / out_dev[b*out_block_stride + s*out_pitch] (strides in
 * elements, bytes for AFP_FMT_I24_PACKED).  The results equal those of n_blocks afp_process calls - to the bit when
 * both forms run on the same kernel (always from 112 taps up; a shorter filter takes the tensor kernel in calls of
 * >= 8 x 512 frames per stream, from 60 taps also in calls of >= 2 x 512 frames, and the packed-FFMA2 kernel otherwise:
 * the two agree to ~1e-7, not to the bit).  For a
 * chain that is a single filter - U = 1 from 48 taps (the collapsed reference chain included), or a polyphase
 * interpolator with U = 2 / 3 / 4 (from 72 / 84 / 96 taps) - with blocks of a multiple of 512 frames this is ONE launch: the
 * CTA that owns 16 streams walks their blocks with [history | block] resident in shared memory, the carried history
 * (the reference's sliding input_buffer, stream_process.py:45-46,97-98) is read once and written once.  Every other
 * chain runs block by block. */
int afp_process_blocks(afp_ctx* ctx, const float* in_dev, int64_t in_pitch, int64_t in_block_stride, void* out_dev,
                       int64_t out_pitch, int64_t out_block_stride, int32_t frames, int32_t n_blocks, void* cuda_stream);
/* The tensor path for long filters splits every sample into two FP16 numbers and therefore needs |x| <= 65504
 * (float audio is +-1).  A larger sample makes the affected outputs invalid (inf / NaN) and sets a flag that the
 * host-buffer entry points report as AFP_ERR_RANGE; after asynchronous afp_process / afp_process_blocks calls
 * this function waits for the device, returns the flag in *exceeded and clears it. */
int afp_range_status(afp_ctx* ctx, int32_t* exceeded);
/* Same with host buffers (any host memory; pinned memory from afp_host_alloc is fastest).
 * Copies in, runs, copies out and waits.  Dense layout: in[S][frames], out[S][afp_out_frames]. */
int afp_process_host(afp_ctx* ctx, const float* in_host, void* out_host, int32_t frames);
/* n_blocks consecutive blocks, block-major: in[n_blocks][S][frames] -> out[n_blocks][S][afp_out_frames]
 * (the buffers of n_blocks consecutive callbacks back to back).  Host->device copies, kernels and
 * device->host copies of successive blocks / stream chunks overlap on three CUDA streams. */
int afp_process_host_blocks(afp_ctx* ctx, const float* in_host, void* out_host, int32_t frames,
                            int32_t n_blocks);

/* AGC as the first stage of every following afp_process* call (ahead of the interpolator, as in
 * stream_process_AGC.py:131-146); params == NULL switches it off.  Carries no state. */
int afp_set_agc(afp_ctx* ctx, const afp_agc_params* params);
/* The AGC alone: out[s][i] = clip(in[s][i] * gain[s][i]) for n_streams rows of `frames` samples (device
 * pointers, pitches in elements, asynchronous on cuda_stream; in and out must not overlap). */
int afp_agc_block(afp_ctx* ctx, const afp_agc_params* params, const float* in_dev, int64_t in_pitch,
                  float* out_dev, int64_t out_pitch, int32_t n_streams, int32_t frames, void* cuda_stream);

/* One block in the wire format of a sounddevice / PortAudio stream (stream_process.py:128-136): frame-major
 * host buffers in[session][frames][channels] float32 -> out[session][afp_out_frames][channels] in
 * `out_format` (float32, or int16 / int24 PCM).  Stream s = session*channels + channel; n_streams must be a
 * multiple of `channels`.  The (de)interleave runs on the device; synchronous like afp_process_host. */
int afp_process_host_interleaved(afp_ctx* ctx, const float* in_host, void* out_host, int32_t frames, int32_t channels);

/* The waterfall row the reference computes from every output block while its display is on
 * (stream_process_GUI_Waterfall.py:471-476): out[s][k] = 20*log10(|rfft(in[s][0:frames] * hanning(frames))[k]|
 * + 1e-12), k = 0..frames/2, for n_streams rows (device pointers, pitches in elements, asynchronous on
 * cuda_stream).  frames: a power of two, 64..4096 (the GUI's block sizes); anything else is AFP_ERR_INVALID. */
int afp_spectrum_db(afp_ctx* ctx, const float* in_dev, int64_t in_pitch, float* out_dev, int64_t out_pitch,
                    int32_t n_streams, int32_t frames, void* cuda_stream);

int32_t afp_out_frames(const afp_ctx* ctx, int32_t frames);
int32_t afp_out_elem_bytes(const afp_ctx* ctx);

/* State: zero it, or save / restore it (carried history of every stage + dither sample counter). */
int afp_reset(afp_ctx* ctx);
int afp_state_size(const afp_ctx* ctx, size_t* bytes);
int afp_get_state(afp_ctx* ctx, void* host_buf, size_t bytes);
int afp_set_state(afp_ctx* ctx, const void* host_buf, size_t bytes);

/* Introspection for the benchmark harness and DESIGN.md tables. */
int afp_get_stage_info(const afp_ctx* ctx, int32_t stage, afp_stage_info* info);   /* AFP_ERR_INVALID past the end */
int64_t afp_kernel_launches(const afp_ctx* ctx);

/* Pinned host memory for afp_process_host.  afp_host_alloc_wc: write-combined pinned memory for buffers the host only
 * WRITES (input staging: no cache snooping on the device's reads; host reads of it are very slow). */
int afp_host_alloc(void** ptr, size_t bytes);
int afp_host_alloc_wc(void** ptr, size_t bytes);
int afp_host_free(void* ptr);

/* FP32 FMA roof of the device, measured with a packed-FFMA2 issue loop (there is no
 * FP32 entry in MEASURED_PEAKS.json).  Returns TFLOP/s and the SM clock seen during the loop. */
int afp_measure_fma_peak(int32_t device, double* tflops, double* sm_mhz);

/* Tensor-pipe roof of the device for the split-FP16 FIR kernel's roofline: a bare tcgen05.mma loop (kind::f16,
 * M = 128, K = 16, N = n_cols in 16..256, operands in shared memory in the kernel's own un-swizzled layout).
 * n_cols = 256: the pipe's rate; n_cols = 64: the kernel's instruction shape, bound by operand fetch.  Returns
 * TFLOP/s, the SM clock during the loop and clocks per MMA. */
int afp_measure_tensor_peak(int32_t device, int32_t n_cols, double* tflops, double* sm_mhz, double* clk_per_mma);

#ifdef __cplusplus
}
#endif
#endif /* AFP_H */
