// Batched streaming FIR / polyphase-interpolator kernel for sm_100a (kernels 1-5 of the north star).
//
// What it replaces in the reference (file:line under /root/reference):
//   * sliding float32 buffer + oaconvolve(buffer, h, 'valid')          stream_process.py:45-46,97-102
//   * per-band / fused EQ 'same' convolution over that buffer          stream_process_EQ.py:124-129,
//                                                                      stream_process_EQ_GUI.py:86-90
//   * zero-stuff + anti-imaging filter (soxr.resample per block)       stream_process.py:66-77
//   * processed[::U] decimation                                        stream_process.py:106
//   * apply_dither                                                     stream_process.py:61-64
//
// Formulation.  Per stream the kernel sees a logical base-rate sequence
//       X = [ src0 (len0 samples) | src1 (len1 samples) | zeros ... ],   X[n<0] = 0
// (streaming: src0 = carried history, src1 = the new block) and computes, for base positions
// q in [q_begin, q_begin+n_q) and phases p in [0,U):
//       y[(q-q_begin)*U + p] = sum_j  h[j*U + p] * X[q - j]
// i.e. a direct-form FIR for U = 1 and a polyphase interpolator that touches only the non-zero
// products for U > 1 (taps are passed in their natural order h[k], zero-padded to U*T*n_chunks).
//
// Mapping to the SM (design notes, see DESIGN.md for the measurements behind them):
//   * Packed FP32: the inner product is issued as FFMA2 (fma.rn.f32x2), two FMAs per lane per
//     instruction, which reaches 128 FMA/clk/SM with half the issue slots of scalar FFMA, so the
//     LDS / address / loop instructions ride in the free slots.  The two halves of every pair are
//     two different rows (row A, row B = two segments that share the same taps); the shared-memory
//     tile stores them interleaved as (A[n], B[n]) so one LDS.128 yields two ready operand pairs.
//   * The tap is the scalar-broadcast operand of FFMA2 and lives in a uniform register loaded
//     straight from the kernel-parameter constant bank (LDCU): taps cost no vector register, no
//     shared memory bandwidth and one issue slot per 2-4 taps.
//   * Register blocking: each thread owns T consecutive base positions (x U phases) of both rows and
//     walks the taps in chunks of T with a 2T-deep register window -> one shared-memory word per
//     U*T FMAs.
//   * Shared-memory layout: 16-byte granules (2 pairs) with one pad granule per 8, which makes the
//     per-thread-contiguous LDS.128 pattern (lane stride T/2 granules) bank-conflict free.
//   * Outputs leave from registers as 16-byte stores; dither/quantisation is applied on the way.
//   * The history for the next block (last hist_len samples of [src0|src1]) is written by the
//     loader while the tile is in registers - no second pass, no extra launch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "afp_philox.cuh"

namespace afp {

constexpr int MAX_TAPS = 6144;   // floats in the per-launch tap table (all phases / bands, zero padded)

struct FirParams {
    const float* src0;           // [n_streams][pitch0]  first len0 logical samples
    const float* src1;           // [n_streams][pitch1]  next len1 logical samples
    long long pitch0, pitch1;
    int len0, len1;
    int q_begin, n_q;            // base positions computed per stream
    int tg_log2;                 // log2(threads per segment); a segment is (1<<tg_log2)*T positions
    int segs_per_stream;
    int segs_shift;              // log2(segs_per_stream) if it is a power of two, else -1
    int n_seg_total;             // n_streams * segs_per_stream
    int n_pairs;                 // work items: ceil(n_seg_total / 2)
    int n_chunks;                // padded taps per phase / T
    int n_sub;                   // passes over the tap axis accumulated into one output (EQ bands), >= 1
    int sub_stride;              // floats between the tables of consecutive passes
    void* out;                   // [n_streams][out_pitch] elements of the output format
    long long out_pitch;
    int decim;                   // keep output samples whose index % decim == 0
    int fmt, dither;
    unsigned long long seed;
    unsigned long long sample_base;   // absolute index (per stream) of this call's first kept output
    int stream_id0;              // global id of stream 0 (Philox counter word 2)
    float* hist_out;             // [n_streams][hist_pitch]; hist_out[s][i] = X[hist_from + i]
    long long hist_pitch;
    int hist_len, hist_from;
    int fast0, fast1, fast_hist, fast_out;   // 16-byte alignment guarantees, computed on the host
    int fast_all;                // every 4-sample group of every region is aligned and lies in one source
    float gain;                  // applied to the accumulators before the epilogue (1.0 = none)
    float dither_lsb;            // float-output dither: value of one LSB (2^-(bits-1))
    // per-stream-group filters (GROUPED instantiation): the taps of stream s are gtaps + group_of_stream[s] * gstride,
    // same layout as `taps`; pair_single = 1: one row per work item (the two rows of a pair must share their taps)
    const float* gtaps;
    const int* group_of_stream;
    long long gstride;
    int pair_single;
    alignas(16) float taps[MAX_TAPS];
};

__device__ __forceinline__ unsigned long long as_u64(const float2& v) {
    return *reinterpret_cast<const unsigned long long*>(&v);
}
__device__ __forceinline__ void ffma2(float2& acc, const float2& x, float h) {
    unsigned long long d = as_u64(acc);
    const float2 hh = make_float2(h, h);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(as_u64(x)), "l"(as_u64(hh)));
    acc = *reinterpret_cast<float2*>(&d);
}

// Shared-memory layout: 16-byte granules (2 pairs); a block of T pairs is T/2 granules followed by one
// pad granule.  A thread reads whole blocks and consecutive lanes read consecutive blocks, so the lane
// stride is T/2+1 granules - odd - and any 8 consecutive lanes of an LDS.128 hit 8 different 16-byte bank
// groups whatever the chunk offset (one pad per 8 granules, the first layout, was only conflict free
// for T = 16: ncu showed 44 % conflicting wavefronts for T = 4).
template <int T>
__device__ __forceinline__ int phys_granule(int g) { return g + g / (T / 2); }
__host__ __device__ constexpr int region_granules(int pairs, int T) { return (pairs / T) * (T / 2 + 1) + 2; }

template <int T>
__device__ __forceinline__ void load_block(float2 (&w)[T], const float4* __restrict__ region, int block) {
    const float4* p = region + block * (T / 2 + 1);
#pragma unroll
    for (int m = 0; m < T / 2; ++m) {
        const float4 v = p[m];
        w[2 * m] = make_float2(v.x, v.y);
        w[2 * m + 1] = make_float2(v.z, v.w);
    }
}

// One chunk of T taps: hi holds X[q0 .. q0+T-1] relative to the chunk origin, lo the T samples below.
template <int U, int T>
__device__ __forceinline__ void chunk_fma(float2 (&acc)[U][T], const float2 (&lo)[T], const float2 (&hi)[T],
                                          const float* __restrict__ taps) {
#pragma unroll
    for (int j = 0; j < T; ++j) {
#pragma unroll
        for (int p = 0; p < U; ++p) {
            const float h = taps[j * U + p];
#pragma unroll
            for (int r = 0; r < T; ++r) {
                if (r >= j) ffma2(acc[p][r], hi[r - j], h);
                else        ffma2(acc[p][r], lo[T + r - j], h);
            }
        }
    }
}

struct RowSrc {
    const float* r0;
    const float* r1;
    bool valid;
};

__device__ __forceinline__ float4 fetch4(const FirParams& P, const RowSrc& rs, int n) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!rs.valid) return v;
    const int L = P.len0 + P.len1;
    if (P.fast0 && n >= 0 && n + 4 <= P.len0) {
        v = __ldg(reinterpret_cast<const float4*>(rs.r0 + n));
    } else if (P.fast1 && n >= P.len0 && n + 4 <= L) {
        v = __ldg(reinterpret_cast<const float4*>(rs.r1 + (n - P.len0)));
    } else if (n + 4 > 0 && n < L) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = n + i;
            e[i] = (m < 0 || m >= L) ? 0.f : (m < P.len0 ? __ldg(rs.r0 + m) : __ldg(rs.r1 + (m - P.len0)));
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
    }
    return v;
}

__device__ __forceinline__ void hist_store4(const FirParams& P, float* hrow, int n, const float4& v, int own_from) {
    // logical samples n..n+3; keep those >= own_from that fall in [hist_from, hist_from + hist_len)
    const int lo = max(P.hist_from, own_from), hi = P.hist_from + P.hist_len;
    if (n + 4 <= lo || n >= hi) return;
    if (P.fast_hist && n >= lo && n + 4 <= hi) {
        *reinterpret_cast<float4*>(hrow + (n - P.hist_from)) = v;
    } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (n + i >= lo && n + i < hi) hrow[n + i - P.hist_from] = e[i];
    }
}

// ---- epilogue ---------------------------------------------------------------------------------
template <int NOUT, int MODE, bool SIDE_BY_SIDE = false>
__device__ __forceinline__ void dither_values(float (&d)[NOUT], unsigned long long seed, uint32_t sid, unsigned long long idx0) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (SIDE_BY_SIDE && NOUT % 16 == 0 && (idx0 & 3ull) == 0) {     // four calls at a time, round by round
#pragma unroll
        for (int i0 = 0; i0 < NOUT; i0 += 16) {
            uint32_t c0[4], c1[4], c2[4], c3[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long c = (idx0 + i0 + 4 * j) >> 2;
                c0[j] = (uint32_t)c; c1[j] = (uint32_t)(c >> 32); c2[j] = sid; c3[j] = 0u;
            }
            philox4x32_10_multi<4>(c0, c1, c2, c3, k0, k1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                d[i0 + 4 * j] = dither_lsb<MODE>(c0[j]); d[i0 + 4 * j + 1] = dither_lsb<MODE>(c1[j]);
                d[i0 + 4 * j + 2] = dither_lsb<MODE>(c2[j]); d[i0 + 4 * j + 3] = dither_lsb<MODE>(c3[j]);
            }
        }
        return;
    }
    if (NOUT % 4 == 0 && (idx0 & 3ull) == 0) {     // the normal case: one Philox call per four samples
#pragma unroll
        for (int i = 0; i < NOUT; i += 4) {
            const unsigned long long c = (idx0 + i) >> 2;
            const Philox4 w = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), sid, 0u, k0, k1);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + j < NOUT) d[i + j] = dither_lsb<MODE>(w.w[j]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const unsigned long long idx = idx0 + i, c = idx >> 2;
            const Philox4 w = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), sid, 0u, k0, k1);
            const uint32_t j = (uint32_t)idx & 3u;
            d[i] = dither_lsb<MODE>(j == 0 ? w.w[0] : (j == 1 ? w.w[1] : (j == 2 ? w.w[2] : w.w[3])));
        }
    }
}

template <int NB>
__device__ __forceinline__ void store_words(void* dst, const uint32_t (&w)[NB]) {
    // NB 32-bit words to a destination whose alignment matches the widest vector dividing NB*4
    if (NB % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NB / 4; ++i)
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else if (NB % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NB / 2; ++i) reinterpret_cast<uint2*>(dst)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < NB; ++i) reinterpret_cast<uint32_t*>(dst)[i] = w[i];
    }
}

__device__ __forceinline__ int elem_bytes(int fmt) {
    return fmt == FMT_I16 ? 2 : (fmt == FMT_I24_PACKED ? 3 : 4);
}

// EPI: epilogue specialisation.  EPI_GENERIC reads format / dither / decimation from the parameters;
// EPI = fmt*4 + dither fixes them at compile time (decim == 1), which removes every branch and all
// dead conversion code from the short-filter kernels, where the epilogue is a large share of the work.
constexpr int EPI_GENERIC = -1;
__host__ __device__ constexpr int epi_code(int fmt, int dither) { return fmt * 4 + dither; }

// Where a thread's finished, format-converted words go.  DirectSink: vector stores from registers.
// (afp_fir_row.cuh adds a sink that stages them in shared memory and hands them to the bulk-copy engine.)
struct DirectSink {
    static constexpr bool kPhiloxSideBySide = false;
    static constexpr bool kAllFull = false;   // true: the caller guarantees n_valid == NOUT and aligned output on every lane
    template <int NB>
    __device__ __forceinline__ void put(void* dst, const uint32_t (&w)[NB]) const { store_words<NB>(dst, w); }
};

// y[NOUT]: consecutive output samples o_base .. o_base+NOUT-1 (pre-decimation) of one stream row.
template <int NOUT, int EPI, class PT, class SINK = DirectSink>
__device__ __forceinline__ void store_row(const PT& P_, float (&y)[NOUT], int stream, long long o_base, int n_valid,
                                          SINK sink = SINK()) {
    if (!SINK::kAllFull && n_valid <= 0) return;
    struct View {   // the epilogue fields of FirParams / TcParams, fmt / dither / decim possibly compile-time
        int fmt, dither, decim;
        float gain, dither_lsb;
        void* out; long long out_pitch;
        unsigned long long seed, sample_base;
        int stream_id0, fast_out;
    };
    const View P = {EPI >= 0 ? (EPI >> 2) : P_.fmt, EPI >= 0 ? (EPI & 3) : P_.dither, EPI >= 0 ? 1 : P_.decim,
                    P_.gain, P_.dither_lsb, P_.out, P_.out_pitch, P_.seed, P_.sample_base, P_.stream_id0, P_.fast_out};
    if (EPI < 0 && P.gain != 1.0f) {   // specialised epilogues are only selected with gain == 1
#pragma unroll
        for (int i = 0; i < NOUT; ++i) y[i] = __fmul_rn(y[i], P.gain);
    }
    unsigned char* row = reinterpret_cast<unsigned char*>(P.out) + (long long)stream * P.out_pitch * elem_bytes(P.fmt);
    if (P.decim == 1) {
        float d[NOUT];
        const unsigned long long idx0 = P.sample_base + (unsigned long long)o_base;
        if (P.dither == DITHER_TPDF) dither_values<NOUT, DITHER_TPDF, SINK::kPhiloxSideBySide>(d, P.seed, (uint32_t)(P.stream_id0 + stream), idx0);
        else if (P.dither == DITHER_RPDF) dither_values<NOUT, DITHER_RPDF, SINK::kPhiloxSideBySide>(d, P.seed, (uint32_t)(P.stream_id0 + stream), idx0);
        else {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) d[i] = 0.f;
        }
        const bool full = SINK::kAllFull || (P.fast_out && n_valid == NOUT);
        if (P.fmt == FMT_F32) {
            if (P.dither != DITHER_NONE) {
#pragma unroll
                for (int i = 0; i < NOUT; ++i) y[i] = dither_float(y[i], d[i], P.dither_lsb);
            }
            float* o = reinterpret_cast<float*>(row) + o_base;
            if (full) {
                uint32_t w[NOUT];
#pragma unroll
                for (int i = 0; i < NOUT; ++i) w[i] = __float_as_uint(y[i]);
                sink.template put<NOUT>(o, w);
            } else {
#pragma unroll
                for (int i = 0; i < NOUT; ++i) if (i < n_valid) o[i] = y[i];
            }
        } else if (P.fmt == FMT_I16) {
            int q[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; ++i) q[i] = quantise<16>(y[i], d[i]);
            short* o = reinterpret_cast<short*>(row) + o_base;
            if (full) {
                uint32_t w[NOUT / 2];
#pragma unroll
                for (int i = 0; i < NOUT / 2; ++i) w[i] = ((uint32_t)q[2 * i] & 0xFFFFu) | ((uint32_t)q[2 * i + 1] << 16);
                sink.template put<NOUT / 2>(o, w);
            } else {
#pragma unroll
                for (int i = 0; i < NOUT; ++i) if (i < n_valid) o[i] = (short)q[i];
            }
        } else {
            int q[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; ++i) q[i] = quantise<24>(y[i], d[i]);
            if (P.fmt == FMT_I24_IN_I32) {
                int* o = reinterpret_cast<int*>(row) + o_base;
                if (full) {
                    uint32_t w[NOUT];
#pragma unroll
                    for (int i = 0; i < NOUT; ++i) w[i] = (uint32_t)q[i];
                    sink.template put<NOUT>(o, w);
                } else {
#pragma unroll
                    for (int i = 0; i < NOUT; ++i) if (i < n_valid) o[i] = q[i];
                }
            } else {  // packed little-endian 3-byte samples
                unsigned char* o = row + o_base * 3;
                if (full) {
                    uint32_t w[NOUT * 3 / 4];
#pragma unroll
                    for (int i = 0; i < NOUT / 4; ++i) {
                        const uint32_t a = (uint32_t)q[4 * i] & 0xFFFFFFu, b = (uint32_t)q[4 * i + 1] & 0xFFFFFFu;
                        const uint32_t c = (uint32_t)q[4 * i + 2] & 0xFFFFFFu, e = (uint32_t)q[4 * i + 3] & 0xFFFFFFu;
                        w[3 * i] = a | (b << 24);
                        w[3 * i + 1] = (b >> 8) | (c << 16);
                        w[3 * i + 2] = (c >> 16) | (e << 8);
                    }
                    sink.template put<NOUT * 3 / 4>(o, w);
                } else {
#pragma unroll
                    for (int i = 0; i < NOUT; ++i) if (i < n_valid) {
                        o[3 * i] = (unsigned char)(q[i] & 0xFF);
                        o[3 * i + 1] = (unsigned char)((q[i] >> 8) & 0xFF);
                        o[3 * i + 2] = (unsigned char)((q[i] >> 16) & 0xFF);
                    }
                }
            }
        }
    } else {
        // decimating store (reference chain, stream_process.py:106): element-wise, parity path
        const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
        const uint32_t sid = (uint32_t)(P.stream_id0 + stream);
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            const long long o = o_base + i;
            if (i >= n_valid || (o % P.decim) != 0) continue;
            const long long od = o / P.decim;
            float d = 0.f;
            if (P.dither != DITHER_NONE) {
                const unsigned long long idx = P.sample_base + (unsigned long long)od, c = idx >> 2;
                const Philox4 w = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), sid, 0u, k0, k1);
                const uint32_t j = (uint32_t)idx & 3u;
                const uint32_t wj = j == 0 ? w.w[0] : (j == 1 ? w.w[1] : (j == 2 ? w.w[2] : w.w[3]));
                d = (P.dither == DITHER_TPDF) ? dither_lsb<DITHER_TPDF>(wj) : dither_lsb<DITHER_RPDF>(wj);
            }
            if (P.fmt == FMT_F32) {
                reinterpret_cast<float*>(row)[od] = (P.dither != DITHER_NONE) ? dither_float(y[i], d, P.dither_lsb) : y[i];
            } else if (P.fmt == FMT_I16) {
                reinterpret_cast<short*>(row)[od] = (short)quantise<16>(y[i], d);
            } else {
                const int q = quantise<24>(y[i], d);
                if (P.fmt == FMT_I24_IN_I32) reinterpret_cast<int*>(row)[od] = q;
                else {
                    row[3 * od] = (unsigned char)(q & 0xFF);
                    row[3 * od + 1] = (unsigned char)((q >> 8) & 0xFF);
                    row[3 * od + 2] = (unsigned char)((q >> 16) & 0xFF);
                }
            }
        }
    }
}

// L2 prefetch of the part of row `r` (element range [e0, e1)) an item will read: one bulk-prefetch
// instruction, fire and forget.  Issued for the group's NEXT item while the current one computes, so
// the HBM traffic of the load phases is spread over the FMA-bound phase instead of arriving as one
// burst per round (all groups run in lock step: same taps, same item size).
__device__ __forceinline__ void prefetch_row_l2(const float* r, int e0, int e1) {
    if (e1 <= e0) return;
    const unsigned long long a0 = reinterpret_cast<unsigned long long>(r + e0) & ~15ull;
    const unsigned long long a1 = (reinterpret_cast<unsigned long long>(r + e1) + 15ull) & ~15ull;
    // stay inside the row's own bytes: shrink to the aligned interior if rounding left it
    const unsigned long long lo = reinterpret_cast<unsigned long long>(r + e0);
    const unsigned long long hi = reinterpret_cast<unsigned long long>(r + e1);
    const unsigned long long b0 = a0 < lo ? a0 + 16 : a0;
    const unsigned long long b1 = a1 > hi ? a1 - 16 : a1;
    if (b1 <= b0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0), "r"((unsigned)(b1 - b0)) : "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void group_sync(int tg, int slot) {
    if (tg <= 32) __syncwarp();
    else if (tg == NT) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(tg) : "memory");
}

// Persistent over work items.  An item is one segment pair (row A, row B); it belongs to one thread
// group (tg threads, one shared-memory region).  Items are dealt round-robin with the CTA index
// varying fastest, so a partly filled last round leaves every SM with the same number of busy
// groups instead of leaving whole SMs idle.
struct RowItem {
    const float* r0;     // src0 row
    const float* r1;     // src1 row, pre-offset by -len0 so that logical sample n is r1[n]
    float* h;            // history row, pre-offset so that logical sample n goes to h[n]
    int start;           // logical index of pair 0 of the region
    int stream, seg;     // stream index, segment index inside the stream
    bool valid;
};

// resident CTAs per SM the register budget is compiled for: the T = 4 kernels hold few accumulators, so
// a third CTA fits and covers more of the (latency-bound) load phase; T = 8 spills at 80 registers
// (measured slower), T = 16 needs all 128
__host__ __device__ constexpr int min_ctas(int T, int NT) { return NT == 256 ? (T == 4 ? 3 : 2) : (T == 4 ? 6 : 4); }

template <int U, int T, int NT, bool MULTI, int EPI, bool GROUPED = false>
__global__ void __launch_bounds__(NT, min_ctas(T, NT))
fir_stream_kernel(const __grid_constant__ FirParams P) {
    extern __shared__ float4 smem[];
    const int tid = threadIdx.x;
    const int tg = 1 << P.tg_log2;                 // threads per segment
    const int slots = NT >> P.tg_log2;             // thread groups per CTA
    const int lq = tg * T;                         // base positions per segment
    const int kq = P.n_chunks * T;                 // padded taps per phase = halo
    const int region_pairs = kq + lq;
    const int region_g = region_granules(region_pairs, T);
    const int slot = tid >> P.tg_log2;
    const int lt = tid & (tg - 1);
    float4* region_w = smem + slot * region_g;
    const float4* region = region_w;
    const int b0 = P.n_chunks + lt;               // block holding this thread's own T positions
    const int stride = gridDim.x * slots;
    const int n4 = region_pairs >> 2;
    const int L = P.len0 + P.len1;

    for (int pair_idx = slot * gridDim.x + blockIdx.x; pair_idx < P.n_pairs; pair_idx += stride) {
        // ---- which two rows
        RowItem it[2];
        {
            const unsigned segA = (GROUPED && P.pair_single) ? (unsigned)pair_idx : 2u * (unsigned)pair_idx;
            unsigned sA, gA;
            if (P.segs_shift >= 0) { sA = segA >> P.segs_shift; gA = segA & ((1u << P.segs_shift) - 1u); }
            else { sA = segA / (unsigned)P.segs_per_stream; gA = segA - sA * (unsigned)P.segs_per_stream; }
            unsigned sB = sA, gB = gA + 1;
            if (gB == (unsigned)P.segs_per_stream) { gB = 0; sB = sA + 1; }
            it[0].stream = (int)sA; it[0].seg = (int)gA; it[0].valid = true;
            it[1].stream = (int)sB; it[1].seg = (int)gB;
            it[1].valid = (int)(segA + 1) < P.n_seg_total && !(GROUPED && P.pair_single);
            if (!it[1].valid) { it[1].stream = (int)sA; it[1].seg = (int)gA; }   // keep addresses in range
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                it[r].start = P.q_begin + it[r].seg * lq - kq;
                it[r].r0 = P.src0 + (long long)it[r].stream * P.pitch0;
                it[r].r1 = P.src1 + (long long)it[r].stream * P.pitch1 - P.len0;
                it[r].h = P.hist_out ? P.hist_out + (long long)it[r].stream * P.hist_pitch - P.hist_from : nullptr;
            }
        }

        // ---- load phase: global -> (row A, row B) interleaved pairs in shared memory, history out.
        // Every thread group loads and waits for its own region only, so groups drift apart and one
        // group's load / epilogue hides behind the others' FMA work.
        if (P.fast_all) {
            // Aligned case: every 4-sample group lies wholly in src0, in src1 or outside, so a group
            // is one predicated LDG.128.  LOAD_B groups per row are requested before any is used
            // (2*LOAD_B independent 16-byte loads in flight per thread, one memory round trip per batch).
            constexpr int LOAD_B = T == 16 ? 6 : (T == 8 ? 3 : 2);
            // per row, in units of 4-sample groups: [ld0, ld1) is inside the sequence, < split reads src0,
            // [hs0, hs1) goes to the next block's history
            int ld0[2], ld1[2], split[2], hs0[2], hs1[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int st = it[r].start;
                ld0[r] = st < 0 ? (-st) >> 2 : 0;
                ld1[r] = it[r].valid ? min(n4, (L - st) >> 2) : 0;
                split[r] = (P.len0 - st) >> 2;
                const int own = it[r].seg == 0 ? INT_MIN : st + kq;
                const int hlo = max(P.hist_from, own), hhi = P.hist_from + P.hist_len;
                hs0[r] = (it[r].h && it[r].valid) ? max((hlo - st) >> 2, 0) : 0;
                hs1[r] = (it[r].h && it[r].valid) ? min((hhi - st) >> 2, n4) : 0;
            }
            for (int base = lt; base < n4; base += tg * LOAD_B) {
                float4 v[2][LOAD_B];
#pragma unroll
                for (int k = 0; k < LOAD_B; ++k) {
                    const int i4 = base + k * tg;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        v[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i4 >= ld0[r] && i4 < ld1[r]) {
                            const float* src = (i4 < split[r] ? it[r].r0 : it[r].r1) + it[r].start + 4 * i4;
                            v[r][k] = __ldg(reinterpret_cast<const float4*>(src));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < LOAD_B; ++k) {
                    const int i4 = base + k * tg;
                    if (i4 < n4) {
                        const int g = phys_granule<T>(2 * i4);
                        region_w[g] = make_float4(v[0][k].x, v[1][k].x, v[0][k].y, v[1][k].y);
                        region_w[g + 1] = make_float4(v[0][k].z, v[1][k].z, v[0][k].w, v[1][k].w);
#pragma unroll
                        for (int r = 0; r < 2; ++r)
                            if (i4 >= hs0[r] && i4 < hs1[r])
                                *reinterpret_cast<float4*>(it[r].h + it[r].start + 4 * i4) = v[r][k];
                    }
                }
            }
        } else {
            RowSrc rs[2];
            float* hrow[2];
            int own[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                rs[r].r0 = it[r].r0; rs[r].r1 = it[r].r1 + P.len0; rs[r].valid = it[r].valid;
                hrow[r] = (it[r].h && it[r].valid) ? it[r].h + P.hist_from : nullptr;
                own[r] = it[r].seg == 0 ? INT_MIN : it[r].start + kq;
            }
#pragma unroll 2
            for (int i4 = lt; i4 < n4; i4 += tg) {
                const float4 a = fetch4(P, rs[0], it[0].start + 4 * i4);
                const float4 b = fetch4(P, rs[1], it[1].start + 4 * i4);
                const int g = phys_granule<T>(2 * i4);
                region_w[g] = make_float4(a.x, b.x, a.y, b.y);
                region_w[g + 1] = make_float4(a.z, b.z, a.w, b.w);
                if (hrow[0]) hist_store4(P, hrow[0], it[0].start + 4 * i4, a, own[0]);
                if (hrow[1]) hist_store4(P, hrow[1], it[1].start + 4 * i4, b, own[1]);
            }
        }
        group_sync<NT>(tg, slot);

        // ---- L2 prefetch of the next item's rows (lanes 0..3 of the group: row A/B x src0/src1).
        // Only worth its instructions when a region is large (long filters: history >> block).
        if (T == 16 && lt < 4 && pair_idx + stride < P.n_pairs) {
            const int seg = (GROUPED && P.pair_single) ? (pair_idx + stride) : 2 * (pair_idx + stride) + (lt >> 1);
            if (seg < P.n_seg_total) {
                const int s_ = seg / P.segs_per_stream, g_ = seg - s_ * P.segs_per_stream;
                const int n0 = P.q_begin + g_ * lq - kq, n1 = n0 + region_pairs;   // logical range
                if (lt & 1) prefetch_row_l2(P.src1 + (long long)s_ * P.pitch1, max(n0, P.len0) - P.len0,
                                            min(n1, L) - P.len0);
                else        prefetch_row_l2(P.src0 + (long long)s_ * P.pitch0, max(n0, 0), min(n1, P.len0));
            }
        }

        // ---- compute phase
        float2 acc[U][T];
#pragma unroll
        for (int p = 0; p < U; ++p)
#pragma unroll
            for (int r = 0; r < T; ++r) acc[p][r] = make_float2(0.f, 0.f);

        float2 wa[T], wb[T];
        // MULTI: one pass per EQ band accumulated into the same outputs.  The single-pass instantiation
        // keeps the tap address a pure function of the chunk counter, which is what lets ptxas hold the
        // taps in uniform registers (LDCU + FFMA2 ..., UR.F32, ...).
        const int n_sub = MULTI ? P.n_sub : 1;
        for (int sb = 0; sb < n_sub; ++sb) {
            // GROUPED: the item's own table in global memory (every lane reads the same address: one L1 broadcast per load)
            const float* tp = GROUPED ? P.gtaps + (long long)__ldg(P.group_of_stream + it[0].stream) * P.gstride
                                      : (MULTI ? P.taps + sb * P.sub_stride : P.taps);
            load_block<T>(wb, region, b0);
            int c = 0;
            for (; c + 2 <= P.n_chunks; c += 2) {
                load_block<T>(wa, region, b0 - c - 1);
                chunk_fma<U, T>(acc, wa, wb, tp + c * (T * U));
                load_block<T>(wb, region, b0 - c - 2);
                chunk_fma<U, T>(acc, wb, wa, tp + (c + 1) * (T * U));
            }
            if (c < P.n_chunks) {
                load_block<T>(wa, region, b0 - c - 1);
                chunk_fma<U, T>(acc, wa, wb, tp + c * (T * U));
            }
        }
        group_sync<NT>(tg, slot);      // the region may be overwritten by the next item's load

        // ---- epilogue: registers -> global, dither / quantise on the way.  The generic epilogue is
        // large: one copy for both rows keeps it out of the i-cache's way; specialised ones are unrolled.
#pragma unroll (EPI >= 0 ? 2 : 1)
        for (int row = 0; row < 2; ++row) {
            const bool rvalid = row == 0 ? it[0].valid : it[1].valid;
            const int rseg = row == 0 ? it[0].seg : it[1].seg;
            const int rstream = row == 0 ? it[0].stream : it[1].stream;
            if (!rvalid) continue;
            const int q_local = rseg * lq + lt * T;               // first base position (relative to q_begin)
            const int nv = min(max(P.n_q - q_local, 0), T) * U;
            float y[U * T];
#pragma unroll
            for (int r = 0; r < T; ++r)
#pragma unroll
                for (int p = 0; p < U; ++p) y[r * U + p] = row == 0 ? acc[p][r].x : acc[p][r].y;
            store_row<U * T, EPI>(P, y, rstream, (long long)q_local * U, nv);
        }
    }
}

}  // namespace afp
