// Tensor-core FIR, second form: split-FP16 operands, Hankel tap cores resident in shared memory, consecutive
// blocks of a stream group processed by one CTA (afp_process_blocks).  U = 1 filters of 48..1025 taps and 2x .. 4x polyphase
// interpolators of up to 1025 taps per phase (tc2_build_cores).
//
// Same Toeplitz blocking as afp_fir_tc.cuh - rows of R = 64 samples, y[R m + j] = sum_d sum_i X[m-d][i] W_d[i][j],
// the A operand of chunk d being the one sample tile moved up d rows through the descriptor start address -
// with three changes, each from a measurement (bench/microbench/umma_peak.cu, profiles/r02/umma_peak.txt):
//
//  1. No tap traffic.  The first kernel streams the Toeplitz blocks W_d (R x R floats per chunk, 557 KB for
//     1001 taps) from L2 for every work item: at the tensor pipe's own pace that is ~43 B/clk/SM, the whole
//     chip's L2 bandwidth.  Here the samples of a row are stored in REVERSED order, X'[m][i] = x[R m + R-1-i],
//     which turns the tap block into a Hankel matrix,  W'_d[i][j] = hp[R d + i + j],  hp[t] = h[t - (R-1)].
//     In the K-major core-matrix layout a core (8 rows j x 16 bytes of i) of a Hankel matrix depends on
//     (i0 + j0) only, so ONE array of cores  C_q[jj][e] = hp[8 q + jj + e]  serves every chunk, k-step and
//     column group: the B descriptor of (chunk d, k-step ks) starts at core index 8 d + 2 ks.  1001 taps -> 143
//     core indices x (128 B high part + 128 B low part) = 36 KB, loaded once per CTA.
//  2. FP16 split instead of TF32 split.  x = hi + lo * 2^-11 with hi = fp16(x), lo = fp16((x - hi) * 2^11),
//     the same for the taps: 11 + 11 significant bits per operand, exactly what the 3xTF32 split keeps, but
//     kind::f16 runs at twice the rate (K = 16 per instruction at the cost of tf32's K = 8) and a sample
//     needs 4 bytes of shared memory (hi + lo) instead of 8, which is what makes room for the ring of window rows
//     (below).  Products  hi*hi and hi*lo' come out of ONE N = 128 instruction: the core array holds the
//     high and the low core of every index side by side (slot 2q / 2q + 1), so a B descriptor that walks the slots
//     one by one (stride 128 B) sees [hi core q | lo core q | hi core q+1 | ...] - accumulator A, column 16 g + c =
//     x_hi * W_hi column 8 g + c, column 16 g + 8 + c = x_hi * W_lo' of the same column.  lo'*hi is an N = 64
//     instruction over the high cores alone (stride 256 B) into accumulator B.  y = A_hh + 2^-11 (A_hl + B);
//     lo*lo (2^-22 relative) is dropped as before.  An MMA costs its operand bytes at 128 B/clk (umma_peak):
//     8 KB + 6 KB = 112 clk per chunk and k-step where three N = 64 instructions took 144.  Range: |x| < 65504
//     (audio is +-1); a larger sample sets the context's range flag and the call reports it.
//  3. The CTA keeps a stream group for all the blocks of the call.  A step = 16 streams x 8 rows (512
//     samples); its window [H history rows | 8 new rows] slides over a ring of H + 16 tile rows (tc2_step; H = 16, or 8
//     for filters of up to 513 taps per phase), so the
//     loader warps bring in only the 8 new rows of step k+1 while the tensor pipe works on step k and the epilogue
//     warps drain step k-1 from TMEM (two accumulator sets).  History is read from the carried buffer only for the
//     first step of a call and written once at its end; in between the "history" is the previous blocks' input,
//     still in shared memory.
//
//  4. Warps: 1 MMA issuer, 16 loaders (two groups of 8 alternate steps), 8 epilogue.  A persistent CTA reads HBM at a rate set
//     by the number of LOADING WARPS - a warp sustains ~4 128-bit loads whatever it issues (bench/microbench/stream_mlp.cu:
//     4 / 8 / 16 warps read 3.0 / 5.1 / 6.6 TB/s) - so 16 warps x 8 loads replaced 8 x 16 (51 taps: 456 -> 553 Gsamples/s).
//     25 warps leave 72 registers per thread, which shapes the epilogue (16 columns at a time for the quantised formats).
//
// Cost model (measured): an M = 128, K = 16 MMA takes max(pipe, operand bytes / 128 B/clk): N = 64 48 clk, N = 128
// 64 clk;  a U = 1 step is n_chunks x 4 k-steps x (64 + 48) clk: 1001 taps 17 x 4 x 112 = 7.6 kclk per 8192 samples; a
// polyphase step n_chunks x 4 x 3 x 64 clk per 128 output columns: 1001 taps at 2x 9 x 768 = 6.9 kclk per 16384 samples.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "afp_fir_tc.cuh"

namespace afp {

constexpr int TC2_R = 64;                  // samples per row
constexpr int TC2_STREAMS = 16;            // streams per work item (M = 128 = 16 x 8 rows)
constexpr int TC2_ROWS = 41;               // rows per stream tile (16 history rows): a ring of 32, its first 8 repeated, + 1 (row pitch 41 x 16 B = 16 mod 128: conflict-free loader stores)
constexpr int TC2_STEP = 8 * TC2_R;        // new samples per stream and step
constexpr int TC2_LOADERS = 16, TC2_EPIS = 8;           // warps (loaders: a warp keeps ~4 128-bit loads in flight whatever it issues - bench/microbench/stream_mlp.cu - so HBM reads scale with the number of loading warps)
constexpr int TC2_THREADS = 32 * (1 + TC2_LOADERS + TC2_EPIS);
constexpr int TC2_A_LBO = TC2_ROWS * 16;                 // bytes between column chunks (8 samples) of one stream tile
constexpr int TC2_PITCH = (TC2_R / 8) * TC2_A_LBO;       // bytes per stream tile
constexpr int TC2_TILE_BYTES = TC2_STREAMS * TC2_PITCH;  // one of hi / lo

struct Tc2Params {
    const float* hist_in;      // [n_streams][hist_pitch] carried samples, oldest first (the last hist_len of the past input)
    float* hist_out;           // same layout: the last hist_len samples after this call
    long long hist_pitch;
    int hist_len;              // multiple of 4, <= 1536
    const float* in;           // block b, stream s, frame f: in[b * in_block_stride + s * in_pitch + f]
    long long in_pitch, in_block_stride;
    void* out;                 // same indexing in elements of the output format (bytes / 3 for packed int24)
    long long out_pitch, out_block_stride;
    const __half* w;           // [n_cores][2][8][8]: Hankel core q of the taps' high parts, then core q of the low parts * 2^11
    int n_cores;
    int n_streams, frames, n_blocks;
    int n_chunks;              // D + 1
    int n_groups;              // ceil(n_streams / 16)
    unsigned int* range_flag;  // set to 1 when a sample does not fit FP16
    long long* prof;           // nullptr, or [grid][16] cycle counters (bench/microbench/tc2_probe.cu): 0-5 MMA warp (total, waits for
                               // accumulators / early rows / late rows, issue, steps), 6-8 loader warp 1 (load latency, wait for the
                               // window slot, split + store), 9-11 epilogue warp 0 (wait for the MMAs, TMEM read, store)
    // GROUPED instantiation (per-stream filters, afp_assign_groups): a work item is ONE stream; its 16 tile slots are 16
    // consecutive runs of steps_per_slot steps of that stream (slot v starts where slot v-1 ends), so all 16 share the
    // stream's tap cores gw + group_of_stream[s] * gw_stride, which the loaders bring in when the group changes
    const __half* gw;
    long long gw_stride;
    const int* group_of_stream;
    int steps_per_slot;
    int epi_stage;             // 1: the launch carries TC2_STAGE_BYTES of shared memory for the epilogue's transposed stores
    // epilogue (same meaning as FirParams)
    int fmt, dither, decim, fast_out;
    unsigned long long seed, sample_base;
    int stream_id0;
    float gain, dither_lsb;
};

typedef void (*tc2_kern_t)(const Tc2Params);

__host__ __device__ constexpr int tc2_n_cores(int n_chunks, int U = 1) { return 8 * U * (n_chunks + 1) - U; }
constexpr int TC2_STAGE_STRIDE = 9;                                  // granules per staged row: 8 + 1 (odd: conflict-free)
constexpr int TC2_STAGE_BYTES = TC2_EPIS * 32 * TC2_STAGE_STRIDE * 16;  // one 32-row tile per epilogue warp
// History rows of a window: 16 cover 1025 taps per phase; filters of up to 8 chunks beyond the first (513 taps per phase -
// every polyphase case) keep 8, which shrinks the stream tile from 41 to 33 rows (a ring of 24) and makes room for the
// epilogue's staging tiles beside the larger polyphase core arrays.
__host__ __device__ constexpr int tc2_hist_rows(int n_chunks) { return n_chunks - 1 <= 8 ? 8 : 16; }
__host__ __device__ constexpr int tc2_tile_rows(int n_chunks) { return tc2_hist_rows(n_chunks) + 16 + 9; }     // ring + 8 repeated + 1
__host__ __device__ constexpr int tc2_tile_bytes(int n_chunks) { return TC2_STREAMS * (TC2_R / 8) * tc2_tile_rows(n_chunks) * 16; }
__host__ __device__ constexpr int tc2_smem_bytes(int n_chunks, bool stage = false, int U = 1, bool grouped = false) {
    // (the per-stream-filter instantiation keeps the 16-row geometry as a compile-time constant: its loaders have no
    // registers to spare for a runtime one - 170 vs 134 Gsamples/s on C3 with 4096 gain sets)
    return 2 * (grouped ? TC2_TILE_BYTES : tc2_tile_bytes(n_chunks)) + 2 * tc2_n_cores(n_chunks, U) * 128 + 256 + (stage ? TC2_STAGE_BYTES : 0);
}
static_assert(tc2_tile_bytes(17) == TC2_TILE_BYTES, "16 history rows: the 41-row tile");

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// instruction descriptor: D f32, A / B f16 K-major, M = 128
__host__ __device__ constexpr uint32_t tc2_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// Host side: the Hankel core array of a filter.  C_q[jj][e] = hp[8 q + jj + e], hp[t] = h[t - (R-1)]; every tap
// h = hi + lo * 2^-11 with hi = fp16(h), lo = fp16((h - hi) * 2^11).  Layout: hi core q in slot 2 q, lo core q in
// slot 2 q + 1, 64 halves each.  Returns the chunk count D + 1.
//
// U = 2 / 3 / 4: the polyphase interpolator  out[U n + p] = sum_k h[U k + p] x[n - k]  (zero-stuff by U, FIR at the high rate,
// only the non-zero products) as the SAME Toeplitz form with wider rows: a row of R input samples makes U R consecutive
// output samples, column c = U j + p, and the tap that joins element i of row m - d (sample R (m-d) + R-1 - i) to column c of
// row m is  h[U (R d + j + i - (R-1)) + p] = g[U R d + U i + c],  g[t] = h[t - U (R-1)].  A core (8 columns x 8 elements)
// again depends on one index only,  C_q[jj][e] = g[8 q + jj + U e]  with q = 8 U d + U (i / 8) + c / 8:  the next core in
// K is U cores on, the next 8 columns one core on, 128 columns (16 cores) per instruction, 64 U columns per row in passes
// of 128 (3x: 128 + 64).  One 64-sample window serves all U phases - the loaders do not change.  Layout: all hi cores, then all lo cores.
template <class VEC>
inline int tc2_build_cores(const double* taps, int n, VEC& w, int U = 1) {
    const int D = (n - 1 + (TC2_R - 1) * U) / (TC2_R * U), Q = tc2_n_cores(D + 1, U);
    w.assign((size_t)2 * Q * 64, __float2half_rn(0.f));
    for (int q = 0; q < Q; ++q)
        for (int jj = 0; jj < 8; ++jj)
            for (int e = 0; e < 8; ++e) {
                const int k = 8 * q + jj + U * e - U * (TC2_R - 1);
                const float hv = (k >= 0 && k < n) ? (float)taps[k] : 0.f;
                const __half hi = __float2half_rn(hv);
                const __half lo = __float2half_rn((hv - __half2float(hi)) * 2048.0f);
                if (U == 1) {
                    w[(size_t)(2 * q) * 64 + jj * 8 + e] = hi;
                    w[(size_t)(2 * q + 1) * 64 + jj * 8 + e] = lo;
                } else {
                    w[(size_t)q * 64 + jj * 8 + e] = hi;
                    w[(size_t)(Q + q) * 64 + jj * 8 + e] = lo;
                }
            }
    return D + 1;
}

// 4 consecutive samples (one float4, the low or high half of an 8-sample granule) -> 8 bytes of the hi tile and 8 of
// the lo tile.  A granule holds its samples in reversed order (element e = sample 7 - e), so within the half the
// words are (x3, x2), (x1, x0), the first of each pair in the low 16 bits.
__device__ __forceinline__ void tc2_split4(const float4& v, uint2& hi, uint2& lo, float& amax) {
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    const __half2 h32 = __floats2half2_rn(v.w, v.z), h10 = __floats2half2_rn(v.y, v.x);
    const float2 f32 = __half22float2(h32), f10 = __half22float2(h10);
    const __half2 l32 = __floats2half2_rn(__fmul_rn(__fsub_rn(v.w, f32.x), 2048.0f), __fmul_rn(__fsub_rn(v.z, f32.y), 2048.0f));
    const __half2 l10 = __floats2half2_rn(__fmul_rn(__fsub_rn(v.y, f10.x), 2048.0f), __fmul_rn(__fsub_rn(v.x, f10.y), 2048.0f));
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h32), *reinterpret_cast<const uint32_t*>(&h10));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l32), *reinterpret_cast<const uint32_t*>(&l10));
}

// Where the window of a step lives - the same function in the MMA and the loader role.
// A window is H + 8 rows [H history | 8 new], H = 16 or 8 (tc2_hist_rows).  The stream tile is a RING of H + 16 rows (the
// physical rows after it repeat rows 0..7, so the eight consecutive rows an MMA reads never wrap): step k of a stream group
// has its window at ring rows 8 k .. 8 k + H + 7 (mod H + 16) and writes only its 8 NEW rows, into the eight rows that are not
// part of the previous step's window - the oldest rows of the step before that, free once that step's MMAs are done.  Only
// the first step of a group loads a whole window (chunk d reads window rows H - d .. H + 7 - d).
// (The first version walked three fixed positions of a 40-row tile and reloaded all 24 rows on every third step; the
// probe counters showed the MMA warp waiting for exactly those reloads: 6.1 of 8.0 kclk per step at 51 taps.)
struct Tc2Step { int pos; bool fresh; };
// Window start of the next step given the previous step's: +8 inside a group; a new group starts a whole window (H + 8
// rows) on, so that its eight OLDEST rows are the ring rows the previous step's window does not cover.  `w` starts at 8.
__device__ __forceinline__ Tc2Step tc2_step(int& w, int k, int H) {      // (ring of H + 16 rows, window of H + 8)
    w += k == 0 ? H + 8 : 8;
    if (w >= H + 16) w -= H + 16;
    Tc2Step st;
    st.pos = w;
    st.fresh = k == 0;
    return st;
}

// The carried samples of the next call: the last hist_len samples of [history | all blocks] of the 16 streams from s0
// on, copied global -> global (L2 hits) by `nthreads` threads.  Loads go in batches of 8 ahead of their stores (the
// compiler cannot reorder them itself: the buffers may alias).
__device__ __forceinline__ void tc2_copy_history(const Tc2Params& P, int s0, int n_rows, int total, int ltid, int nthreads) {
    const int q4 = P.hist_len >> 2, n_copy = n_rows * q4;
    for (int base = ltid; base < n_copy; base += 8 * nthreads) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = base + u * nthreads;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n_copy) {
                const int sl = j / q4, i4 = j - sl * q4, s = s0 + sl;
                const int t = total - P.hist_len + 4 * i4;
                if (s < P.n_streams) {
                    const float* src;
                    if (t < 0) src = P.hist_in + (long long)s * P.hist_pitch + (P.hist_len + t);
                    else {
                        const int b_ = t / P.frames, f = t - b_ * P.frames;
                        src = P.in + (long long)b_ * P.in_block_stride + (long long)s * P.in_pitch + f;
                    }
                    v[u] = __ldg(reinterpret_cast<const float4*>(src));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = base + u * nthreads;
            if (j < n_copy) {
                const int sl = j / q4, i4 = j - sl * q4, s = s0 + sl;
                if (s < P.n_streams) reinterpret_cast<float4*>(P.hist_out + (long long)s * P.hist_pitch)[i4] = v[u];
            }
        }
    }
}

// Epilogue store sink: warp transpose through shared memory.  A lane owns one row's run of NB words (one 128-byte line for
// 4-byte samples), and the 32 rows of a warp lie 256 bytes apart: stored straight from registers, every STG.128 puts 16
// bytes into 32 different lines - the partial-line traffic slowed the loaders' reads as much as the epilogue itself
// (probe without the stores: 383 -> 545 Gsamples/s at 51 taps).  Here the lanes park their runs in the warp's tile and
// read them back so that eight consecutive lanes store one row's eight granules: every instruction writes four whole
// lines.  Row addresses travel by shuffle; rows outside the batch carry a null address.
struct RowTransposeSink {
    static constexpr bool kPhiloxSideBySide = true;
    static constexpr bool kAllFull = true;    // (store_row's aligned full-run path; the caller checks fast_out)
    uint4* tile;
    bool ok;
    template <int NB>
    __device__ __forceinline__ void put(void* dst, const uint32_t (&w)[NB]) const {
        constexpr int G = NB / 4;                               // granules per row (packed int24: 3 per 16 samples)
        if (NB % 4 != 0 || G > 8) {
            if (ok) store_words<NB>(dst, w);
            return;
        }
        const int lane = threadIdx.x & 31;
        __syncwarp();                                           // the previous read-back is done
#pragma unroll
        for (int i = 0; i < G; ++i)
            tile[lane * TC2_STAGE_STRIDE + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        __syncwarp();
        const unsigned long long d = ok ? reinterpret_cast<unsigned long long>(dst) : 0ull;
        const unsigned dlo = (unsigned)d, dhi = (unsigned)(d >> 32);
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int l = i * 32 + lane, r = l / G, g = l - r * G;   // the warp's 32 G granules in row-major order, 32 per instruction
            const unsigned long long p = ((unsigned long long)__shfl_sync(0xffffffffu, dhi, r) << 32) | __shfl_sync(0xffffffffu, dlo, r);
            if (p) reinterpret_cast<uint4*>(p)[g] = tile[r * TC2_STAGE_STRIDE + g];
        }
    }
};

template <int EPI, bool GROUPED = false, int U = 1>
__global__ void __launch_bounds__(TC2_THREADS, 1)
fir_tc2_kernel(const __grid_constant__ Tc2Params P) {
    extern __shared__ __align__(128) unsigned char tc2_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* tile_hi = tc2_smem;
    const int H = GROUPED ? 16 : tc2_hist_rows(P.n_chunks);   // history rows of a window (16 cover 1025 taps per phase; up to 513: 8)
    const int RING = H + 16, ROWS = RING + 9, A_LBO = ROWS * 16, PITCH = (TC2_R / 8) * A_LBO, TILE = TC2_STREAMS * PITCH;
    unsigned char* tile_lo = tile_hi + TILE;
    unsigned char* w_hi = tile_lo + TILE;       // core slots: [hi q | lo q], 256 B per index
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_hi + 2 * P.n_cores * 128);
    uint64_t* win_early = bars;         // [2] step parity p: the rows the first chunks read are in place
    uint64_t* win_late = bars + 2;      // [2] ... and the rows only the later chunks read
    uint64_t* mma_done = bars + 4;      // [2] step parity p: the window of the step has been read for the last time
    uint64_t* acc_empty = bars + 6;     // [2] accumulator set drained
    uint64_t* acc_full = bars + 8;      // [2] accumulator set final (U = 4: a step fills a set twice, 128 output columns at a time)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
    static_assert(U >= 1 && U <= 4, "polyphase factors of the tensor kernel");
    static_assert(!GROUPED || U == 1, "per-stream tap cores: U = 1 only");
    constexpr int PASSES = U >= 3 ? 2 : 1;   // accumulator fills per step: 128 output columns each (U = 3: 128 + 64)

    const int n_steps = GROUPED ? P.steps_per_slot : (P.frames / TC2_STEP) * P.n_blocks;   // steps of one work item

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&win_early[i], TC2_LOADERS / 2); mbar_init(&win_late[i], TC2_LOADERS / 2);   // the owning loader group's warps
            mbar_init(&mma_done[i], 1); mbar_init(&acc_empty[i], TC2_EPIS); mbar_init(&acc_full[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // tap cores: global -> shared, once
        const uint4* src = reinterpret_cast<const uint4*>(P.w);
        uint4* dst = reinterpret_cast<uint4*>(w_hi);
        for (int i = tid; i < 2 * P.n_cores * 8; i += TC2_THREADS) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

    // (Tried: roles in whole warpgroups and setmaxnreg 56 / 104 for loaders / epilogue - the loaders' spills at 56 registers cost
    // more than the epilogue gained: 553 -> 461 Gsamples/s at 51 taps.  25 warps compile to 72 registers each.)
    if (warp_u == 0) {
        // ================= MMA issuer: one elected lane, operands warp-uniform =================
        const uint64_t a_hi_d = umma_desc(smem_u32(tile_hi), A_LBO, PITCH), a_lo_d = umma_desc(smem_u32(tile_lo), A_LBO, PITCH);
        // B: the next core in K is the next index (256 B on); in N the wide product walks the slots (hi q, lo q, hi q+1, ...),
        // the narrow one the high cores only
        // U = 1.  B: the next core in K is the next index (256 B on); in N the wide product walks the slots (hi q, lo q, hi q+1,
        // ...), the narrow one the high cores only.
        // U = 2 .. 4 (polyphase interpolator, file header of tc2_build_cores): all high cores, then all low cores; the next core
        // in K is U cores on, in N one core on; 128 output columns per instruction.
        const uint32_t lo_off = U == 1 ? 0u : (uint32_t)P.n_cores * 128u;
        const uint64_t w_wide_d = U == 1 ? umma_desc(smem_u32(w_hi), 256, 128) : umma_desc(smem_u32(w_hi) + lo_off, U * 128, 128);
        const uint64_t w_hi_d = U == 1 ? umma_desc(smem_u32(w_hi), 256, 256) : umma_desc(smem_u32(w_hi), U * 128, 128);
        const uint32_t top_a = (uint32_t)(a_hi_d >> 32), top_ww = (uint32_t)(w_wide_d >> 32), top_wh = (uint32_t)(w_hi_d >> 32);
        const uint32_t idesc_wide = tc2_idesc(2 * TC2_R), idesc = tc2_idesc(TC2_R);
        const int D = P.n_chunks - 1;
        uint32_t kk = 0, vk = 0;                               // steps / accumulator fills issued by this CTA
        int wpos = 8;
        long long t_acc = 0, t_dir = 0, t_stg = 0, t_iss = 0;  // cycles waiting for: accumulators, early rows, late rows; issuing
        const long long t_begin = clock64();
        for (int grp = blockIdx.x; grp < P.n_groups; grp += gridDim.x) {
            for (int k = 0; k < n_steps; ++k, ++kk) {
                const uint32_t p = kk & 1u, ph = (kk >> 1) & 1u;
                const Tc2Step st = tc2_step(wpos, k, H);
                bool late_seen = false;
#pragma unroll 1
                for (int h = 0; h < PASSES; ++h, ++vk) {
                    const uint32_t set = vk & 1u, aph = (vk >> 1) & 1u;
                    const long long c0 = clock64();
                    mbar_wait(&acc_empty[set], aph ^ 1u);      // (passes at once the first time round)
                    const long long c1 = clock64();
                    if (h == 0) mbar_wait(&win_early[p], ph);
                    const long long c2 = clock64();
                    t_acc += c1 - c0; t_dir += c2 - c1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t acc_a = tmem_base + set * 256u, acc_b = acc_a + 128u;
                    bool first = true;
                    // Chunk order, the same for every step so that rounding does not depend on where a step ran:
                    // 8, D..9, 7..0 (what exists of it).  Chunk d reads window rows 16-d .. 23-d: the chunks >= 8 only history
                    // rows - a sliding step starts them while its new rows are still on their way.
                    for (int idx = 0; idx <= D; ++idx) {
                        int d;
                        if (D >= 8) d = idx == 0 ? 8 : (idx <= D - 8 ? D + 1 - idx : D - idx);
                        else d = D - idx;
                        const bool late = d <= 7;              // reads new rows (a first step publishes both barriers together)
                        if (late && !late_seen) {
                            const long long w0 = clock64();
                            mbar_wait(&win_late[p], ph);
                            t_stg += clock64() - w0;
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            late_seen = true;
                        }
                        if (elect_one()) {
                            const uint32_t arow = (uint32_t)(st.pos + H - d >= RING ? st.pos + H - d - RING : st.pos + H - d);
                            const uint32_t ah = (uint32_t)a_hi_d + arow, al = (uint32_t)a_lo_d + arow;
                            if (U == 1) {
                                const uint32_t ww = (uint32_t)w_wide_d + (uint32_t)(8 * d) * 16u, wh = (uint32_t)w_hi_d + (uint32_t)(8 * d) * 16u;
#pragma unroll
                                for (int ks = 0; ks < TC2_R / 16; ++ks) {
                                    const uint32_t ao = (uint32_t)(2 * ks * ROWS), wo = (uint32_t)(2 * ks) * 16u;
                                    const uint32_t f = (first && ks == 0) ? 0u : 1u;
                                    umma_f16(acc_a, ah + ao, top_a, ww + wo, top_ww, idesc_wide, f);   // x_hi * [W_hi | W_lo'] (interleaved by 8 columns)
                                    umma_f16(acc_b, al + ao, top_a, wh + wo, top_wh, idesc, f);        // x_lo' * W_hi
                                }
                            } else {
                                // core index of (chunk d, k-step ks, pass h): 8 U d + 2 U ks + 16 h; 8 descriptor units per core
                                const uint32_t c0_ = (uint32_t)(8 * U * d + 16 * h) * 8u;
                                const uint32_t wl = (uint32_t)w_wide_d + c0_, wh = (uint32_t)w_hi_d + c0_;
                                const uint32_t idesc_p = (U == 3 && h == 1) ? idesc : idesc_wide;   // 3x: the second pass has 64 columns
#pragma unroll
                                for (int ks = 0; ks < TC2_R / 16; ++ks) {
                                    const uint32_t ao = (uint32_t)(2 * ks * ROWS), wo = (uint32_t)(2 * U * ks) * 8u;
                                    const uint32_t f = (first && ks == 0) ? 0u : 1u;
                                    umma_f16(acc_a, ah + ao, top_a, wh + wo, top_wh, idesc_p, f);    // x_hi * W_hi   -> accumulator 0
                                    umma_f16(acc_b, ah + ao, top_a, wl + wo, top_ww, idesc_p, f);    // x_hi * W_lo'  -> accumulator 1
                                    umma_f16(acc_b, al + ao, top_a, wh + wo, top_wh, idesc_p, 1u);   // x_lo' * W_hi  -> accumulator 1
                                }
                            }
                        }
                        __syncwarp();
                        first = false;
                    }
                    if (!late_seen) { mbar_wait(&win_late[p], ph); late_seen = true; }
                    if (elect_one()) {
                        umma_commit(&acc_full[set]);
                        if (h == PASSES - 1) umma_commit(&mma_done[p]);
                    }
                    __syncwarp();
                    t_iss += clock64() - c2;
                }
            }
        }
        if (P.prof && lane == 0) {
            long long* o = P.prof + (size_t)blockIdx.x * 16;
            o[0] = clock64() - t_begin; o[1] = t_acc; o[2] = t_dir; o[3] = t_stg; o[4] = t_iss; o[5] = kk;
        }
    } else if (warp_u <= TC2_LOADERS) {
        // ================= loaders: global -> hi / lo halves of the window rows =================
        // Two groups of four warps; group g owns the steps of parity g (the accumulator set and barrier pair of that
        // parity), so while one group splits and stores the rows of step k the other already waits for the loads of step
        // k+1: a global-memory round trip (2-3.5 kclk under load) overlaps the other group's arithmetic instead of
        // sitting between two steps.
        const int ltid = tid - 32;
        constexpr int LT = 32 * TC2_LOADERS / 2;           // threads of a loader group: 256
        const int lg = ltid / LT, lt = ltid % LT;
        const int sh = lt >> 7;                            // which streams of a pair this thread takes
        uint32_t kk = 0;
        float amax = 0.f;
        float4 v[8];                                       // this thread's loads: in flight from one own step to the next
        bool pre_fresh = false;                            // v holds this thread's share of the next group's first window
        long long t_ld = 0, t_slot = 0, t_place = 0;       // probe counters (P.prof): remaining load latency, wait for the slot, split + store
        int cur_group = -1, wpos = 8;
        for (int grp = blockIdx.x; grp < P.n_groups; grp += gridDim.x) {
            const int D = P.n_chunks - 1;
            // 8 window rows for all 16 streams: 2048 float4, 8 per thread of the owning group.  Thread lt takes float4 (lt & 15)
            // of row (lt >> 4) & 7 of every second stream (consecutive lanes: consecutive float4 of a stream's rows, which are
            // contiguous in memory), so row, column and - one filter set per batch - the block and frame of the samples are the
            // same for its 8 loads: one address computation per batch, then a stream pitch per load.
            const int prow = (lt >> 4) & 7, pf4 = lt & 15;
            // one float4 (4 samples) of window row `wrow` (0..23) of local stream sl at step k
            auto fetch_of = [&](int g, int k, int sl, int wrow, int f4) -> float4 {
                const int s = GROUPED ? g : g * TC2_STREAMS + sl;
                // logical sample index of the stream: < 0 = carried history
                const int t = ((GROUPED ? sl * n_steps : 0) + k) * TC2_STEP - H * TC2_R + wrow * TC2_R + f4 * 4;
                if (s >= P.n_streams || t < -P.hist_len) return make_float4(0.f, 0.f, 0.f, 0.f);
                const float* src;
                if (t < 0) src = P.hist_in + (long long)s * P.hist_pitch + (P.hist_len + t);
                else {
                    const int b_ = t / P.frames, f = t - b_ * P.frames;
                    src = P.in + (long long)b_ * P.in_block_stride + (long long)s * P.in_pitch + f;
                }
                return __ldg(reinterpret_cast<const float4*>(src));
            };
            // One filter set per batch: a thread's loads of one batch share row, column, block and frame - ONE address
            // computation (an integer division by the block length among it), then a stream pitch per load.  (Written as
            // eight fetch_of calls the compiler repeats the division per load: ~30 instructions each.)
            struct Loc { const float* base; long long pitch; };          // base == nullptr: before the carried samples - zeros
            auto locate = [&](int k, int wrow, int f4) -> Loc {
                const int t = k * TC2_STEP - H * TC2_R + wrow * TC2_R + f4 * 4;
                if (t < -P.hist_len) return Loc{nullptr, 0};
                if (t < 0) return Loc{P.hist_in + (P.hist_len + t), P.hist_pitch};
                const int b_ = t / P.frames, f = t - b_ * P.frames;
                return Loc{P.in + (long long)b_ * P.in_block_stride + f, P.in_pitch};
            };
            auto fetch_at = [&](const Loc& L, int g, int sl) -> float4 {
                const int s = g * TC2_STREAMS + sl;
                return (L.base && s < P.n_streams) ? __ldg(reinterpret_cast<const float4*>(L.base + (long long)s * L.pitch))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto fetch = [&](int k, int sl, int wrow, int f4) -> float4 { return fetch_of(grp, k, sl, wrow, f4); };
            // a first window is loaded by both groups: this thread's streams there are fs + 4 u; its share of the two newest
            // batches: rows 16.. in v[0..3], rows 8.. in v[4..7]
            const int fs = 2 * lg + sh;
            const int grp_next = grp + gridDim.x;
            // the two newest batches of group g's first window into v (rows H.. in v[0..3], rows H-8.. in v[4..7])
            auto fresh_fetch_all = [&](int g) {
                if (GROUPED) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u] = fetch_of(g, 0, 4 * u + fs, H + prow, pf4);
                        v[4 + u] = fetch_of(g, 0, 4 * u + fs, H - 8 + prow, pf4);
                    }
                } else {
                    const Loc a = locate(0, H + prow, pf4), b = locate(0, H - 8 + prow, pf4);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { v[u] = fetch_at(a, g, 4 * u + fs); v[4 + u] = fetch_at(b, g, 4 * u + fs); }
                }
            };
            auto place = [&](int pos, int sl, int wrow, int f4, const float4& v) {
                uint2 hi, lo;
                tc2_split4(v, hi, lo, amax);
                const int row = pos + wrow >= RING ? pos + wrow - RING : pos + wrow;
                const int off = sl * PITCH + (7 - (f4 >> 1)) * A_LBO + row * 16 + ((f4 & 1) ? 0 : 8);
                *reinterpret_cast<uint2*>(tile_hi + off) = hi;
                *reinterpret_cast<uint2*>(tile_lo + off) = lo;
                if (row < 8) {                             // rows 0..7 are repeated after row 31
                    *reinterpret_cast<uint2*>(tile_hi + off + RING * 16) = hi;
                    *reinterpret_cast<uint2*>(tile_lo + off + RING * 16) = lo;
                }
            };
            auto publish = [&](uint64_t* bar) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);
            };
            // What this thread loads next, issued right after it has published a step (never in front of a publication: under
            // load the issue of sixteen loads itself takes kclk): the rows of its group's next own step `kn`, or - none left -
            // its share of the next stream group's first window, which then finds its two newest batches on their way.
            auto next_loads = [&](int kn) {
                if (kn < n_steps) {
                    if (GROUPED) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = fetch(kn, 2 * u + sh, H + prow, pf4);
                    } else {
                        const Loc a = locate(kn, H + prow, pf4);
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = fetch_at(a, grp, 2 * u + sh);
                    }
                } else if (grp_next < P.n_groups && D < 9) {
                    fresh_fetch_all(grp_next);
                    pre_fresh = true;
                }
            };
            // GROUPED: this stream's tap cores replace the previous stream's unless they share a group
            bool new_taps = false;
            const uint4* wsrc = nullptr;
            if (GROUPED) {
                const int g = __ldg(P.group_of_stream + grp);
                new_taps = g != cur_group;
                cur_group = g;
                wsrc = reinterpret_cast<const uint4*>(P.gw + (long long)g * P.gw_stride);
            }
            // The loads of a step are issued as soon as the thread's registers are free: right after the publication of its
            // previous own step (next_loads), so they travel while the other group works.
            for (int k = 0; k < n_steps; ++k, ++kk) {
                const uint32_t p = kk & 1u;
                const Tc2Step st = tc2_step(wpos, k, H);
                if (st.fresh) {
                    // First step of a group: the window rows its MMAs read, [16 - D, 24), in batches of 8 rows.  Both loader
                    // groups work on it (a thread takes one float4 position of every second stream), the two newest batches
                    // in flight together and before the tile is free (every earlier MMA done); the oldest batch only exists
                    // for filters of more than 513 taps.  GROUPED: this stream's tap cores replace the previous ones here too.
                    // The groups meet at a named barrier, then the owning group publishes.
                    // Who may wait for what: an mbarrier wait names a phase by its PARITY, so a thread must never ask about a
                    // phase while the barrier is still two uses behind it (the answer would be the completed phase of equal
                    // parity, two steps too early).  Each loader group follows the barrier of its own step parity use by use
                    // (it waits on it at every step it owns), but not the other one.  So here the group that owns step kk-1
                    // waits for it, the group that owns step kk-2 for that, and a named barrier carries the news across.
                    if (D >= 9) {
                        // the oldest batch lies outside the previous step's window: free once step kk-2 is done
                        // (its loads use the registers a prefetched window would sit in: next_loads prefetches for D < 9 only)
                        if (GROUPED) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = fetch(k, 4 * u + fs, prow, pf4);
                        } else {
                            const Loc a = locate(k, prow, pf4);
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = fetch_at(a, grp, 4 * u + fs);
                        }
                        if ((int)p == lg && kk >= 2) mbar_wait(&mma_done[p], ((kk - 2) >> 1) & 1u);
                        asm volatile("bar.sync 2, %0;" ::"r"(32 * TC2_LOADERS) : "memory");
#pragma unroll
                        for (int u = 0; u < 4; ++u) place(st.pos, 4 * u + fs, prow, pf4, v[u]);
                    }
                    if (!pre_fresh) fresh_fetch_all(grp);
                    pre_fresh = false;
                    if ((int)p != lg && kk >= 1) mbar_wait(&mma_done[p ^ 1u], ((kk - 1) >> 1) & 1u);
                    asm volatile("bar.sync 3, %0;" ::"r"(32 * TC2_LOADERS) : "memory");
                    if (GROUPED && new_taps) {
                        const int n_w = 2 * P.n_cores * 8;
                        for (int j = ltid; j < n_w; j += 2 * LT) reinterpret_cast<uint4*>(w_hi)[j] = __ldg(wsrc + j);
                    }
                    // this group's next own step: k + 1 for the group that does not own the fresh step, k + 2 for the owner
                    const int kn = (int)p == lg ? k + 2 : k + 1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        place(st.pos, 4 * u + fs, H + prow, pf4, v[u]);
                        place(st.pos, 4 * u + fs, H - 8 + prow, pf4, v[4 + u]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"r"(32 * TC2_LOADERS) : "memory");
                    if ((int)p == lg && lane == 0) { mbar_arrive(&win_early[p]); mbar_arrive(&win_late[p]); }
                    next_loads(kn);
                    continue;
                }
                if ((int)p != lg) continue;                 // the other group's step
                {
                    const long long q0 = P.prof ? clock64() : 0;
                    if (P.prof) {                                       // (probe only: make the loads land here)
#pragma unroll
                        for (int u = 0; u < 8; ++u) asm volatile("" ::"f"(v[u].x), "f"(v[u].w));
                    }
                    const long long q1 = P.prof ? clock64() : 0;
                    // the new rows replace the oldest rows of step kk-2
                    if (kk >= 2) mbar_wait(&mma_done[p], ((kk - 2) >> 1) & 1u);
                    // The history rows are where the last steps left them.  (Not before the wait above: this group comes
                    // here straight from step kk-2, and an arrival that completes this barrier's phase kk before the MMA
                    // warp has looked at phase kk-2 would make that wait see the old parity again - a deadlock.)
                    if (lane == 0) mbar_arrive(&win_early[p]);
                    const long long q2 = P.prof ? clock64() : 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) place(st.pos, 2 * u + sh, H + prow, pf4, v[u]);
                    publish(&win_late[p]);
                    if (P.prof) { t_ld += q1 - q0; t_slot += q2 - q1; t_place += clock64() - q2; }
                    next_loads(k + 2);
                }
            }
        }
        if (amax > 65000.0f) atomicOr(P.range_flag, 1u);
        if (P.prof && ltid == 0) { long long* o = P.prof + (size_t)blockIdx.x * 16; o[6] = t_ld; o[7] = t_slot; o[8] = t_place; }
    } else {
        // ================= epilogue: TMEM -> registers -> dither / quantise -> HBM =================
        const int ew = warp_u - 1 - TC2_LOADERS;                // 0..7
        const int quarter = warp_u & 3;                         // the TMEM lanes this warp may read
        const int half = ew >> 2;                               // columns 32*half .. +31 of each 64-column accumulator
        const int m = 32 * quarter + lane;                      // accumulator row: stream m/8 of the group, row m%8 of the step
        struct EpiView {   // what store_row reads; out / sample_base move with the block
            int fmt, dither, decim, fast_out;
            float gain, dither_lsb;
            void* out; long long out_pitch;
            unsigned long long seed, sample_base;
            int stream_id0;
        } V = {P.fmt, P.dither, 1, P.fast_out, P.gain, P.dither_lsb, P.out, P.out_pitch, P.seed, P.sample_base, P.stream_id0};
        const int eb = P.fmt == FMT_I16 ? 2 : (P.fmt == FMT_I24_PACKED ? 3 : 4);
        uint4* stage_tile = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(bars) + 256) + ew * 32 * TC2_STAGE_STRIDE;
        uint32_t kk = 0;
        long long e_wait = 0, e_tmem = 0, e_store = 0;      // probe counters (P.prof)
        for (int grp = blockIdx.x; grp < P.n_groups; grp += gridDim.x) {
            const int s = GROUPED ? grp : grp * TC2_STREAMS + (m >> 3);
            for (int k = 0; k < n_steps; ++k, ++kk) {
              if constexpr (U == 1) {
                const uint32_t p = kk & 1u, ph = (kk >> 1) & 1u;
                // first output sample of this thread's 32 within the call: slot m/8 runs through its own steps (GROUPED)
                const int idx = ((GROUPED ? (m >> 3) * n_steps : 0) + k) * TC2_STEP + (m & 7) * TC2_R + 32 * half;
                const int blk = idx / P.frames;
                const long long r0 = P.prof ? clock64() : 0;
                mbar_wait(&acc_full[p], ph);
                const long long r1 = P.prof ? clock64() : 0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // accumulator A columns 64 half .. +63 (this half's 32 output columns, hi*hi and hi*lo' interleaved by 8),
                // accumulator B columns 32 half .. +31, 16 output columns at a time
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + p * 256u;
                V.out = static_cast<unsigned char*>(P.out) + (long long)blk * P.out_block_stride * eb;
                V.sample_base = P.sample_base + (unsigned long long)blk * (unsigned long long)P.frames;
                long long r2 = r1;
                constexpr bool kWholeRow = EPI == epi_code(FMT_F32, DITHER_NONE);
                if (kWholeRow) {
                    // plain float output: the row's 32 columns at once - one 128-byte line per row and instruction in the
                    // transposed store (553 vs 409 Gsamples/s at 51 taps against the two-pass form below)
                    float y[32];
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        uint32_t b[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
                              "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                            : "r"(taddr + 128u + (uint32_t)(32 * half + 16 * f)));
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t a[16];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                                  "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                                : "r"(taddr + (uint32_t)(64 * half + 32 * f + 16 * g)));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 8; ++c)                                    // y = A_hh + 2^-11 (A_hl + B)
                                y[16 * f + 8 * g + c] = __fmaf_rn(__fadd_rn(__uint_as_float(a[8 + c]), __uint_as_float(b[8 * g + c])),
                                                                  4.8828125e-4f, __uint_as_float(a[c]));
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[p]);
                    if (P.prof) r2 = clock64();
                    if (!(P.prof && P.gain == 0.f)) {
                        if (P.epi_stage && P.fast_out)
                            store_row<32, EPI>(V, y, s, (long long)(idx - blk * P.frames), 32, RowTransposeSink{stage_tile, s < P.n_streams});
                        else
                            store_row<32, EPI>(V, y, s, (long long)(idx - blk * P.frames), s < P.n_streams ? 32 : 0);
                    }
                } else {
                    // 16 output columns at a time, each finished (dither, codes, stores) before the next is read: a row's 32 floats
                    // next to the dither values and codes do not fit the 72 registers 25 warps leave a thread, and a spilled value
                    // costs an L1 round trip behind the kernel's own global traffic (ncu: long-scoreboard stalls on the reloads).
                    // The accumulator set goes back to the MMA warp after the second read.
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const long long t0 = P.prof ? clock64() : 0;
                        float y[16];
                        uint32_t b[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
                              "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                            : "r"(taddr + 128u + (uint32_t)(32 * half + 16 * f)));
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t a[16];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                                  "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                                : "r"(taddr + (uint32_t)(64 * half + 32 * f + 16 * g)));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 8; ++c)                                    // y = A_hh + 2^-11 (A_hl + B)
                                y[8 * g + c] = __fmaf_rn(__fadd_rn(__uint_as_float(a[8 + c]), __uint_as_float(b[8 * g + c])),
                                                         4.8828125e-4f, __uint_as_float(a[c]));
                        }
                        if (f == 1) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[p]);
                        }
                        if (P.prof) r2 += clock64() - t0;
                        if (!(P.prof && P.gain == 0.f)) {    // (probe: gain 0 = run without the stores)
                            if (P.epi_stage && P.fast_out)
                                store_row<16, EPI>(V, y, s, (long long)(idx - blk * P.frames + 16 * f), 16, RowTransposeSink{stage_tile, s < P.n_streams});
                            else
                                store_row<16, EPI>(V, y, s, (long long)(idx - blk * P.frames + 16 * f), s < P.n_streams ? 16 : 0);
                        }
                    }
                }
                if (P.prof) { e_wait += r1 - r0; e_tmem += r2 - r1; e_store += clock64() - r1 - (r2 - r1); }
              } else {
                // Polyphase interpolator: a step fills an accumulator set PASSES times with 128 output columns; this thread owns
                // row m and columns 64 half .. +63 of each fill = 64 consecutive output samples of its stream.
                const int ti = k * TC2_STEP + (m & 7) * TC2_R;                  // first input sample of the row within the call
                const int blk = ti / P.frames;
                V.out = static_cast<unsigned char*>(P.out) + (long long)blk * P.out_block_stride * eb;
                V.sample_base = P.sample_base + (unsigned long long)blk * (unsigned long long)(U * P.frames);
                constexpr bool kWholeRow = EPI == epi_code(FMT_F32, DITHER_NONE);
                constexpr int W = kWholeRow ? 32 : 16;                          // output samples per store pass
#pragma unroll 1
                for (int h = 0; h < PASSES; ++h) {
                    const uint32_t vk = kk * PASSES + h, set = vk & 1u, aph = (vk >> 1) & 1u;
                    const long long r0 = P.prof ? clock64() : 0;
                    mbar_wait(&acc_full[set], aph);
                    const long long r1 = P.prof ? clock64() : 0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int cols_half = (U == 3 && h == 1) ? 32 : 64;       // this thread's columns of the fill (3x: 128 + 64 columns)
                    const int nf = cols_half / W;
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + set * 256u + (uint32_t)(cols_half * half);
                    const long long o0 = (long long)U * (ti - blk * P.frames) + 128 * h + cols_half * half;   // within the output block
                    long long r2 = 0;
#pragma unroll
                    for (int f = 0; f < 64 / W; ++f) {
                        if (U == 3 && f >= nf) break;
                        const long long t0 = P.prof ? clock64() : 0;
                        float y[W];
#pragma unroll
                        for (int g = 0; g < W / 16; ++g) {
                            uint32_t a0[16], a1[16];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]), "=r"(a0[4]), "=r"(a0[5]), "=r"(a0[6]), "=r"(a0[7]),
                                  "=r"(a0[8]), "=r"(a0[9]), "=r"(a0[10]), "=r"(a0[11]), "=r"(a0[12]), "=r"(a0[13]), "=r"(a0[14]), "=r"(a0[15])
                                : "r"(taddr + (uint32_t)(W * f + 16 * g)));
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]), "=r"(a1[4]), "=r"(a1[5]), "=r"(a1[6]), "=r"(a1[7]),
                                  "=r"(a1[8]), "=r"(a1[9]), "=r"(a1[10]), "=r"(a1[11]), "=r"(a1[12]), "=r"(a1[13]), "=r"(a1[14]), "=r"(a1[15])
                                : "r"(taddr + 128u + (uint32_t)(W * f + 16 * g)));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 16; ++c)                               // y = acc0 + 2^-11 acc1
                                y[16 * g + c] = __fmaf_rn(__uint_as_float(a1[c]), 4.8828125e-4f, __uint_as_float(a0[c]));
                        }
                        if (f == nf - 1) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[set]);
                        }
                        if (P.prof) r2 += clock64() - t0;
                        if (!(P.prof && P.gain == 0.f)) {
                            if (P.epi_stage && P.fast_out)
                                store_row<W, EPI>(V, y, s, o0 + W * f, W, RowTransposeSink{stage_tile, s < P.n_streams});
                            else
                                store_row<W, EPI>(V, y, s, o0 + W * f, s < P.n_streams ? W : 0);
                        }
                    }
                    if (P.prof) { e_wait += r1 - r0; e_tmem += r2; e_store += clock64() - r1 - r2; }
                }
              }
            }
            // (the epilogue warps have the slack for this; in the loaders it would delay the next group's first window)
            if (GROUPED) tc2_copy_history(P, grp, 1, TC2_STREAMS * n_steps * TC2_STEP, tid - 32 * (1 + TC2_LOADERS), 32 * TC2_EPIS);
            else tc2_copy_history(P, grp * TC2_STREAMS, TC2_STREAMS, n_steps * TC2_STEP, tid - 32 * (1 + TC2_LOADERS), 32 * TC2_EPIS);
        }
        if (P.prof && tid == 32 * (1 + TC2_LOADERS)) { long long* o = P.prof + (size_t)blockIdx.x * 16; o[9] = e_wait; o[10] = e_tmem; o[11] = e_store; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace afp
