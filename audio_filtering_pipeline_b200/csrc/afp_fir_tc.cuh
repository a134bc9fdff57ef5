// Toeplitz-tiled tensor-core FIR for long filters (513..1025 taps, U = 1): tcgen05.mma kind::tf32 with
// FP32 accumulators in TMEM and 3xTF32 split precision.
//
// Formulation.  Cut every stream into rows of R = 64 samples, X[r][i] = x[R r + i].  For an output row m
//     y[R m + j] = sum_d sum_i  X[m - d][i] * W_d[i][j],      W_d[i][j] = h[R d + j - i]  (0 outside [0,N))
// i.e. a GEMM  Y[M x R] += A_d[M x R] * W_d[R x R]  over d = 0..D (D = ceil((N+R-2)/R)), whose A
// operand for chunk d is the same row-major sample matrix shifted up by d rows.  No im2col copy exists
// anywhere: the samples of a stream sit ONCE in shared memory in the un-swizzled K-major core-matrix
// layout [8 column chunks of 4 floats][rows][16 B] - a core matrix (8 rows x 16 B) is then 128
// contiguous bytes starting at ANY row - and chunk d is addressed by moving the UMMA shared-memory
// descriptor's start address up by d rows (16 B each).  (The 128-byte-swizzled layout was tried first:
// it is only usable when the start row is a multiple of 8, measured on the B200.)  The 8-row core
// groups of one MMA (M = 128 = 16 groups) come from 16 different streams: the descriptor's stride-byte-
// offset is the per-stream pitch.  W_d (a Toeplitz block of taps, the same for every stream) is laid
// out on the host and streamed through a 2-slot ring with cp.async.bulk + mbarrier.
//
// Precision: x = hi + lo with hi = x with the 13 low mantissa bits cleared (what a tf32 operand read
// keeps), lo = x - hi (exact); same for the taps; three products per step (hi*hi, hi*lo, lo*hi), lo*lo
// (2^-22 relative) dropped.  Products of 11-bit mantissas are exact in FP32; the accumulators are FP32.
// The W tile stacks [W_hi ; W_lo] along N, so x_hi*[W_hi | W_lo] is ONE MMA (N = 128) and x_lo*W_hi a
// second (N = 64): the operand fetch from shared memory is what bounds this kernel (measured:
// ~38 clk + bytes/128 per MMA), so every A tile should be read as few times as possible.  The three
// accumulators are added in the epilogue.
//
// One CTA per SM, 512 threads, one work item = 16 streams x one block of `frames` (multiple of 512)
// samples:  load: [history | block] -> hi/lo tiles (+ next history to HBM);  MMA: warps 1.. issue while
// warp 0 feeds W slices (one K = 8 slice = 4 KB per ring slot);  epilogue: warps 0-7 read the accumulators
// (tcgen05.ld), dither / quantise and store.  The sample tiles fill shared memory, so the next item cannot
// be loaded under the MMAs - but it can be loaded under the EPILOGUE: once the accumulators are final the
// tiles are free, and warps 8-15 fill them with the next item while warps 0-7 drain TMEM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "afp_fir_kernel.cuh"

namespace afp {

constexpr int TC_STREAMS = 16;             // streams per work item (16 core groups of 8 rows = M 128)
constexpr int TC_THREADS = 512;            // warps 0-7: tap ring, MMA issue, epilogue; warps 8-15: loaders of the NEXT item
constexpr int TC_MAIN_THREADS = 256;

// Row length R (samples per row = outputs per accumulator row = K per chunk).  R = 64 halves the number
// of MMAs (they are bound by operand fetch, ~64 B/clk measured, and by a per-instruction cost) but
// needs blocks of 512 frames; R = 32 also takes 256-frame blocks.
template <int R>
struct TcCfg {
    static constexpr int ROW = R;
    static constexpr int KSTEPS = R / 8;                    // K = 8 per tf32 MMA
    static constexpr int W_SLICE = 2 * 2 * R * 4;           // floats per K = 8 slice of taps: 2 column chunks x (R hi + R lo rows) x 4
    static constexpr int W_FLOATS = KSTEPS * W_SLICE;       // one chunk d of taps (8 KB at R = 32, 32 KB at R = 64)
    static constexpr int KPS = R == 32 ? 4 : 2;             // k-steps per ring slot (8 KB slots)
    static constexpr int UPC = KSTEPS / KPS;                // ring units per chunk
    static constexpr int SLOT_FLOATS = KPS * W_SLICE;
    static constexpr int W_LBO = 2 * R * 16;                // W slice: 2R rows (R hi, R lo) per column chunk
    static constexpr int W_SBO = 8 * 16;
};

struct TcParams {
    const float* hist_in;      // [n_streams][hist_rows*32]  carried samples (oldest first)
    const float* in;           // [n_streams][in_pitch]      the new block
    long long in_pitch;
    float* hist_out;           // [n_streams][hist_rows*32]
    void* out;                 // [n_streams][out_pitch] elements
    long long out_pitch;
    const float* w;            // [n_chunks][8 k-steps][2 col chunks][128 rows: 64 hi, 64 lo][4]  Toeplitz tap slices
    int n_streams;
    int frames;                // multiple of 512
    int hist_rows;             // rows of history (>= n_chunks - 1)
    int n_chunks;              // D + 1
    int n_items;               // ceil(n_streams / 16)
    unsigned idesc_override;   // timing experiments only (0 = the real instruction descriptor)
    int w_slots;               // depth of the W ring in shared memory (2, 4 or 8): hides the L2 latency of the tap tiles
    // epilogue (same meaning as FirParams)
    int fmt, dither, decim, fast_out;      // decim is always 1 here
    unsigned long long seed, sample_base;
    int stream_id0;
    float gain, dither_lsb;
};

typedef void (*tc_kern_t)(const TcParams);

// ---- small PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: wait in hardware, not in a spin
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// descriptors travel as (low word, high word): only the low word (start address) changes between MMAs,
// so each MMA costs one 32-bit uniform add per operand instead of 64-bit arithmetic
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp (the way CUTLASS issues tcgen05.mma / TMA from warp-uniform code)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): core
// matrices of 8 rows x 16 B; lbo = byte distance between core matrices adjacent in K, sbo = between
// 8-row groups adjacent in M / N
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address        bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;        // leading byte offset  bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;        // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version   bits [46,48) = 1 (sm_100)
    return d;                                                // base offset 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor: D f32, A/B tf32 K-major, M = 128 (cute::UMMA::InstrDescriptor layout)
__host__ __device__ constexpr uint32_t tc_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// byte offset of float (row r, column i) in a K-major core-matrix tile of `rows_padded` rows:
// [column chunk i/4][row][4 floats]
__host__ __device__ constexpr int tile_off(int r, int i, int rows_padded) { return (i >> 2) * rows_padded * 16 + r * 16 + ((i & 3) << 2); }

template <int EPI, int NTS, int R>   // R = row length, NTS = frames / (8 R): 8-row groups per stream and block
__global__ void __launch_bounds__(TC_THREADS, 1)
fir_tc_kernel(const __grid_constant__ TcParams P) {
    using Cfg = TcCfg<R>;
    constexpr int TC_ROW = R, TC_KSTEPS = Cfg::KSTEPS, TC_W_SLICE = Cfg::W_SLICE, TC_W_FLOATS = Cfg::W_FLOATS;
    constexpr int TC_KPS = Cfg::KPS, TC_UPC = Cfg::UPC, TC_SLOT_FLOATS = Cfg::SLOT_FLOATS, TC_W_LBO = Cfg::W_LBO, TC_W_SBO = Cfg::W_SBO;
    (void)TC_KSTEPS; (void)TC_W_FLOATS;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int block_rows = P.frames / TC_ROW;                 // rows of the new block per stream
    const int rows = P.hist_rows + block_rows;                // rows per stream tile
    const int rows_p = rows;                                  // (no row padding: the bytes go to a deeper W ring)
    const int a_lbo = rows_p * 16;                            // bytes between column chunks of one stream
    const int pitch = (TC_ROW / 4) * a_lbo;                   // bytes per stream tile
    constexpr int n_t = NTS;                                  // 8-row groups per stream = accumulator sets
    constexpr int N_ROLES = NTS == 1 ? 2 : 4;                 // MMA-issuing warps (warps 1..N_ROLES)

    // shared memory carve-up: data hi | data lo | W ring (2 x 8 KB) | barriers | tmem ptr
    unsigned char* data_hi = tc_smem;
    unsigned char* data_lo = data_hi + TC_STREAMS * pitch;
    unsigned char* wring = data_lo + TC_STREAMS * pitch;
    const int NS = P.w_slots;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + NS * TC_SLOT_FLOATS * 4);
    uint64_t* w_full = bars;          // [NS]
    uint64_t* w_empty = bars + 16;    // [NS]
    uint64_t* acc_full = bars + 32;   // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 33);

    const int tmem_cols_needed = n_t * 3 * TC_ROW;            // per set t: [hh | hl] 128 columns + [lh] 64 columns
    const uint32_t tmem_cols = tmem_cols_needed <= 128 ? 128 : (tmem_cols_needed <= 256 ? 256 : 512);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < NS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], N_ROLES); }
        mbar_init(acc_full, N_ROLES);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp-uniform copies: tcgen05.mma takes its operands from uniform registers; with per-thread values
    // ptxas wraps every MMA in an ELECT / R2UR waterfall loop (~100 clk per MMA, measured)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

    uint32_t w_prod = 0, w_cons = 0, acc_phase = 0;           // ring positions (chunks produced / consumed so far)
    const int hist_len = P.hist_rows * TC_ROW;
    const int total = hist_len + P.frames;                    // samples per stream tile
    const int n4 = total >> 2;

    // ---- load of one item by `lthreads` threads (ltid = index among them)
    auto load_item = [&](int item_l, int ltid, int lthreads) {
        const int s0 = item_l * TC_STREAMS;
        // ---- load: [history | block] -> hi / lo swizzled tiles; next block's history -> HBM
        // consecutive threads take consecutive ROWS of one column chunk: their 16-byte shared-memory
        // stores are contiguous (no bank conflicts in the core-matrix layout); the global loads are 16 B
        // per 128-byte row, the other chunks of the row follow from L1/L2
        constexpr int LB = 8;                                 // loads in flight per thread
        for (int base = ltid; base < TC_STREAMS * n4; base += lthreads * LB) {
            float4 v[LB];
#pragma unroll
            for (int u = 0; u < LB; ++u) {
                const int idx = base + u * lthreads;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < TC_STREAMS * n4) {
                    const int sl = idx / n4, g = idx - sl * n4;   // local stream, (chunk, row) index in its tile
                    const int c = g / rows, r = g - c * rows;
                    const int s = s0 + sl, pos = r * TC_ROW + c * 4;   // sample index in [history | block]
                    if (s < P.n_streams)
                        v[u] = pos < hist_len ? __ldg(reinterpret_cast<const float4*>(P.hist_in + (long long)s * hist_len + pos))
                                              : __ldg(reinterpret_cast<const float4*>(P.in + (long long)s * P.in_pitch + (pos - hist_len)));
                }
            }
#pragma unroll
            for (int u = 0; u < LB; ++u) {
                const int idx = base + u * lthreads;
                if (idx < TC_STREAMS * n4) {
                    const int sl = idx / n4, g = idx - sl * n4;
                    const int c = g / rows, r = g - c * rows;
                    const int s = s0 + sl, pos = r * TC_ROW + c * 4;
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u); lo.x = v[u].x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u); lo.y = v[u].y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u); lo.z = v[u].z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u); lo.w = v[u].w - hi.w;
                    const int off = sl * pitch + c * a_lbo + r * 16;
                    *reinterpret_cast<float4*>(data_hi + off) = hi;
                    *reinterpret_cast<float4*>(data_lo + off) = lo;
                    if (s < P.n_streams && pos >= P.frames)       // last hist_len samples of [history | block]
                        *reinterpret_cast<float4*>(P.hist_out + (long long)s * hist_len + (pos - P.frames)) = v[u];
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA unit
    };
    if (blockIdx.x < P.n_items) load_item(blockIdx.x, tid, TC_THREADS);   // first item: everybody loads
    __syncthreads();

    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int s0 = item * TC_STREAMS;
        // ---- phase 2: W producer (warp 0) and MMA issuer (warp 1); whole warps run the loops, one
        // elected lane issues, so every operand stays warp-uniform (uniform registers, no waterfall)
        if (warp_u == 0) {
            uint32_t slot = w_prod % NS, round = w_prod / NS;   // (one division per item; NS need not be a power of two)
            for (int u = 0; u < P.n_chunks * TC_UPC; ++u) {
                if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(&w_full[slot], TC_SLOT_FLOATS * 4);
                    bulk_g2s(wring + slot * TC_SLOT_FLOATS * 4, P.w + (size_t)u * TC_SLOT_FLOATS, TC_SLOT_FLOATS * 4, &w_full[slot]);
                }
                __syncwarp();
                if (++slot == (uint32_t)NS) { slot = 0; ++round; }
            }
        } else if (warp_u >= 1 && warp_u <= N_ROLES) {
            // MMA issuers.  The MMAs are small (128 x 64 x 8 and 128 x 32 x 8) and a single issuing warp
            // spends ~90 clk per MMA on uniform-datapath bookkeeping (ncu: tensor pipe 2.5 % active with one
            // issuer), so the accumulators are split over N_ROLES warps that issue independently: role =
            // (row-group set, product family).  Each role commits its own MMAs to the slot's "empty"
            // barrier and to acc_full (both initialised with N_ROLES arrivals).
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int role = warp_u - 1;
            // Descriptors are affine in (chunk, set, k-step) - the start-address field counts 16-byte
            // units and one row is one unit, so moving the A operand up a row is "low word - 1".
            const uint64_t a_hi_d = umma_desc(smem_u32(data_hi), a_lbo, pitch), a_lo_d = umma_desc(smem_u32(data_lo), a_lbo, pitch);
            const uint64_t w_d = umma_desc(smem_u32(wring), TC_W_LBO, TC_W_SBO);
            const uint32_t dhi_a = (uint32_t)(a_hi_d >> 32), dhi_w = (uint32_t)(w_d >> 32);   // same high word for hi / lo tiles
            const uint32_t a_k = (2u * a_lbo) >> 4;           // K = 8 = two column chunks
            const uint32_t id_wide = P.idesc_override ? P.idesc_override : tc_idesc(2 * TC_ROW), id_narrow = tc_idesc(TC_ROW);
            // what this role issues per k-step: sets [t_lo, t_hi), wide and / or narrow product
            const int t_lo = NTS == 1 ? 0 : (NTS == 2 ? role >> 1 : role), t_hi = t_lo + 1;
            const bool do_wide = NTS == 4 ? true : (role & 1) == 0, do_narrow = NTS == 4 ? true : (role & 1) == 1;
            uint32_t slot = w_cons % NS, phase = (w_cons / NS) & 1;
            for (int u = 0; u < P.n_chunks * TC_UPC; ++u) {
                const int d = u / TC_UPC, k0 = (u % TC_UPC) * TC_KPS;    // chunk and first k-step of this ring unit
                mbar_wait(&w_full[slot], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t ah = (uint32_t)a_hi_d + (uint32_t)P.hist_rows - (uint32_t)d + k0 * a_k;
                    const uint32_t al = (uint32_t)a_lo_d + (uint32_t)P.hist_rows - (uint32_t)d + k0 * a_k;
                    const uint32_t wl = (uint32_t)w_d + ((slot * (uint32_t)(TC_SLOT_FLOATS * 4)) >> 4);
                    const uint32_t fl = u == 0 ? 0u : 1u;      // first write of the accumulators
                    for (int t = t_lo; t < t_hi; ++t) {
                        const uint32_t acc = tmem_base + (uint32_t)(t * 3 * TC_ROW);
#pragma unroll
                        for (int k = 0; k < TC_KPS; ++k) {
                            const uint32_t wk = wl + (uint32_t)((k * TC_W_SLICE * 4) >> 4);
                            const uint32_t off = (uint32_t)(8 * t) + k * a_k;
                            const uint32_t f = k == 0 ? fl : 1u;
                            if (do_wide) umma_tf32(acc, ah + off, dhi_a, wk, dhi_w, id_wide, f);                    // x_hi * [W_hi | W_lo]
                            if (do_narrow) umma_tf32(acc + 2 * TC_ROW, al + off, dhi_a, wk, dhi_w, id_narrow, f);   // x_lo * W_hi
                        }
                    }
                    umma_commit(&w_empty[slot]);              // this role's MMAs on the slot are done
                }
                __syncwarp();
                if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full);           // this role's accumulators are final
            __syncwarp();
        } else if (warp_u == N_ROLES + 1) {
            // a warp that idles through the MMA phase: pull the next item's rows into L2 meanwhile
            const int nitem = item + gridDim.x;
            if (nitem < P.n_items && lane < TC_STREAMS) {
                const int s = nitem * TC_STREAMS + lane;
                if (s < P.n_streams) {
                    prefetch_row_l2(P.hist_in + (long long)s * hist_len, 0, hist_len);
                    prefetch_row_l2(P.in + (long long)s * P.in_pitch, 0, P.frames);
                }
            }
        }
        w_cons += P.n_chunks * TC_UPC;                        // ring position after this item (same in every thread)
        w_prod = w_cons;
        // ---- phase 3: accumulators -> registers -> dither / quantise -> HBM
        mbar_wait(acc_full, acc_phase & 1);
        acc_phase++;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncwarp();
        // warp w reads TMEM lanes 32*(w%4)..+31 (accumulator rows); warps 0-3 take the low 32 outputs of
        // each 64-output row, warps 4-7 the high 32
        constexpr int H = TC_ROW / 32;                        // 32-output pieces per accumulator row
        if (warp_u >= TC_MAIN_THREADS / 32) {
            // the accumulators are final, so the MMAs have read the sample tiles for the last time: warps 8-15
            // refill them with the next item while warps 0-7 drain TMEM
            const int nitem = item + gridDim.x;
            if (nitem < P.n_items) load_item(nitem, tid - TC_MAIN_THREADS, TC_THREADS - TC_MAIN_THREADS);
        } else
        for (int u = warp >> 2; u < n_t * H; u += 2) {        // warps 0-3 and 4-7 alternate over (set, piece)
            const int t = u / H, half = u % H;
            const int m = 32 * (warp & 3) + lane;             // accumulator row = TMEM lane
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * 3 * TC_ROW + 32 * half);
            float y[32];
#pragma unroll
            for (int f = 0; f < 3; ++f) {                     // hh + hl + lh: column blocks 0, 64, 128
                uint32_t a[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                      "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
                      "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),
                      "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
                    : "r"(taddr + f * TC_ROW));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = f == 0 ? __uint_as_float(a[j]) : __fadd_rn(y[j], __uint_as_float(a[j]));
            }
            const int s = s0 + (m >> 3);
            const long long o_base = (long long)(8 * t + (m & 7)) * TC_ROW + 32 * half;
            store_row<32, EPI>(P, y, s, o_base, s < P.n_streams ? 32 : 0);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                      // tiles and accumulators free for the next item
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

}  // namespace afp
