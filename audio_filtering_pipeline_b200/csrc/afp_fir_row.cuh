// Single-row streaming FIR / polyphase kernel for sm_100a ("kernel C"): the short- and mid-length
// filter path (HBM-bound or co-limited cases of the north star: <= 101 taps, 1x-4x up-sampling).
//
// Same contract as fir_stream_kernel (afp_fir_kernel.cuh; reference stream_process.py:45-46,97-107):
//       X = [ src0 | src1 | zeros ],  y[(q-q_begin)*U + p] = sum_j h[j*U+p] * X[q-j]
// but a different mapping to the SM, built around one measured fact: FFMA2 also exists in the form
//       FFMA2 Racc.F32x2, Rx.F32 (scalar broadcast, ORDINARY register), URh.F32x2 (tap PAIR, uniform)
// which runs at 121-125 of the 128 FMA/clk/SM (bench/microbench/ffma_peak.cu, variant x2_bc).  The two
// lanes of a packed FMA are then two CONSECUTIVE OUTPUT SAMPLES of one row that share the input sample:
//   * U even: lanes = phases (p, p+1) of one base position, taps (h[jU+p], h[jU+p+1]) - an aligned pair
//     of the natural tap table;
//   * U = 1:  lanes = positions (r, r+1), taps (h[k], h[k+1]); k takes both parities, so the launch
//     carries a second copy of the table shifted by one sample (pairs starting at odd k);
//   * U odd > 1 (U = 3): every phase p is a U = 1 problem with the sub-filter h_p[j] = h[jU+p]: lanes =
//     positions (r, r+1) of one phase, one table pair per phase; the epilogue interleaves the phases.
// Consequences, compared with the row-pair kernel:
//   * no row interleave: shared memory holds the plain, contiguous samples of ONE row segment, so the
//     loader is one or two bulk copies (cp.async.bulk global -> shared, completion on an mbarrier) issued
//     by a single lane, double buffered: the next item's samples are in flight while the current one
//     computes, and the next block's history leaves the same way (one bulk copy shared -> global).  The
//     first version used per-thread cp.async into a padded, bank-conflict-free layout: ncu showed ~30 %
//     of all executed instructions in its address arithmetic.  The unpadded layout costs 2- to 4-way
//     conflicts on the window LDS.128, which is < 35 % of the LSU wavefront budget at the measured rates;
//   * the register window is scalars, not pairs: half the registers per window sample;
//   * accumulator pairs ARE consecutive output samples: no shuffling between FMA and store format;
//   * work items are single segments (finer grain, no odd-row padding);
//   * the last chunk of the tap loop is specialised on how many taps it really holds (multiples of 4 /
//     2), so a 51-tap filter costs 52 tap slots per output, not 64.
// Stores: a thread's run of T*U consecutive outputs, written directly, puts 16 bytes into 32 different
// 128-byte lines per instruction - measured at 2.3-3.5 TB/s against 6.2 TB/s when 4 or more lanes write
// contiguous bytes (bench/microbench/store_pattern2.cu), and the phase-skip experiment showed exactly
// that store rate bounding the U = 4 short-filter case.  STORE = 2 therefore transposes the warp's
// outputs through a warp-private shared-memory tile: every STG.128 covers 512 contiguous bytes.
#pragma once
#include "afp_fir_kernel.cuh"

// phase-skip experiment hooks (bench/microbench/fir_row_test.cu -DAFP_ROW_DBG): mask in P.n_sub, which
// then does not carry the tail size.  bit0 skip the FMA phase, bit1 skip stores, bit2 skip global loads
#ifdef AFP_ROW_DBG
#define AFP_ROW_DBG_LOADS(x) ((x) && !(P.n_sub & 4))
#define AFP_ROW_DBG_NOFMA (P.n_sub & 1)
#define AFP_ROW_DBG_NOSTORE (P.n_sub & 2)
#else
#define AFP_ROW_DBG_LOADS(x) (x)
#define AFP_ROW_DBG_NOFMA false
#define AFP_ROW_DBG_NOSTORE false
#endif

namespace afp {

typedef unsigned long long u64;

// volatile: keeps the FMAs in source order, which is written so that each tap pair is loaded once
// (LDCU.64) and used by T or T/2 consecutive, mutually independent FMAs.  Left to itself ptxas
// interleaves the tap loop differently and reloads every pair about twice (measured: one LDCU.64 per 4
// FFMA2 instead of one per 8).
__device__ __forceinline__ void ffma2_bc(u64& acc, float x, u64 hpair) {
    u64 xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
#ifdef AFP_ROW_NOVOL
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(hpair));
#else
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(hpair));
#endif
}

// Shared-memory layout of a region; a block of T samples is T/4 granules of 16 B.
//   PAD = 0: the plain samples, contiguous - what a bulk copy writes.  The per-thread-contiguous LDS.128
//            window pattern (lane stride = one block) then has T/4-way bank conflicts.
//   PAD = 1: every block is followed by one pad granule when that makes the block stride odd: conflict
//            free, filled by per-thread cp.async.  Used where the FMA pipe is the bound (U = 1, T = 16:
//            measured 440 vs 360 Gsamples/s at 51 taps), because the window loads sit on its critical path.
template <int T, int PAD> struct RowLayout {
    static constexpr int GPB = T / 4;
    static constexpr int BSTRIDE = GPB + ((PAD && GPB % 2 == 0) ? 1 : 0);
    __host__ __device__ static constexpr int granules(int samples) { return (samples / T) * BSTRIDE; }
    __device__ static __forceinline__ int phys(int g) { return BSTRIDE == GPB ? g : g + g / GPB; }
};

template <int T, int PAD>
__device__ __forceinline__ void load_wblock(float (&w)[T], const float4* __restrict__ region, int block) {
    const float4* p = region + block * RowLayout<T, PAD>::BSTRIDE;
#pragma unroll
    for (int m = 0; m < T / 4; ++m) {
        const float4 v = p[m];
        w[4 * m] = v.x; w[4 * m + 1] = v.y; w[4 * m + 2] = v.z; w[4 * m + 3] = v.w;
    }
}

// One chunk of up to T taps per phase.  hi = the T samples at the thread's positions moved down by c*T,
// lo = the T samples below them.  Window index m in [0, 2T): W[m] = m < T ? lo[m] : hi[m-T].
// J = taps of this chunk that are really there (J = T: full chunk).
//
// U even: acc[(r*U + p)/2] (+)= (h[(cT+j)U+p], h[(cT+j)U+p+1]) * W[T + r - j].
template <int U, int T, int J>
__device__ __forceinline__ void chunk_fma_even(u64 (&acc)[U * T / 2], const float (&lo)[T], const float (&hi)[T],
                                               const float* __restrict__ tp) {
#pragma unroll
    if (U % 2 != 0) return;   // (odd-U kernels instantiate this in dead code)
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int p2 = 0; p2 < U / 2; ++p2) {
            const u64 h = *reinterpret_cast<const u64*>(tp + j * U + 2 * p2);
#pragma unroll
            for (int r = 0; r < T; ++r) ffma2_bc(acc[(r * U) / 2 + p2], r >= j ? hi[r - j] : lo[T + r - j], h);
        }
    }
}

// U = 1: acc[r/2] = (y[r], y[r+1]).  Tap pair (h[k], h[k+1]), k = cT + off, meets sample W[r + T - off]
// for every even r: the even lane's chunk sits one tap lower (off = -1 .. J-2), which over all chunks
// covers k = -1 .. K'-2 on the even lane (h[-1] = 0) and 0 .. K'-1 on the odd lane, K' = padded length:
// K' must be >= taps + 1.  ta = natural table + cT (pairs at even k), tb = shifted table + cT with
// tb[i] = h[i-1] (pairs at odd k live at the even index k+1).
template <int T, int J>
__device__ __forceinline__ void chunk_fma_u1(u64 (&acc)[T / 2], const float (&lo)[T], const float (&hi)[T],
                                             const float* __restrict__ ta, const float* __restrict__ tb) {
#pragma unroll
    for (int off = -1; off < J - 1; ++off) {
        const u64 h = *reinterpret_cast<const u64*>((off & 1) ? tb + off + 1 : ta + off);
#pragma unroll
        for (int r = 0; r < T; r += 2) {
            const int m = r + T - off;
            ffma2_bc(acc[r / 2], m < T ? lo[m] : hi[m - T], h);
        }
    }
}

// one lane of a converged warp, chosen in warp-uniform control flow: the bulk-copy instructions take
// uniform-register operands, and issuing them under a per-thread condition (lt == 0) makes ptxas wrap
// them in R2UR code and give up the uniform datapath for the tap loads of the whole item loop
__device__ __forceinline__ bool row_elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ unsigned row_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ROW_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ROW_DONE;\n"
        "bra ROW_WAIT;\n"
        "ROW_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy, bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- store sinks (see store_row in afp_fir_kernel.cuh) ------------------------------------------------
// Per-thread bulk copy (cp.async.bulk shared -> global) of the thread's run: kept as a measured
// alternative (5.8-6.4 TB/s for 128-byte runs in isolation, but slower than the transposed stores inside
// the kernel and poor for 64-byte runs).
struct BulkSink {
    static constexpr bool kPhiloxSideBySide = false;
    static constexpr bool kAllFull = false;
    float4* stage;     // this thread's staging row (16-byte aligned)
    template <int NB>
    __device__ __forceinline__ void put(void* dst, const uint32_t (&w)[NB]) const {
        if (NB % 4 != 0) { store_words<NB>(dst, w); return; }   // bulk copies move whole 16-byte granules
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous copy finished reading
#pragma unroll
        for (int i = 0; i < NB / 4; ++i)
            reinterpret_cast<uint4*>(stage)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned sa = (unsigned)__cvta_generic_to_shared(stage);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(dst), "r"(sa), "r"(NB * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
};

// Warp transpose through shared memory.  All 32 lanes call put() together (the kernel checks that with a
// vote); lane l owns the NB words that go to dst(l) = dst(0) + l*NB*4.  Lanes write their run into row l
// of the warp's tile (row stride odd in granules: conflict-free STS.128), then instruction i of the
// read-back has lane l fetch granule i*32+l of the tile in global order and store it: one STG.128 covers
// 512 contiguous bytes.
struct TransposeSink {
    static constexpr bool kPhiloxSideBySide = false;
    static constexpr bool kAllFull = true;   // keeps the warp barriers below in uniform control flow
    float4* tile;      // the warp's tile: 32 rows of stage_stride_g(run bytes) granules
    template <int NB>
    __device__ __forceinline__ void put(void* dst, const uint32_t (&w)[NB]) const {
        if (NB % 4 != 0) { store_words<NB>(dst, w); return; }
        constexpr int GPT = NB / 4;                             // granules per thread
        constexpr int STR = GPT + ((GPT % 2 == 0) ? 1 : 0);     // row stride, odd
        const int lane = threadIdx.x & 31;
        uint4* t = reinterpret_cast<uint4*>(tile);
        __syncwarp();                                           // previous item's read-back is done
#pragma unroll
        for (int i = 0; i < GPT; ++i) t[lane * STR + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        __syncwarp();
        const u64 base = __shfl_sync(0xffffffffu, reinterpret_cast<u64>(dst), 0);
        uint4* g = reinterpret_cast<uint4*>(base);
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int n = i * 32 + lane;
            g[n] = t[(n / GPT) * STR + (n % GPT)];
        }
    }
};
// staging row stride in granules: the run plus one pad granule when that makes the stride odd
__host__ __device__ constexpr int stage_stride_g(int run_bytes) {
    return run_bytes / 16 + (((run_bytes / 16) % 2 == 0) ? 1 : 0);
}

// dynamic shared memory of a launch: mbarriers, two region buffers per thread group, store staging
constexpr int kRowBarGranules = 8;
__host__ __device__ constexpr int row_smem_bytes(int U, int T, int NT, int tg_log2, int kq, int store, int pad) {
    const int gpb = T / 4, bstride = gpb + ((pad && gpb % 2 == 0) ? 1 : 0);
    return (kRowBarGranules + (NT >> tg_log2) * 2 * (((kq + (T << tg_log2)) / T) * bstride)
            + (store ? NT * stage_stride_g(U * T * 4) : 0)) * 16;
}

// resident CTAs per SM the register budget is compiled for
__host__ __device__ constexpr int row_min_ctas(int U, int T, int NT) {
    return (NT == 256 ? 1 : 2) * (U * T >= 24 ? 3 : 4);
}

// Tap-table layout in FirParams::taps for this kernel: natural table (U * kq floats), then - U = 1 only -
// the shifted copy at offset kq + 4 (kq + 4 floats: [0, h[0], h[1], ...]).
__host__ __device__ constexpr int row_tabb_offset(int kq) { return kq + 4; }
// FirParams::n_sub and sub_stride are free in this kernel's launches (single table only): sub_stride
// carries the number of chunks that hold taps (the history / halo, n_chunks * T, may be longer) and n_sub
// the number of tap slots in the last of them (0 = it is full).  Granularity of the specialised tails:
__host__ __device__ constexpr int row_tail_step(int U, int T) { return U % 2 == 1 ? 4 : 2; }
// odd U: phase p's natural table sits at p * row_phase_stride(kq), its shifted copy row_tabb_offset(kq) further
__host__ __device__ constexpr int row_phase_stride(int kq) { return 2 * (kq + 4); }

// TGL: log2(threads per group) fixed at compile time (0 = read it from the parameters); STREAM: the launch is
// the streaming form (q_begin == kq == len0, so no copy range needs clamping at the front and nothing lies
// before the sequence) - both turn per-item uniform-datapath arithmetic into immediates.
template <int U, int T, int NT, int EPI, int STORE, int PAD, int TGL = 0, bool STREAM = false>
__global__ void __launch_bounds__(NT, row_min_ctas(U, T, NT))
fir_row_kernel(const __grid_constant__ FirParams P) {
    static_assert(U >= 1 && U <= 4, "row kernel: U in 1..4");
    using LY = RowLayout<T, PAD>;
    extern __shared__ float4 smem[];
    const int tid = threadIdx.x;
    const int tgl = TGL ? TGL : P.tg_log2;
    const int tg = 1 << tgl;                       // threads per segment (>= 32)
    const int slots = NT >> tgl;                   // thread groups per CTA
    const int lq = tg * T;                         // base positions per segment
    const int kq = P.n_chunks * T;                 // padded taps per phase = halo
    const int region_s = kq + lq;
    const int region_g = LY::granules(region_s);
    // warp-uniform by construction (a group is one or more whole warps)
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int slot = warp_u >> (tgl - 5);
    const bool lead_warp = ((warp_u << 5) & (tg - 1)) == 0;   // first warp of its group: issues the copies
    const int lt = tid & (tg - 1);
    // shared memory: [2 mbarriers per thread group | region buffers | store staging]
    constexpr int BAR_G = kRowBarGranules;         // granules reserved for mbarriers (16 x 8 bytes)
    const unsigned bar0 = row_smem_u32(smem) + slot * 16;                  // this group's two mbarriers
    float4* const buf0 = smem + BAR_G + slot * 2 * region_g;
    // store staging (sized for 4-byte output elements) follows the region buffers: one row per thread
    float4* const stage = smem + BAR_G + slots * 2 * region_g + (STORE == 2 ? (tid & ~31) : tid) * stage_stride_g(U * T * 4);
    const int b0 = P.n_chunks + lt;                // block holding this thread's own T positions
    const int stride = gridDim.x * slots;
    const int n4 = region_s >> 2;
    const int L = P.len0 + P.len1;
    const int n_items = P.n_seg_total;
#ifdef AFP_ROW_DBG
    const int tail = 0, n_full = AFP_ROW_DBG_NOFMA ? 0 : P.n_chunks;
#else
    const int tail = P.n_sub;                      // tap slots in the last chunk, 0 = full
    const int n_full = tail ? P.sub_stride - 1 : P.sub_stride;   // sub_stride = chunks that hold taps (<= n_chunks)
#endif

    auto decode = [&](int item, int& s, int& g) {
        if (P.segs_shift >= 0) { s = (int)((unsigned)item >> P.segs_shift); g = item & ((1 << P.segs_shift) - 1); }
        else { s = (int)((unsigned)item / (unsigned)P.segs_per_stream); g = item - s * P.segs_per_stream; }
    };
    // One lane starts the copy of an item's region - logical samples [start, start + region_s) of stream
    // s: granules [i_lo, i_sp) come from src0, [i_sp, i_hi) from src1; what lies outside the sequence
    // (never in the streaming form) is zeroed by the whole group when the item is consumed.
    auto issue = [&](int item, float4* dst, unsigned bar) {
        int s, g;
        decode(item, s, g);
        const int start = P.q_begin + g * lq - kq;
        const float* r0 = P.src0 + ((long long)s * P.pitch0 + start);
        const float* r1 = P.src1 + ((long long)s * P.pitch1 - P.len0 + start);
        const int i_lo = STREAM ? 0 : min(max((-start) >> 2, 0), n4);
        const int i_hi = STREAM ? min((L - start) >> 2, n4) : min(max((L - start) >> 2, i_lo), n4);
        const int i_sp = STREAM ? max((P.len0 - start) >> 2, 0) : min(max((P.len0 - start) >> 2, i_lo), i_hi);
        if (PAD == 0) {          // one elected lane: one bulk copy per source
            const unsigned n0 = AFP_ROW_DBG_LOADS(true) ? (unsigned)(i_sp - i_lo) * 16u : 0u;
            const unsigned n1 = AFP_ROW_DBG_LOADS(true) ? (unsigned)(i_hi - i_sp) * 16u : 0u;
            mbar_arrive_expect_tx(bar, n0 + n1);
            if (n0) bulk_g2s(row_smem_u32(dst + i_lo), r0 + 4 * i_lo, n0, bar);
            if (n1) bulk_g2s(row_smem_u32(dst + i_sp), r1 + 4 * i_sp, n1, bar);
        } else {                 // whole group: 16-byte cp.async into the padded layout, addresses by increment
            const unsigned sstep = (unsigned)(tg + tg / LY::GPB) * 16u;
            const unsigned sb = row_smem_u32(dst);
            {
                int i4 = i_lo + lt;
                const char* gp = reinterpret_cast<const char*>(r0) + 16ll * i4;
                unsigned sa = sb + 16u * (unsigned)LY::phys(i4);
                for (; AFP_ROW_DBG_LOADS(i4 < i_sp); i4 += tg, gp += 16 * tg, sa += sstep) cp_async16(sa, gp);
            }
            {
                int i4 = i_sp + lt;
                const char* gp = reinterpret_cast<const char*>(r1) + 16ll * i4;
                unsigned sa = sb + 16u * (unsigned)LY::phys(i4);
                for (; AFP_ROW_DBG_LOADS(i4 < i_hi); i4 += tg, gp += 16 * tg, sa += sstep) cp_async16(sa, gp);
            }
            cp_async_commit();
        }
    };

    if (PAD == 0) {
        if (lead_warp && row_elect_one()) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        group_sync<NT>(tg, slot);
    }

    int item = slot * gridDim.x + blockIdx.x;
    if (PAD == 0) { if (lead_warp && item < n_items && row_elect_one()) issue(item, buf0, bar0); }
    else if (item < n_items) issue(item, buf0, bar0);
    for (int it_n = 0; item < n_items; item += stride, ++it_n) {
        const int cur = it_n & 1;
        float4* const region_w = buf0 + cur * region_g;
        const float4* const region = region_w;
        if (PAD == 0) mbar_wait(bar0 + cur * 8, (it_n >> 1) & 1);
        else cp_async_wait_all();
        group_sync<NT>(tg, slot);      // everyone has left the other buffer (and sees this one)

        int stream, seg;
        decode(item, stream, seg);
        const int start = P.q_begin + seg * lq - kq;

        if (PAD == 0) {
            if (lead_warp && row_elect_one()) {
                // the other buffer's history copy (previous item) must have been read before it is refilled
                bulk_wait_read_all();
                if (item + stride < n_items) issue(item + stride, buf0 + (cur ^ 1) * region_g, bar0 + (cur ^ 1) * 8);
                // ---- carried history for the next block, straight from the staged samples
                if (P.hist_out) {
                    const int hlo = max(P.hist_from, seg == 0 ? INT_MIN : start + kq), hhi = P.hist_from + P.hist_len;
                    const int h0 = max((hlo - start) >> 2, 0), h1 = min((hhi - start) >> 2, n4);
                    if (h0 < h1)
                        bulk_s2g(P.hist_out + ((long long)stream * P.hist_pitch - P.hist_from + start + 4 * h0),
                                 row_smem_u32(region + h0), (unsigned)(h1 - h0) * 16u);
                }
            }
        } else {
            if (item + stride < n_items) issue(item + stride, buf0 + (cur ^ 1) * region_g, 0u);
            if (P.hist_out) {
                const int hlo = max(P.hist_from, seg == 0 ? INT_MIN : start + kq), hhi = P.hist_from + P.hist_len;
                const int h0 = max((hlo - start) >> 2, 0), h1 = min((hhi - start) >> 2, n4);
                if (h0 < h1) {
                    float* hrow = P.hist_out + ((long long)stream * P.hist_pitch - P.hist_from + start);
                    for (int i4 = h0 + lt; i4 < h1; i4 += tg)
                        *reinterpret_cast<float4*>(hrow + 4 * i4) = region[LY::phys(i4)];
                }
            }
        }
        if ((!STREAM && start < 0) || start + region_s > L) {   // group-uniform; 'same' modes, ragged last segment
            const int i_lo = min(max((-start) >> 2, 0), n4), i_hi = min(max((L - start) >> 2, i_lo), n4);
            for (int i4 = lt; i4 < n4; i4 += tg)
                if (i4 < i_lo || i4 >= i_hi) region_w[LY::phys(i4)] = make_float4(0.f, 0.f, 0.f, 0.f);
            group_sync<NT>(tg, slot);
        }

        // ---- compute
        u64 acc[U * T / 2];
#pragma unroll
        for (int i = 0; i < U * T / 2; ++i) acc[i] = 0ull;
        float wa[T], wb[T];
        load_wblock<T, PAD>(wb, region, b0);
        int c = 0;
        if (U % 2 == 1) {
            // phase-major accumulators: acc[p*(T/2) + r/2] = (y_p[r], y_p[r+1])
            u64(&ap)[U][T / 2] = reinterpret_cast<u64(&)[U][T / 2]>(acc);
            const float* ta = P.taps;
            const float* tb = P.taps + row_tabb_offset(kq);
            const int ps = row_phase_stride(kq);
            for (; c + 2 <= n_full; c += 2) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
#pragma unroll
                for (int p = 0; p < U; ++p) chunk_fma_u1<T, T>(ap[p], wa, wb, ta + p * ps + c * T, tb + p * ps + c * T);
                load_wblock<T, PAD>(wb, region, b0 - c - 2);
#pragma unroll
                for (int p = 0; p < U; ++p) chunk_fma_u1<T, T>(ap[p], wb, wa, ta + p * ps + (c + 1) * T, tb + p * ps + (c + 1) * T);
            }
            if (c < n_full) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
#pragma unroll
                for (int p = 0; p < U; ++p) chunk_fma_u1<T, T>(ap[p], wa, wb, ta + p * ps + c * T, tb + p * ps + c * T);
#pragma unroll
                for (int i = 0; i < T; ++i) wb[i] = wa[i];
                ++c;
            }
            if (tail) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
                ta += c * T; tb += c * T;
#pragma unroll
                for (int p = 0; p < U; ++p) {
                    if (T >= 8 && tail == 4) chunk_fma_u1<T, 4>(ap[p], wa, wb, ta + p * ps, tb + p * ps);
                    else if (T >= 16 && tail == 8) chunk_fma_u1<T, (T >= 16 ? 8 : T)>(ap[p], wa, wb, ta + p * ps, tb + p * ps);
                    else if (T >= 16 && tail == 12) chunk_fma_u1<T, (T >= 16 ? 12 : T)>(ap[p], wa, wb, ta + p * ps, tb + p * ps);
                    else chunk_fma_u1<T, T>(ap[p], wa, wb, ta + p * ps, tb + p * ps);
                }
            }
        } else {
            const float* tp = P.taps;
            for (; c + 2 <= n_full; c += 2) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
                chunk_fma_even<U, T, T>(acc, wa, wb, tp + c * (T * U));
                load_wblock<T, PAD>(wb, region, b0 - c - 2);
                chunk_fma_even<U, T, T>(acc, wb, wa, tp + (c + 1) * (T * U));
            }
            if (c < n_full) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
                chunk_fma_even<U, T, T>(acc, wa, wb, tp + c * (T * U));
#pragma unroll
                for (int i = 0; i < T; ++i) wb[i] = wa[i];
                ++c;
            }
            if (tail) {
                load_wblock<T, PAD>(wa, region, b0 - c - 1);
                tp += c * (T * U);
                if (tail == 2) chunk_fma_even<U, T, 2>(acc, wa, wb, tp);
                else if (T >= 8 && tail == 4) chunk_fma_even<U, T, (T >= 8 ? 4 : T)>(acc, wa, wb, tp);
                else if (T >= 8 && tail == 6) chunk_fma_even<U, T, (T >= 8 ? 6 : T)>(acc, wa, wb, tp);
                else chunk_fma_even<U, T, T>(acc, wa, wb, tp);
            }
        }

        // ---- epilogue: accumulator pairs are consecutive output samples
        if (AFP_ROW_DBG_NOSTORE) continue;
        const int q_local = seg * lq + lt * T;
        const int nv = min(max(P.n_q - q_local, 0), T) * U;
        float y[U * T];
        if (U % 2 == 1 && U > 1) {     // phase-major pairs -> output order y[r*U + p]
#pragma unroll
            for (int p = 0; p < U; ++p)
#pragma unroll
                for (int r = 0; r < T; r += 2) {
                    const u64 a = acc[p * (T / 2) + r / 2];
                    y[r * U + p] = __uint_as_float((unsigned)a);
                    y[(r + 1) * U + p] = __uint_as_float((unsigned)(a >> 32));
                }
        } else {
#pragma unroll
            for (int i = 0; i < U * T / 2; ++i) {
                y[2 * i] = __uint_as_float((unsigned)acc[i]);
                y[2 * i + 1] = __uint_as_float((unsigned)(acc[i] >> 32));
            }
        }
        const long long o_base = (long long)q_local * U;
        if (STORE == 2) {
            // whole warp full and aligned -> transposed stores; otherwise (ragged tail) direct
            const bool decim1 = EPI >= 0 || P.decim == 1;
            if (__all_sync(0xffffffffu, nv == U * T) && P.fast_out && decim1)
                store_row<U * T, EPI>(P, y, stream, o_base, nv, TransposeSink{stage});
            else
                store_row<U * T, EPI>(P, y, stream, o_base, nv);
        } else if (STORE == 1) {
            store_row<U * T, EPI>(P, y, stream, o_base, nv, BulkSink{stage});
        } else {
            store_row<U * T, EPI>(P, y, stream, o_base, nv);
        }
    }
    bulk_wait_read_all();   // shared memory must outlive the bulk copies that read it
}

}  // namespace afp
