// Automatic gain control of one block per stream: the device form of the reference's apply_agc
// (stream_process_AGC.py:43-89) and of the clip at its call site (:143).  Per block and stateless:
//
//   rms[i]      = sqrt( mean of x^2 over the window centred on i, zero-extended: np.convolve 'same' )   :60-62
//   desired[i]  = clip(target / (rms[i] + 1e-10), 0, max_gain)                                          :65
//   smoothed[0] = desired[0];  smoothed[i] = a*desired[i] + (1-a)*smoothed[i-1],
//                 a = a_attack if desired[i] > smoothed[i-1] else a_release                             :68-80
//   y[i]        = clip(x[i] * clip(smoothed[i], 0.1, max_gain), -0.99, 0.99)                            :83-86, :143
//
// The smoother is a data-dependent recurrence, f_i(s) = max(L_att(s), L_rel(s)) for a_attack >= a_release
// (min otherwise): compositions of such maps are piecewise linear with a growing number of pieces, so
// there is no fixed-size associative scan - every stream is one sequential chain of `frames` steps of
// FFMA2 -> FMNMX (13.5 clk per step, bench/microbench/agc_chain.cu).  Parallelism comes from the batch: one
// LANE per stream in that phase, and the HBM roof (2.8 samples per clock and SM) needs ~38 such lanes busy
// per SM all the time - so what matters is how many streams an SM holds at once.
//
// Time-chunked rows.  A work item is `group` streams x the whole block, walked in chunks of 256 samples;
// shared memory holds ONE 256-float row per stream (1 KB: the first version kept a stream's whole block
// plus a float64 prefix-sum row, 12.7 KB per 1024-frame stream, and an SM held 38 streams; now it holds
// ~190).  Per chunk:
//   phase A  warp per stream, 8 consecutive samples per lane.  The window sum slides:
//            S(i+1) = S(i) + x^2[i+R+1] - x^2[i-L], so the chunk's sums are  carry + exclusive scan of
//            d[m] = x^2[lead m] - x^2[lag m]  (x^2 in float32 as numpy computes it, the differences and
//            the scan in float64: the terms that leave the window cancel the ones that entered it to
//            1e-16, a quiet window after a loud passage keeps its digits); carry = S(chunk start), one
//            float64 per stream.  No prefix row exists.  x is read three times (as lead, as lag W samples
//            later, for the multiply): the first read is the HBM read, the others L1 / L2 hits.
//            desired gain -> the stream's row.  The same pass applies the PREVIOUS chunk (its smoothed
//            gains are still in the row): multiply, clip, 16-byte stores - so a chunk costs two CTA
//            barriers, and the loads of both halves are in flight together.
//   phase B  lane per stream (warp 0): the recurrence over the chunk in place, smoother state carried in
//            a register; both branches in packed FFMA2s (s * (1-a_att, 1-a_rel) + d * (a_att, a_rel)),
//            then FMNMX; four steps per LDS.128 / STS.128 (row pitch 65 x 16 B: lane = stream accesses
//            are conflict-free), the next 16 steps loaded before the current 16 are computed.  The other
//            warps pull the next chunk's lead samples into L2 meanwhile.
// 128-thread CTAs, up to 6 per SM: one CTA's serial phase B overlaps the others' memory phases.
//
// The reference's error path is kept: int(attack*window) == 0 (or release) makes its loop divide by
// zero the first time that branch is taken; it catches the exception and returns the block unchanged
// (:87-89).  CHECKED instantiation: the recurrence is evaluated with the explicit comparison, a stream
// that takes an invalid branch is flagged and its whole row is rewritten with gain 1 when the item ends.
// A window longer than the block fails in the reference when the gains are applied (shape mismatch):
// whole launch passes through.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace afp {

constexpr int AGC_CHUNK = 256;                 // samples of a stream per pass: 8 per lane
constexpr int AGC_PITCH = AGC_CHUNK + 4;       // floats per row: 65 x 16 B, odd
constexpr int AGC_THREADS = 128;
#ifndef AGC_MIN_CTAS
#define AGC_MIN_CTAS 6
#endif
constexpr int AGC_MAX_GROUP = 32;              // streams per work item (one lane each in phase B)

struct AgcParams {
    const float* in;           // [n_streams][in_pitch]
    long long in_pitch;
    float* out;                // [n_streams][out_pitch]   (must not overlap `in`)
    long long out_pitch;
    int n_streams, frames;
    int window, left, right;   // window length; samples before / after the centre: left + 1 + right = window
    int group;                 // streams per work item, <= AGC_MAX_GROUP
    int n_items;
    float inv_window;
    float fast_scale;          // target * sqrt(window): desired = min(fast_scale * rsqrt(window sum), max_gain)
    float target, max_gain, min_gain, clip;   // clip <= 0: no output clip
    float a_att, c_att, a_rel, c_rel;          // alpha and 1 - alpha of the two branches
    int att_ok, rel_ok;        // 0: that branch raises in the reference
    int passthrough;           // window > frames
    int probe;                 // timing probes (AFP_AGC_PROBE, results are wrong): 2 = skip the recurrence, 4 = skip the gain
                               // arithmetic, 8 = skip the apply; 1 selects the float32-sum instantiation on the host
};

typedef void (*agc_kern_t)(const AgcParams);

// kernel variants
constexpr int AGC_FAST = 0;      // both constants valid, epsilon-free gain
constexpr int AGC_EXACT = 1;     // both constants valid, target / (rms + 1e-10)
constexpr int AGC_CHECKED = 2;   // a zero-sample constant: explicit comparison, failing streams pass through
constexpr int AGC_PROBE_F32 = 3; // TIMING PROBE ONLY (AFP_AGC_PROBE=1): window sums in float32 - shows what the float64 arithmetic costs

__host__ __device__ constexpr size_t agc_smem_bytes(int group) {
    return (size_t)group * AGC_PITCH * sizeof(float) + (size_t)group * sizeof(double) + (size_t)group * sizeof(int);
}

__device__ __forceinline__ float shfl_up_f64(float v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }
__device__ __forceinline__ float shfl_f64(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ float shfl_xor_f64(float v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}
// approximate sqrt / reciprocal sqrt (MUFU, ~2 ulp): the gain only has to hold 1e-5
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ unsigned long long agc_pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

// smoothed[i] = max|min(c_att*s + a_att*d, c_rel*s + a_rel*d): two FFMA2 (d and s are the scalar-broadcast
// operands, the coefficient pairs sit in uniform registers) and FMNMX; the dependent chain per step is
// FFMA2 -> FMNMX.
__device__ __forceinline__ unsigned long long agc_scaled(float d, unsigned long long aa) {   // (a_att*d, a_rel*d)
    unsigned long long t;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(agc_pack(d, d)), "l"(aa), "l"(0ULL));
    return t;
}
template <bool TAKE_MAX>
__device__ __forceinline__ float agc_step(float s, unsigned long long t, unsigned long long cc) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(agc_pack(s, s)), "l"(cc));
    const float2 br = *reinterpret_cast<float2*>(&t);
    return TAKE_MAX ? fmaxf(br.x, br.y) : fminf(br.x, br.y);
}
template <bool TAKE_MAX>
__device__ __forceinline__ float agc_group(float4& v, float s, unsigned long long cc, unsigned long long aa) {
    // the four products first: they do not depend on s and stay off the chain
    const unsigned long long t0 = agc_scaled(v.x, aa), t1 = agc_scaled(v.y, aa), t2 = agc_scaled(v.z, aa),
                             t3 = agc_scaled(v.w, aa);
    v.x = s = agc_step<TAKE_MAX>(s, t0, cc);
    v.y = s = agc_step<TAKE_MAX>(s, t1, cc);
    v.z = s = agc_step<TAKE_MAX>(s, t2, cc);
    v.w = s = agc_step<TAKE_MAX>(s, t3, cc);
    return s;
}
// Row layout: the 64 groups of four samples of a chunk are stored even groups first (slots 0..31), odd groups after
// (slots 32..63): in phase A a lane owns groups 2 lane and 2 lane + 1, i.e. slots lane and 32 + lane - consecutive
// lanes, consecutive 16-byte slots, conflict-free LDS.128 / STS.128 (8 contiguous samples per lane were a 2-way
// conflict); in phase B a lane walks its own row, where any order is conflict-free.
//
// One chunk of the recurrence, in place; returns the smoother state after it.  A shared-memory load issued one
// group ahead is not back in time (29 clk + the store queue in front of it, against 13.5 clk per step), so: 16
// steps per iteration, the NEXT 16 loaded before the current ones are computed.  `first`: smoothed[0] = desired[0] (:69).
template <bool TAKE_MAX>
__device__ __forceinline__ float agc_recur(float* row, int n, float s, bool first, unsigned long long cc, unsigned long long aa) {
    float4* ev = reinterpret_cast<float4*>(row);     // groups 0, 2, 4, ...
    float4* od = ev + 32;                            // groups 1, 3, 5, ...
    const int np = (n + 7) >> 3;                     // pairs of groups (the slots past n are computed, never read back)
    float4 a = ev[0], b = od[0];
    if (first) s = a.x;
    else a.x = s = agc_step<TAKE_MAX>(s, agc_scaled(a.x, aa), cc);
    a.y = s = agc_step<TAKE_MAX>(s, agc_scaled(a.y, aa), cc);
    a.z = s = agc_step<TAKE_MAX>(s, agc_scaled(a.z, aa), cc);
    a.w = s = agc_step<TAKE_MAX>(s, agc_scaled(a.w, aa), cc);
    ev[0] = a;
    s = agc_group<TAKE_MAX>(b, s, cc, aa); od[0] = b;
    int p = 1;
    if (p + 2 <= np) {
        float4 a0 = ev[p], b0 = od[p], a1 = ev[p + 1], b1 = od[p + 1];
        for (; p + 4 <= np; p += 2) {
            const float4 c0 = ev[p + 2], d0 = od[p + 2], c1 = ev[p + 3], d1 = od[p + 3];
            s = agc_group<TAKE_MAX>(a0, s, cc, aa); ev[p] = a0;
            s = agc_group<TAKE_MAX>(b0, s, cc, aa); od[p] = b0;
            s = agc_group<TAKE_MAX>(a1, s, cc, aa); ev[p + 1] = a1;
            s = agc_group<TAKE_MAX>(b1, s, cc, aa); od[p + 1] = b1;
            a0 = c0; b0 = d0; a1 = c1; b1 = d1;
        }
        s = agc_group<TAKE_MAX>(a0, s, cc, aa); ev[p] = a0;
        s = agc_group<TAKE_MAX>(b0, s, cc, aa); od[p] = b0;
        s = agc_group<TAKE_MAX>(a1, s, cc, aa); ev[p + 1] = a1;
        s = agc_group<TAKE_MAX>(b1, s, cc, aa); od[p + 1] = b1;
        p += 2;
    }
    for (; p < np; ++p) {
        a = ev[p]; b = od[p];
        s = agc_group<TAKE_MAX>(a, s, cc, aa); ev[p] = a;
        s = agc_group<TAKE_MAX>(b, s, cc, aa); od[p] = b;
    }
    return s;
}
// sample i of a chunk in the row layout above
__device__ __forceinline__ int agc_row_index(int i) { return ((i >> 2) & 1) * 128 + (i >> 3) * 4 + (i & 3); }

// (double)(v * v) without the conversion unit: the square in float32 as numpy computes it, then the float64 bit pattern
// by integer arithmetic - exponent re-biased by 896, mantissa moved up 29 bits ({u >> 3, u << 29} is u * 2^29).  Zero and
// float32 denormals come out as ~1e-38 instead of themselves: nothing next to any audible level, and the same value
// when the sample enters the window and when it leaves, so it cancels exactly.  F2F.F64.F32 shares the quarter-rate
// pipe with MUFU.RSQ and the float64 -> float32 conversions; this form runs on the integer pipe.
__device__ __forceinline__ double agc_sq64(float v) {
    const unsigned u = __float_as_uint(__fmul_rn(v, v));
    const unsigned long long w = (unsigned long long)u * 0x20000000ull + 0x3800000000000000ull;
    return __longlong_as_double((long long)w);
}

// 8 consecutive samples x[i .. i+8) of a row of n samples, zero outside [0, n)
__device__ __forceinline__ void agc_load8(float (&v)[8], const float* x, int i, int n) {
    const float* p = x + i;
    if (i >= 0 && i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (i + j >= 0 && i + j < n) ? __ldg(p + j) : 0.f;
    }
}
// the same for a segment the caller knows to be inside the row and 16-byte aligned (warp-uniform fact): no checks.
// STREAMING: last use of these samples, evict first (ld.global.cs)
template <bool STREAMING>
__device__ __forceinline__ void agc_load8v(float (&v)[8], const float* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = STREAMING ? __ldcs(q) : __ldg(q), b = STREAMING ? __ldcs(q + 1) : __ldg(q + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// y[i .. i+8) = v, only inside [0, n)
__device__ __forceinline__ void agc_store8(const float (&v)[8], float* y, int i, int n) {
    float* p = y + i;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i + j < n) p[j] = v[j];
    }
}
__device__ __forceinline__ void agc_store8v(const float (&v)[8], float* p) {   // (the output is not read again: streaming stores)
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
}

// pull a row segment into L2 (16-byte aligned interior of [r, r + n))
__device__ __forceinline__ void agc_prefetch_l2(const float* r, int n) {
    if (n <= 0) return;
    const unsigned long long lo = (reinterpret_cast<unsigned long long>(r) + 15ull) & ~15ull;
    const unsigned long long hi = reinterpret_cast<unsigned long long>(r + n) & ~15ull;
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
}
template <int VARIANT>
__global__ void __launch_bounds__(AGC_THREADS, AGC_MIN_CTAS) agc_kernel(const __grid_constant__ AgcParams P) {
    constexpr bool CHECKED = VARIANT == AGC_CHECKED;
    constexpr bool PROBE = VARIANT == AGC_PROBE_F32;
    using acc_t = typename std::conditional<PROBE, float, double>::type;
    constexpr int C = AGC_CHUNK;
    extern __shared__ __align__(16) unsigned char agc_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int n_warps = AGC_THREADS / 32;
    const int n = P.frames, G = P.group;
    float* rows = reinterpret_cast<float*>(agc_smem);
    double* carry = reinterpret_cast<double*>(agc_smem + (size_t)G * AGC_PITCH * sizeof(float));
    int* failed = reinterpret_cast<int*>(carry + G);
    const float cl = P.clip > 0.f ? P.clip : 3.0e38f;

    if (P.passthrough) {   // the reference's except branch for the whole launch: y = clip(x)
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x)
            for (int sl = warp; sl < G; sl += n_warps) {
                const int s = item * G + sl;
                if (s >= P.n_streams) break;
                const float* x = P.in + (long long)s * P.in_pitch;
                float* y = P.out + (long long)s * P.out_pitch;
                for (int i = 8 * lane; i < n; i += C) {
                    float v[8];
                    agc_load8(v, x, i, n);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(v[j], -cl), cl);
                    agc_store8(v, y, i, n);
                }
            }
        return;
    }

    const int n_chunks = (n + C - 1) / C;
    const int L = P.left, R1 = P.right + 1;
    const unsigned long long cc = agc_pack(P.c_att, P.c_rel), aa = agc_pack(P.a_att, P.a_rel);
    const bool take_max = P.a_att >= P.a_rel;
    // rows and window offsets on 16-byte boundaries: whole 256-sample segments inside the block load / store without checks
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(P.in) | reinterpret_cast<uintptr_t>(P.out)) & 15u) == 0 &&
                        ((P.in_pitch | P.out_pitch) & 3) == 0 && ((L | R1) & 3) == 0;
    const int k = 8 * lane;

    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int s0 = item * G;
        const int g_here = min(G, P.n_streams - s0);     // streams of this item
        float sm_state = 0.f;          // phase B: smoother state of stream s0 + tid
        bool bad = false;
        if (CHECKED && tid < G) failed[tid] = 0;
        for (int c = 0; c <= n_chunks; ++c) {
            const int c0 = c * C;
            // ---- phase A: apply chunk c-1, desired gains of chunk c
            const bool do_apply = c > 0, do_gain = c < n_chunks;
            const bool mid_v = vec_ok && do_apply && c0 <= n;
            const bool lead_v = vec_ok && do_gain && c0 + R1 + C <= n;
            const bool lag_v = vec_ok && do_gain && c0 - L >= 0 && c0 - L + C <= n;
            const bool lag_is_mid = L == C && do_apply && do_gain;
            const float* x = P.in + (long long)(s0 + warp) * P.in_pitch;
            float* y = P.out + (long long)(s0 + warp) * P.out_pitch + (c0 - C + k);
            for (int sl = warp; sl < g_here; sl += n_warps, x += n_warps * P.in_pitch, y += n_warps * P.out_pitch) {
                float* row = rows + (size_t)sl * AGC_PITCH + 4 * lane;     // this lane's two slots: row, row + 128
                float mid[8], lead[8], lag[8];
                if (do_apply) { if (mid_v) agc_load8v<false>(mid, x + (c0 - C + k)); else agc_load8(mid, x, c0 - C + k, n); }
                if (do_gain) {
                    if (lead_v) agc_load8v<false>(lead, x + (c0 + R1 + k)); else agc_load8(lead, x, c0 + R1 + k, n);
                    if (lag_is_mid) {
                        // window = 2 chunks: the samples that leave the window are the ones this pass multiplies
#pragma unroll
                        for (int j = 0; j < 8; ++j) lag[j] = mid[j];
                    } else if (lag_v) agc_load8v<true>(lag, x + (c0 - L + k));
                    else agc_load8(lag, x, c0 - L + k, n);
                }
                acc_t s_carry = 0;
                if (c == 0) {
                    // S(0) = sum of x^2 over [0, R] (the part of the first window inside the block)
                    const int m = min(R1, n);
                    for (int t0 = 0; t0 < m; t0 += C) {
                        float v[8];
                        agc_load8(v, x, t0 + k, m);
#pragma unroll
                        for (int j = 0; j < 8; ++j) s_carry += PROBE ? (acc_t)__fmul_rn(v[j], v[j]) : (acc_t)agc_sq64(v[j]);
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) s_carry += shfl_xor_f64(s_carry, off);
                } else {
                    s_carry = (acc_t)carry[sl];
                }
                if (do_apply && !(P.probe & 8)) {
                    const float4 g0 = *reinterpret_cast<const float4*>(row), g1 = *reinterpret_cast<const float4*>(row + 128);
                    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    // (the smoothed gain is a convex combination of desired gains, all <= max_gain: only the lower limit acts)
#pragma unroll
                    for (int j = 0; j < 8; ++j) mid[j] = fminf(fmaxf(__fmul_rn(mid[j], fmaxf(g[j], P.min_gain)), -cl), cl);
                    if (mid_v) agc_store8v(mid, y); else agc_store8(mid, y - (c0 - C + k), c0 - C + k, n);
                }
                __syncwarp();      // every lane has read its gains (and the carry) before the row is rewritten
                if (do_gain && !(P.probe & 4)) {
                    acc_t run[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // np.square of a float32 array stays float32 (:60); the sums are float64 (:62)
                        const acc_t d = PROBE ? (acc_t)(__fmul_rn(lead[j], lead[j]) - __fmul_rn(lag[j], lag[j])) : (acc_t)(agc_sq64(lead[j]) - agc_sq64(lag[j]));
                        run[j] = j == 0 ? d : run[j - 1] + d;
                    }
                    acc_t tot = run[7];
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const acc_t t = shfl_up_f64(tot, off);
                        if (lane >= off) tot += t;
                    }
                    acc_t excl = shfl_up_f64(tot, 1);
                    excl = (lane == 0 ? (acc_t)0 : excl) + s_carry;      // S(c0 + 8 lane)
                    const acc_t total = shfl_f64(tot, 31);
                    if (lane == 0) carry[sl] = (double)(s_carry + total);     // S(c0 + C)
                    float gq[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float sum = (float)(j == 0 ? excl : excl + run[j - 1]);
                        if (VARIANT == AGC_FAST || PROBE) {
                            // (a window sum that cancelled to a tiny negative number: rsqrt -> NaN -> fminf picks max_gain)
                            gq[j] = fminf(P.fast_scale * rsqrt_approx(sum), P.max_gain);
                        } else {
                            const float rms = sqrt_approx(fmaxf(sum, 0.f) * P.inv_window);
                            gq[j] = fminf(fmaxf(__fdividef(P.target, rms + 1e-10f), 0.f), P.max_gain);
                        }
                    }
                    *reinterpret_cast<float4*>(row) = make_float4(gq[0], gq[1], gq[2], gq[3]);
                    *reinterpret_cast<float4*>(row + 128) = make_float4(gq[4], gq[5], gq[6], gq[7]);
                }
            }
            if (c == n_chunks) break;
            __syncthreads();
            // ---- phase B: the attack / release recurrence over the chunk, one lane per stream; the other warps
            // pull the next chunk's lead samples (next item: its first window and chunk) into L2 meanwhile
            const int nc = min(C, n - c0);
            if (warp == 0 && !(P.probe & 2)) {
                if (lane < G && s0 + lane < P.n_streams) {
                    float* row = rows + (size_t)lane * AGC_PITCH;
                    if (!CHECKED) {
                        sm_state = take_max ? agc_recur<true>(row, nc, sm_state, c == 0, cc, aa)
                                            : agc_recur<false>(row, nc, sm_state, c == 0, cc, aa);
                    } else {
                        float sv = sm_state;
                        int i = 0;
                        if (c == 0) { sv = row[0]; i = 1; }     // (index 0 is slot 0)
#pragma unroll 4
                        for (; i < nc; ++i) {
                            const float d = row[agc_row_index(i)];
                            const bool att = d > sv;
                            bad |= att ? !P.att_ok : !P.rel_ok;
                            sv = fmaf(att ? P.c_att : P.c_rel, sv, (att ? P.a_att : P.a_rel) * d);
                            row[agc_row_index(i)] = sv;
                        }
                        sm_state = sv;
                        if (bad) failed[lane] = 1;
                    }
                }
            } else if (warp == 1) {
                if (lane < G && s0 + lane < P.n_streams && c + 1 < n_chunks) {
                    const int a = c0 + C + R1;
                    agc_prefetch_l2(P.in + (long long)(s0 + lane) * P.in_pitch + a, min(C, n - a));
                }
            } else if (warp == 2 && c == n_chunks - 1) {
                const int sn = (item + gridDim.x) * G + lane;
                if (item + gridDim.x < P.n_items && lane < G && sn < P.n_streams)
                    agc_prefetch_l2(P.in + (long long)sn * P.in_pitch, min(n, R1 + C));
            }
            __syncthreads();
        }
        if (CHECKED) {
            // streams that took a branch the reference cannot evaluate: the block comes back unchanged (then the call
            // site's clip); chunks already written are overwritten
            __syncthreads();
            for (int sl = warp; sl < G; sl += n_warps) {
                const int s = s0 + sl;
                if (s >= P.n_streams) break;
                if (!failed[sl]) continue;
                const float* x = P.in + (long long)s * P.in_pitch;
                float* y = P.out + (long long)s * P.out_pitch;
                for (int i = 8 * lane; i < n; i += C) {
                    float v[8];
                    agc_load8(v, x, i, n);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(v[j], -cl), cl);
                    agc_store8(v, y, i, n);
                }
            }
        }
        __syncthreads();   // rows / flags are rewritten by the next item
    }
}

}  // namespace afp
