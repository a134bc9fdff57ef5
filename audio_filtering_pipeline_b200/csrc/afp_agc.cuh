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
// FFMA -> FMNMX (8 clk).  Parallelism comes from the batch: one LANE per stream in that phase.
//
// One work item = `group` streams x one block.  Shared memory holds one float row per stream (pitch = an
// odd number of 16-byte units: lane = stream 128-bit accesses are conflict-free) plus one float64
// prefix-sum row per warp.
//   phase 1  warp per stream: a lane takes 16 consecutive samples (four 16-byte loads) -> x^2 in float32
//            as numpy does -> float64 running sums in registers -> one warp scan of the lane totals per
//            512 samples -> exclusive prefix sums in the warp's scratch row (one pad double per 16, so the
//            lane stride is 17 doubles: conflict-free) -> window sum as a difference of two prefix sums
//            (float64: a quiet window after a loud passage keeps its digits) -> desired gain, float32,
//            into the stream's row.  FAST instantiation: target*rsqrt(mean) - the reference's "+ 1e-10"
//            under the division only matters where the gain is clamped to max_gain anyway (host check).
//   phase 2  lane per stream: the recurrence, in place; both branches in packed FFMA2s
//            (s * (1-a_att, 1-a_rel) + d * (a_att, a_rel)), then FMNMX; four steps per LDS.128 / STS.128
//   phase 3  warp per stream, 4-sample chunks, 8 loads in flight per lane: x again (an L2 hit, the item
//            was read a few microseconds ago) times the limited gain, clip, coalesced 16-byte stores
// Two CTAs per SM when the rows fit (host plan in afp_abi.cu), so one CTA's serial phase 2 overlaps the
// other's memory phases.  HBM-bound by design: 8 bytes per sample, no intermediate leaves the SM.
//
// The reference's error path is kept: int(attack*window) == 0 (or release) makes its loop divide by
// zero the first time that branch is taken; it catches the exception and returns the block unchanged
// (:87-89).  CHECKED instantiation: the recurrence is evaluated with the explicit comparison, a stream
// that takes an invalid branch is flagged and leaves with gain 1.  A window longer than the block fails
// in the reference when the gains are applied (shape mismatch): whole launch passes through.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace afp {

struct AgcParams {
    const float* in;           // [n_streams][in_pitch]
    long long in_pitch;
    float* out;                // [n_streams][out_pitch]   (must not overlap `in`)
    long long out_pitch;
    int n_streams, frames;
    int window, left, right;   // window length; samples before / after the centre: left + 1 + right = window
    int group;                 // streams per work item
    int pitch;                 // floats per stream row in shared memory: 4 x an odd number
    int n_items;
    int scratch;               // doubles per warp in the prefix-sum scratch
    float inv_window;
    float fast_scale;          // target * sqrt(window): desired = min(fast_scale * rsqrt(window sum), max_gain)
    float target, max_gain, min_gain, clip;   // clip <= 0: no output clip
    float a_att, c_att, a_rel, c_rel;          // alpha and 1 - alpha of the two branches
    int att_ok, rel_ok;        // 0: that branch raises in the reference
    int passthrough;           // window > frames
};

typedef void (*agc_kern_t)(const AgcParams);

// kernel variants
constexpr int AGC_FAST = 0;      // both constants valid, epsilon-free gain
constexpr int AGC_EXACT = 1;     // both constants valid, target / (rms + 1e-10)
constexpr int AGC_CHECKED = 2;   // a zero-sample constant: explicit comparison, failing streams pass through

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
// FFMA2 -> FMNMX.  Rows are read and written four steps at a time (LDS.128 / STS.128: the row pitch is an
// odd number of 16-byte units, so lane = stream accesses are conflict-free): 3.5 instructions per step -
// this phase runs on one warp per CTA and is bound by that warp's issue rate.
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
// The chain costs 13.5 clk per step on this part (bench/microbench/agc_chain.cu: packed and scalar forms
// alike); a shared-memory load issued one group ahead is not back in time (29 clk + the store queue in
// front of it), which made the first version 25 clk per step.  So: 16 steps per iteration, the NEXT 16
// loaded before the current ones are computed.
template <bool TAKE_MAX>
__device__ __forceinline__ void agc_recur(float* row, int n, unsigned long long cc, unsigned long long aa) {
    float4* r4 = reinterpret_cast<float4*>(row);
    const int n4 = (n + 3) >> 2;                     // (the slots past n in the last group are never read back)
    float4 v = r4[0];
    float s = v.x;                                   // smoothed[0] = desired[0]   (:69)
    v.y = s = agc_step<TAKE_MAX>(s, agc_scaled(v.y, aa), cc);
    v.z = s = agc_step<TAKE_MAX>(s, agc_scaled(v.z, aa), cc);
    v.w = s = agc_step<TAKE_MAX>(s, agc_scaled(v.w, aa), cc);
    r4[0] = v;
    int q = 1;
    if (q + 4 <= n4) {
        float4 c0 = r4[q], c1 = r4[q + 1], c2 = r4[q + 2], c3 = r4[q + 3];
        for (; q + 8 <= n4; q += 4) {
            const float4 d0 = r4[q + 4], d1 = r4[q + 5], d2 = r4[q + 6], d3 = r4[q + 7];
            s = agc_group<TAKE_MAX>(c0, s, cc, aa); r4[q] = c0;
            s = agc_group<TAKE_MAX>(c1, s, cc, aa); r4[q + 1] = c1;
            s = agc_group<TAKE_MAX>(c2, s, cc, aa); r4[q + 2] = c2;
            s = agc_group<TAKE_MAX>(c3, s, cc, aa); r4[q + 3] = c3;
            c0 = d0; c1 = d1; c2 = d2; c3 = d3;
        }
        s = agc_group<TAKE_MAX>(c0, s, cc, aa); r4[q] = c0;
        s = agc_group<TAKE_MAX>(c1, s, cc, aa); r4[q + 1] = c1;
        s = agc_group<TAKE_MAX>(c2, s, cc, aa); r4[q + 2] = c2;
        s = agc_group<TAKE_MAX>(c3, s, cc, aa); r4[q + 3] = c3;
        q += 4;
    }
    for (; q < n4; ++q) {
        v = r4[q];
        s = agc_group<TAKE_MAX>(v, s, cc, aa);
        r4[q] = v;
    }
}

constexpr int AGC_LANE_RUN = 16;                   // consecutive samples per lane in the prefix scan (agc_load_run)
constexpr int AGC_SCAN = 32 * AGC_LANE_RUN;        // samples per warp scan
// index of prefix sum k in the scratch row: one pad double per 16
__host__ __device__ constexpr int agc_phys(int k) { return k + (k >> 4); }

// desired gain over samples [i0, i1) of one stream (a warp; lanes stride 32).  LO0: the window starts
// before the block (lower prefix sum is 0); HIN: it ends past the block (upper prefix sum is the total).
// A step of 32 samples is a step of exactly 34 doubles in the padded scratch row, so the addresses are
// affine in the iteration count.
template <int VARIANT, bool LO0, bool HIN>
__device__ __forceinline__ void agc_gain_span(const AgcParams& P, const double* pfx, float* row, int i0, int i1, int lane,
                                              double total) {
    int i = i0 + lane;
    if (i >= i1) return;
    const double* ph = pfx + agc_phys(HIN ? 0 : i + P.right + 1);
    const double* pl = pfx + agc_phys(LO0 ? 0 : i - P.left);
    float* r = row + i;
#pragma unroll 4
    for (; i < i1; i += 32, ph += 34, pl += 34, r += 32) {
        const double hi = HIN ? total : *ph, lo = LO0 ? 0.0 : *pl;
        const float sum = (float)(hi - lo);
        if (VARIANT == AGC_FAST) {
            *r = fminf(P.fast_scale * rsqrt_approx(sum), P.max_gain);
        } else {
            const float rms = sqrt_approx(sum * P.inv_window);
            *r = fminf(fmaxf(__fdividef(P.target, rms + 1e-10f), 0.f), P.max_gain);
        }
    }
}

// 16 consecutive samples per lane and 512-sample scan block, zero past the end of the row
template <int SB>
__device__ __forceinline__ void agc_load_run(float (&v)[SB][16], const float* x, int base, int n, int lane) {
    const bool al = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
#pragma unroll
    for (int b = 0; b < SB; ++b) {
        const int i = base + 512 * b + 16 * lane;
        if (al && i + 16 <= n) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(x + i) + q);
                v[b][4 * q] = t.x; v[b][4 * q + 1] = t.y; v[b][4 * q + 2] = t.z; v[b][4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[b][j] = i + j < n ? __ldg(x + i + j) : 0.f;
        }
    }
}

// SB = 512-sample scans in flight per warp: every load of up to 512*SB samples of a stream is issued
// before the first use.
// pull a row into L2 (16-byte aligned interior of [r, r + n))
__device__ __forceinline__ void agc_prefetch_l2(const float* r, int n) {
    const unsigned long long lo = (reinterpret_cast<unsigned long long>(r) + 15ull) & ~15ull;
    const unsigned long long hi = reinterpret_cast<unsigned long long>(r + n) & ~15ull;
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
}

// TEAM = warps per stream in phase 1 (1 or 2).  The float64 scratch row is what limits the resident warps
// (8.7 KB per 1024-frame stream in flight); with TEAM = 2 two warps share one row and take half a stream
// each - the second adds the first one's total as its carry - so the SM holds twice the warps in the same
// shared memory.  The members meet at a named barrier (id 1 + team, 64 threads) three times per stream:
// carry published, prefix row complete, prefix row free again.
template <int TEAM>
__device__ __forceinline__ void agc_team_sync(int team) {
    if (TEAM == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(32 * TEAM) : "memory");
}

template <int VARIANT, int SB, int TEAM>
__global__ void __launch_bounds__(TEAM == 2 ? 256 : 128, TEAM == 2 ? 2 : 3) agc_kernel(const __grid_constant__ AgcParams P) {
    constexpr bool CHECKED = VARIANT == AGC_CHECKED;
    extern __shared__ __align__(16) unsigned char agc_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
    const int team = warp / TEAM, member = warp % TEAM, n_teams = n_warps / TEAM;
    const int n = P.frames, G = P.group;
    float* rows = reinterpret_cast<float*>(agc_smem);
    const size_t rows_bytes = (((size_t)G * P.pitch * sizeof(float)) + 7) & ~(size_t)7;
    double* pfx = reinterpret_cast<double*>(agc_smem + rows_bytes) + (size_t)team * P.scratch;
    double* team_tot = reinterpret_cast<double*>(agc_smem + rows_bytes) + (size_t)n_teams * P.scratch + team;
    int* failed = reinterpret_cast<int*>(agc_smem + rows_bytes + ((size_t)n_teams * P.scratch + n_teams) * sizeof(double));
    // this warp's part of a stream in phase 1: [h0, h1), a multiple of 16 samples from the start
    const int nh = TEAM == 1 ? n : ((((n + 1) >> 1) + 15) & ~15);
    const int h0 = min(member * nh, n), h1 = min(h0 + nh, n);

    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int s0 = item * G;
        if (!P.passthrough) {
            // ---- phase 1: desired gain of every sample
            for (int sl = team; sl < G; sl += n_teams) {
                const int s = s0 + sl;
                if (s >= P.n_streams) break;
                const float* x = P.in + (long long)s * P.in_pitch;
                if (member == 0 && lane == 0) { pfx[0] = 0.0; if (CHECKED) failed[sl] = 0; }
                double carry = 0.0;
                bool have_carry = member == 0;
                for (int base = h0; base < h1; base += AGC_SCAN * SB) {
                    float v[SB][AGC_LANE_RUN];
                    agc_load_run<SB>(v, x, base, h1, lane);
#pragma unroll
                    for (int b = 0; b < SB; ++b) {
                        // np.square of a float32 array stays float32 (:60); the sums are float64 (:62)
                        double run[AGC_LANE_RUN];
                        run[0] = (double)__fmul_rn(v[b][0], v[b][0]);
#pragma unroll
                        for (int j = 1; j < AGC_LANE_RUN; ++j) run[j] = run[j - 1] + (double)__fmul_rn(v[b][j], v[b][j]);
                        double tot = run[AGC_LANE_RUN - 1];
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const double t = shfl_up_f64(tot, off);
                            if (lane >= off) tot += t;
                        }
                        double excl = shfl_up_f64(tot, 1);
                        if (TEAM == 2 && !have_carry) {      // second member: the first half's total, once
                            agc_team_sync<TEAM>(team);
                            carry = *team_tot;
                            have_carry = true;
                        }
                        excl = (lane == 0 ? 0.0 : excl) + carry;
                        const int i = base + AGC_SCAN * b + AGC_LANE_RUN * lane;
                        if (i + AGC_LANE_RUN <= h1) {
#pragma unroll
                            for (int j = 0; j < AGC_LANE_RUN; ++j) pfx[agc_phys(i + j + 1)] = excl + run[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < AGC_LANE_RUN; ++j)
                                if (i + j < h1) pfx[agc_phys(i + j + 1)] = excl + run[j];
                        }
                        carry += shfl_f64(tot, 31);
                    }
                }
                if (TEAM == 2) {
                    if (member == 0) { if (lane == 0) *team_tot = carry; agc_team_sync<TEAM>(team); }
                    else if (!have_carry) agc_team_sync<TEAM>(team);      // (empty second half)
                }
                agc_team_sync<TEAM>(team);                                // prefix row complete
                float* row = rows + (size_t)sl * P.pitch;
                // three index ranges, so that the loops carry no clamps: [0, m1) / [m1, m2) / [m2, n) with
                // m1, m2 = the first sample whose window no longer starts before the block (left) and the
                // first one whose window ends past it (n - 1 - right), in order; each member its own part
                {
                    const int hc = max(n - 1 - P.right, 0), lc = min(P.left, n);
                    const int m1 = min(lc, hc), m2 = max(lc, hc);
                    const double total = pfx[agc_phys(n)];
                    agc_gain_span<VARIANT, true, false>(P, pfx, row, max(0, h0), min(m1, h1), lane, total);
                    if (lc <= hc) agc_gain_span<VARIANT, false, false>(P, pfx, row, max(m1, h0), min(m2, h1), lane, total);
                    else          agc_gain_span<VARIANT, true, true>(P, pfx, row, max(m1, h0), min(m2, h1), lane, total);   // window covers the block
                    agc_gain_span<VARIANT, false, true>(P, pfx, row, max(m2, h0), min(n, h1), lane, total);
                }
                agc_team_sync<TEAM>(team);                                // prefix row free for the next stream
            }
            __syncthreads();
            // ---- phase 2: the attack / release recurrence, one lane per stream.  Meanwhile the last warp (it
            // has no stream of its own here) pulls the next item's rows into L2, so that phase 1 of the next
            // item starts from L2 instead of HBM latency.
            if (warp == n_warps - 1) {
                const int nitem = item + gridDim.x;
                if (nitem < P.n_items)
                    for (int sl = lane; sl < G; sl += 32) {
                        const int sn = nitem * G + sl;
                        if (sn < P.n_streams) agc_prefetch_l2(P.in + (long long)sn * P.in_pitch, n);
                    }
            }
            if (tid < G && s0 + tid < P.n_streams) {
                float* row = rows + (size_t)tid * P.pitch;
                if (!CHECKED) {
                    const unsigned long long cc = agc_pack(P.c_att, P.c_rel), aa = agc_pack(P.a_att, P.a_rel);
                    if (P.a_att >= P.a_rel) agc_recur<true>(row, n, cc, aa);
                    else                    agc_recur<false>(row, n, cc, aa);
                } else {
                    bool bad = false;
                    float s = row[0];
#pragma unroll 4
                    for (int i = 1; i < n; ++i) {
                        const float d = row[i];
                        const bool att = d > s;
                        bad |= att ? !P.att_ok : !P.rel_ok;
                        s = fmaf(att ? P.c_att : P.c_rel, s, (att ? P.a_att : P.a_rel) * d);
                        row[i] = s;
                    }
                    if (bad) failed[tid] = 1;
                }
            }
            __syncthreads();
        }
        // ---- phase 3: apply.  Warp per stream, 4-sample chunks, 8 loads in flight per lane: the x rows come
        // back from L2, the only latency left to hide in this phase
        const float cl = P.clip > 0.f ? P.clip : 3.0e38f;
        const bool vec = (n & 3) == 0 && ((P.in_pitch | P.out_pitch) & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(P.in) | reinterpret_cast<uintptr_t>(P.out)) & 15u) == 0;
        if (vec) {
            const int n4 = n >> 2;
            constexpr int LB = 8;
            for (int sl = warp; sl < G; sl += n_warps) {
                const int s = s0 + sl;
                if (s >= P.n_streams) break;
                const float4* x4 = reinterpret_cast<const float4*>(P.in + (long long)s * P.in_pitch);
                float4* y4 = reinterpret_cast<float4*>(P.out + (long long)s * P.out_pitch);
                const float* row = rows + (size_t)sl * P.pitch;
                const bool unity = P.passthrough || (CHECKED && failed[sl] != 0);
                const float g_lo = unity ? 1.f : P.min_gain, g_hi = unity ? 1.f : P.max_gain;   // gain pinned to 1
                for (int c0 = lane; c0 < n4; c0 += 32 * LB) {
                    float4 v[LB];
#pragma unroll
                    for (int u = 0; u < LB; ++u)
                        if (c0 + 32 * u < n4) v[u] = __ldg(x4 + c0 + 32 * u);
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        const int c4 = c0 + 32 * u;
                        if (c4 < n4) {
                            const float4 g = *reinterpret_cast<const float4*>(row + 4 * c4);
                            float4 y;
                            y.x = fminf(fmaxf(__fmul_rn(v[u].x, fminf(fmaxf(g.x, g_lo), g_hi)), -cl), cl);
                            y.y = fminf(fmaxf(__fmul_rn(v[u].y, fminf(fmaxf(g.y, g_lo), g_hi)), -cl), cl);
                            y.z = fminf(fmaxf(__fmul_rn(v[u].z, fminf(fmaxf(g.z, g_lo), g_hi)), -cl), cl);
                            y.w = fminf(fmaxf(__fmul_rn(v[u].w, fminf(fmaxf(g.w, g_lo), g_hi)), -cl), cl);
                            y4[c4] = y;
                        }
                    }
                }
            }
        } else {
            for (int sl = warp; sl < G; sl += n_warps) {
                const int s = s0 + sl;
                if (s >= P.n_streams) break;
                const float* x = P.in + (long long)s * P.in_pitch;
                float* y = P.out + (long long)s * P.out_pitch;
                const float* row = rows + (size_t)sl * P.pitch;
                const bool unity = P.passthrough || (CHECKED && failed[sl] != 0);
#pragma unroll 4
                for (int i = lane; i < n; i += 32) {
                    const float gain = unity ? 1.f : fminf(fmaxf(row[i], P.min_gain), P.max_gain);
                    y[i] = fminf(fmaxf(__fmul_rn(__ldg(x + i), gain), -cl), cl);
                }
            }
        }
        __syncthreads();   // rows are rewritten by the next item
    }
}

}  // namespace afp
