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
// One work item = `group` streams x one block.  Shared memory holds one float row per stream (odd
// pitch: lane = stream accesses are conflict-free) plus one float64 prefix-sum row per warp.
//   phase 1  warp per stream: a lane takes 16 consecutive samples (four 16-byte loads) -> x^2 in float32
//            as numpy does -> float64 running sums in registers -> one warp scan of the lane totals per
//            512 samples -> exclusive prefix sums in the warp's scratch row (one pad double per 16, so the
//            lane stride is 17 doubles: conflict-free) -> window sum as a difference of two prefix sums
//            (float64: a quiet window after a loud passage keeps its digits) -> desired gain, float32,
//            into the stream's row.  FAST instantiation: target*rsqrt(mean) - the reference's "+ 1e-10"
//            under the division only matters where the gain is clamped to max_gain anyway (host check).
//   phase 2  lane per stream: the recurrence, in place; both branches in one packed FFMA2
//            (s * (1-a_att, 1-a_rel) + d * (a_att, a_rel)), then FMNMX
//   phase 3  all threads over (stream, 4-sample chunk), 8 loads in flight each: x again (an L2 hit, the
//            item was read a few microseconds ago) times the limited gain, clip, coalesced 16-byte stores
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
    int pitch;                 // floats per stream row in shared memory (odd)
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

// smoothed[i] = max|min(c_att*s + a_att*d, c_rel*s + a_rel*d): FMUL2, FFMA2 (s is the scalar-broadcast
// operand), FMNMX; the dependent chain per step is FFMA2 -> FMNMX
template <bool TAKE_MAX>
__device__ __forceinline__ float agc_recur(float* row, int n, float s, unsigned long long cc, unsigned long long aa) {
#pragma unroll 16
    for (int i = 1; i < n; ++i) {
        const float d = row[i];
        unsigned long long t;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(agc_pack(d, d)), "l"(aa));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(agc_pack(s, s)), "l"(cc));
        const float2 br = *reinterpret_cast<float2*>(&t);
        s = TAKE_MAX ? fmaxf(br.x, br.y) : fminf(br.x, br.y);
        row[i] = s;
    }
    return s;
}

constexpr int AGC_LANE_RUN = 16;                   // consecutive samples per lane in the prefix scan
constexpr int AGC_SCAN = 32 * AGC_LANE_RUN;        // samples per warp scan
// index of prefix sum k in the scratch row: one pad double per 16
__host__ __device__ constexpr int agc_phys(int k) { return k + (k >> 4); }

// SB = 512-sample scans in flight per warp: every load of up to 512*SB samples of a stream is issued
// before the first use
template <int VARIANT, int SB>
__global__ void agc_kernel(const __grid_constant__ AgcParams P) {
    constexpr bool CHECKED = VARIANT == AGC_CHECKED;
    extern __shared__ __align__(16) unsigned char agc_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
    const int n = P.frames, G = P.group;
    float* rows = reinterpret_cast<float*>(agc_smem);
    const size_t rows_bytes = (((size_t)G * P.pitch * sizeof(float)) + 7) & ~(size_t)7;
    double* pfx = reinterpret_cast<double*>(agc_smem + rows_bytes) + (size_t)warp * P.scratch;
    int* failed = reinterpret_cast<int*>(agc_smem + rows_bytes + (size_t)n_warps * P.scratch * sizeof(double));

    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int s0 = item * G;
        if (!P.passthrough) {
            // ---- phase 1: desired gain of every sample
            for (int sl = warp; sl < G; sl += n_warps) {
                const int s = s0 + sl;
                if (s >= P.n_streams) break;
                const float* x = P.in + (long long)s * P.in_pitch;
                const bool al = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
                if (lane == 0) { pfx[0] = 0.0; if (CHECKED) failed[sl] = 0; }
                double carry = 0.0;
                for (int base = 0; base < n; base += AGC_SCAN * SB) {
                    float v[SB][AGC_LANE_RUN];
#pragma unroll
                    for (int b = 0; b < SB; ++b) {
                        const int i = base + AGC_SCAN * b + AGC_LANE_RUN * lane;
                        if (al && i + AGC_LANE_RUN <= n) {
#pragma unroll
                            for (int q = 0; q < AGC_LANE_RUN / 4; ++q) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(x + i) + q);
                                v[b][4 * q] = t.x; v[b][4 * q + 1] = t.y; v[b][4 * q + 2] = t.z; v[b][4 * q + 3] = t.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < AGC_LANE_RUN; ++j) v[b][j] = i + j < n ? __ldg(x + i + j) : 0.f;
                        }
                    }
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
                        excl = (lane == 0 ? 0.0 : excl) + carry;
                        const int i = base + AGC_SCAN * b + AGC_LANE_RUN * lane;
                        if (i + AGC_LANE_RUN <= n) {
#pragma unroll
                            for (int j = 0; j < AGC_LANE_RUN; ++j) pfx[agc_phys(i + j + 1)] = excl + run[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < AGC_LANE_RUN; ++j)
                                if (i + j < n) pfx[agc_phys(i + j + 1)] = excl + run[j];
                        }
                        carry += shfl_f64(tot, 31);
                    }
                }
                __syncwarp();
                float* row = rows + (size_t)sl * P.pitch;
#pragma unroll 4
                for (int i = lane; i < n; i += 32) {
                    const int hi = min(i + P.right, n - 1) + 1, lo = max(i - P.left, 0);
                    const float sum = (float)(pfx[agc_phys(hi)] - pfx[agc_phys(lo)]);
                    if (VARIANT == AGC_FAST) {
                        row[i] = fminf(P.fast_scale * rsqrt_approx(sum), P.max_gain);
                    } else {
                        const float rms = sqrt_approx(sum * P.inv_window);
                        row[i] = fminf(fmaxf(__fdividef(P.target, rms + 1e-10f), 0.f), P.max_gain);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
            // ---- phase 2: the attack / release recurrence, one lane per stream
            if (tid < G && s0 + tid < P.n_streams) {
                float* row = rows + (size_t)tid * P.pitch;
                float s = row[0];
                if (!CHECKED) {
                    const unsigned long long cc = agc_pack(P.c_att, P.c_rel), aa = agc_pack(P.a_att, P.a_rel);
                    if (P.a_att >= P.a_rel) s = agc_recur<true>(row, n, s, cc, aa);
                    else                    s = agc_recur<false>(row, n, s, cc, aa);
                } else {
                    bool bad = false;
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
        // ---- phase 3: apply.  All threads over (stream, 4-sample chunk) with 8 loads in flight each: the
        // x rows come back from L2, the only latency left to hide in this phase
        const float cl = P.clip > 0.f ? P.clip : 3.0e38f;
        const bool vec = (n & 3) == 0 && ((P.in_pitch | P.out_pitch) & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(P.in) | reinterpret_cast<uintptr_t>(P.out)) & 15u) == 0;
        if (vec) {
            const int n4 = n >> 2, g_here = min(G, P.n_streams - s0), total = g_here * n4;
            constexpr int LB = 8;
            for (int base = tid; base < total; base += blockDim.x * LB) {
                float4 v[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u) {
                    const int idx = base + u * blockDim.x;
                    if (idx < total) {
                        const int sl = idx / n4, c4 = idx - sl * n4;
                        v[u] = __ldg(reinterpret_cast<const float4*>(P.in + (long long)(s0 + sl) * P.in_pitch) + c4);
                    }
                }
#pragma unroll
                for (int u = 0; u < LB; ++u) {
                    const int idx = base + u * blockDim.x;
                    if (idx < total) {
                        const int sl = idx / n4, c4 = idx - sl * n4;
                        const float* row = rows + (size_t)sl * P.pitch + 4 * c4;
                        const bool unity = P.passthrough || (CHECKED && failed[sl] != 0);
                        float4 y;
                        y.x = fminf(fmaxf(__fmul_rn(v[u].x, unity ? 1.f : fminf(fmaxf(row[0], P.min_gain), P.max_gain)), -cl), cl);
                        y.y = fminf(fmaxf(__fmul_rn(v[u].y, unity ? 1.f : fminf(fmaxf(row[1], P.min_gain), P.max_gain)), -cl), cl);
                        y.z = fminf(fmaxf(__fmul_rn(v[u].z, unity ? 1.f : fminf(fmaxf(row[2], P.min_gain), P.max_gain)), -cl), cl);
                        y.w = fminf(fmaxf(__fmul_rn(v[u].w, unity ? 1.f : fminf(fmaxf(row[3], P.min_gain), P.max_gain)), -cl), cl);
                        reinterpret_cast<float4*>(P.out + (long long)(s0 + sl) * P.out_pitch)[c4] = y;
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
