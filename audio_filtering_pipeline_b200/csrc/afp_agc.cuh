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
//   phase 1  warp per stream: x (coalesced 16-byte loads) -> x^2 in float32 as numpy does -> float64
//            inclusive prefix sums (warp shuffles; 4 independent 128-sample scans in flight) -> window
//            sum as a difference of two prefix sums (float64: a quiet window after a loud passage keeps
//            its digits) -> desired gain, float32, into the stream's row
//   phase 2  lane per stream: the recurrence, in place
//   phase 3  warp per stream: x again (an L2 hit, the item was read a few microseconds ago) times the
//            limited gain, clip, coalesced store
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
    float* out;                // [n_streams][out_pitch]   (may alias `in`)
    long long out_pitch;
    int n_streams, frames;
    int window, left, right;   // window length; samples before / after the centre: left + 1 + right = window
    int group;                 // streams per work item
    int pitch;                 // floats per stream row in shared memory (odd)
    int n_items;
    float inv_window;
    float target, max_gain, min_gain, clip;   // clip <= 0: no output clip
    float a_att, c_att, a_rel, c_rel;          // alpha and 1 - alpha of the two branches
    int att_ok, rel_ok;        // 0: that branch raises in the reference
    int passthrough;           // window > frames
};

typedef void (*agc_kern_t)(const AgcParams);

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

// approximate sqrt / divide (MUFU + one multiply, ~2 ulp): the gain only has to hold 1e-5
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// SB = independent 128-sample scans in flight per warp: every load of a block of up to 128*SB samples is
// issued before the first use (one exposed memory latency per stream when frames <= 1024)
template <bool CHECKED, int SB>
__global__ void agc_kernel(const __grid_constant__ AgcParams P) {
    constexpr int AGC_SCAN_BLOCKS = SB;
    extern __shared__ __align__(16) unsigned char agc_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
    const int n = P.frames, G = P.group;
    float* rows = reinterpret_cast<float*>(agc_smem);
    const size_t rows_bytes = (((size_t)G * P.pitch * sizeof(float)) + 7) & ~(size_t)7;
    double* pfx = reinterpret_cast<double*>(agc_smem + rows_bytes) + (size_t)warp * (n + 1);
    int* failed = reinterpret_cast<int*>(agc_smem + rows_bytes + (size_t)n_warps * (n + 1) * sizeof(double));

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
                for (int base = 0; base < n; base += 128 * AGC_SCAN_BLOCKS) {
                    double p[AGC_SCAN_BLOCKS][4], tot[AGC_SCAN_BLOCKS];
#pragma unroll
                    for (int b = 0; b < AGC_SCAN_BLOCKS; ++b) {
                        const int i = base + 128 * b + 4 * lane;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (al && i + 3 < n) v = __ldg(reinterpret_cast<const float4*>(x + i));
                        else {
                            if (i < n) v.x = __ldg(x + i);
                            if (i + 1 < n) v.y = __ldg(x + i + 1);
                            if (i + 2 < n) v.z = __ldg(x + i + 2);
                            if (i + 3 < n) v.w = __ldg(x + i + 3);
                        }
                        // np.square of a float32 array stays float32 (:60); the sums are float64 (:62)
                        p[b][0] = (double)__fmul_rn(v.x, v.x);
                        p[b][1] = p[b][0] + (double)__fmul_rn(v.y, v.y);
                        p[b][2] = p[b][1] + (double)__fmul_rn(v.z, v.z);
                        p[b][3] = p[b][2] + (double)__fmul_rn(v.w, v.w);
                        tot[b] = p[b][3];
                    }
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1)
#pragma unroll
                        for (int b = 0; b < AGC_SCAN_BLOCKS; ++b) {
                            const double t = shfl_up_f64(tot[b], off);
                            if (lane >= off) tot[b] += t;
                        }
#pragma unroll
                    for (int b = 0; b < AGC_SCAN_BLOCKS; ++b) {
                        double excl = shfl_up_f64(tot[b], 1);
                        excl = (lane == 0 ? 0.0 : excl) + carry;
                        const int i = base + 128 * b + 4 * lane;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (i + j < n) pfx[i + j + 1] = excl + p[b][j];
                        carry += shfl_f64(tot[b], 31);
                    }
                }
                __syncwarp();
                float* row = rows + (size_t)sl * P.pitch;
#pragma unroll 4
                for (int i = lane; i < n; i += 32) {
                    const int hi = min(i + P.right, n - 1), lo = max(i - P.left, 0);
                    const double sum = pfx[hi + 1] - pfx[lo];
                    const float rms = sqrt_approx((float)sum * P.inv_window);
                    row[i] = fminf(fmaxf(__fdividef(P.target, rms + 1e-10f), 0.f), P.max_gain);
                }
                __syncwarp();
            }
            __syncthreads();
            // ---- phase 2: the attack / release recurrence, one lane per stream
            if (tid < G && s0 + tid < P.n_streams) {
                float* row = rows + (size_t)tid * P.pitch;
                float s = row[0];
                if (!CHECKED) {
                    const float aA = P.a_att, cA = P.c_att, aR = P.a_rel, cR = P.c_rel;
                    if (aA >= aR) {
#pragma unroll 8
                        for (int i = 1; i < n; ++i) {
                            const float d = row[i];
                            s = fmaxf(fmaf(cA, s, aA * d), fmaf(cR, s, aR * d));
                            row[i] = s;
                        }
                    } else {
#pragma unroll 8
                        for (int i = 1; i < n; ++i) {
                            const float d = row[i];
                            s = fminf(fmaf(cA, s, aA * d), fmaf(cR, s, aR * d));
                            row[i] = s;
                        }
                    }
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
