// Waterfall row of one block per stream: out[k] = 20*log10(|rfft(x * hanning(n))[k]| + 1e-12), k = 0..n/2
// (stream_process_GUI_Waterfall.py:471-476).  One CTA per stream at a time (persistent over streams): the
// windowed block goes to shared memory in bit-reversed order as n/2 complex numbers (real-input packing),
// log2(n/2) radix-2 stages in place, two per pass (twiddles and window from host-built float64 -> float32
// tables), then unpacking, magnitude and dB of the n/2 + 1 bins.  n is a power of two, 64..4096 (6 KB of
// shared memory at 1024, 25 KB at 4096).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace afp {

struct SpectrumParams {
    const float* in;           // [n_streams][in_pitch]
    long long in_pitch;
    float* out;                // [n_streams][out_pitch], n/2 + 1 values per row
    long long out_pitch;
    const float* window;       // [n]    np.hanning(n)
    const float2* twiddle;     // [n/2]  exp(-2 pi i k / n)
    int n_streams, n, log2n;
};

constexpr int SPEC_THREADS = 256;

// one pad element per 32: the bit-reversed scatter of the load phase (consecutive threads -> addresses that
// differ in their HIGH bits, i.e. one bank) and the stride-4 accesses of the first pass spread over the banks
__host__ __device__ constexpr int spec_pad(int i) { return i + (i >> 5); }

__device__ __forceinline__ float2 spec_mul(float2 b, float2 w) {
    return make_float2(fmaf(b.x, w.x, -b.y * w.y), fmaf(b.x, w.y, b.y * w.x));
}

// The block is real, so it is transformed as ONE complex FFT of m = n/2 points on z[j] = xw[2j] + i xw[2j+1]
// and unpacked:  X[k] = (Z[k] + conj Z[m-k]) / 2  -  i W_n^k (Z[k] - conj Z[m-k]) / 2,  k = 0..m  (Z[m] = Z[0]).
// Two radix-2 stages per pass: a thread takes the four elements {i, i+h, i+2h, i+3h} (h = 2^st, i with
// bits st and st+1 clear), does the stage-st butterflies on (i, i+h), (i+2h, i+3h) and the stage-(st+1)
// butterflies on (i, i+2h), (i+h, i+3h) in registers: half the barriers and shared-memory round trips.
// Twiddles W_m^k = W_n^{2k} come from a shared-memory copy of every second table entry.
__global__ void __launch_bounds__(SPEC_THREADS) spectrum_db_kernel(const __grid_constant__ SpectrumParams P) {
    extern __shared__ float2 spec_smem[];
    const int n = P.n, m = n >> 1, log2m = P.log2n - 1, tid = threadIdx.x;
    float2* spec_x = spec_smem;                    // [spec_pad(m)]
    float2* tw = spec_smem + spec_pad(m);          // [m/2]   W_m^k
    for (int k = tid; k < (m >> 1); k += SPEC_THREADS) tw[k] = __ldg(P.twiddle + 2 * k);
    for (int s = blockIdx.x; s < P.n_streams; s += gridDim.x) {
        const float* x = P.in + (long long)s * P.in_pitch;
        const bool al8 = (reinterpret_cast<uintptr_t>(x) & 7u) == 0;
        for (int j = tid; j < m; j += SPEC_THREADS) {
            float2 v;
            if (al8) v = __ldg(reinterpret_cast<const float2*>(x) + j);
            else { v.x = __ldg(x + 2 * j); v.y = __ldg(x + 2 * j + 1); }
            const float2 w = __ldg(reinterpret_cast<const float2*>(P.window) + j);
            spec_x[spec_pad((int)(__brev((unsigned)j) >> (32 - log2m)))] = make_float2(v.x * w.x, v.y * w.y);
        }
        __syncthreads();
        int st = 0;
        if (log2m & 1) {                           // odd number of stages: one plain radix-2 stage first (all twiddles 1)
            for (int j = tid; j < (m >> 1); j += SPEC_THREADS) {
                const int p0 = spec_pad(2 * j), p1 = spec_pad(2 * j + 1);
                const float2 a = spec_x[p0], b = spec_x[p1];
                spec_x[p0] = make_float2(a.x + b.x, a.y + b.y);
                spec_x[p1] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
            st = 1;
        }
        for (; st < log2m; st += 2) {
            const int h = 1 << st;
            for (int j = tid; j < (m >> 2); j += SPEC_THREADS) {
                const int k = j & (h - 1);
                const int i0 = ((j >> st) << (st + 2)) + k;
                const float2 w1 = tw[(size_t)k * (m >> (st + 1))];            // stage st:     W_{2h}^k
                const float2 w2a = tw[(size_t)k * (m >> (st + 2))];           // stage st + 1: W_{4h}^k
                const float2 w2b = tw[(size_t)(k + h) * (m >> (st + 2))];     //               W_{4h}^{k+h}
                const int pa = spec_pad(i0), pb = spec_pad(i0 + h), pc = spec_pad(i0 + 2 * h), pd = spec_pad(i0 + 3 * h);
                const float2 a = spec_x[pa], b = spec_x[pb], c = spec_x[pc], d = spec_x[pd];
                const float2 tb = spec_mul(b, w1), td = spec_mul(d, w1);
                const float2 a1 = make_float2(a.x + tb.x, a.y + tb.y), b1 = make_float2(a.x - tb.x, a.y - tb.y);
                const float2 c1 = make_float2(c.x + td.x, c.y + td.y), d1 = make_float2(c.x - td.x, c.y - td.y);
                const float2 tc = spec_mul(c1, w2a), te = spec_mul(d1, w2b);
                spec_x[pa] = make_float2(a1.x + tc.x, a1.y + tc.y);
                spec_x[pc] = make_float2(a1.x - tc.x, a1.y - tc.y);
                spec_x[pb] = make_float2(b1.x + te.x, b1.y + te.y);
                spec_x[pd] = make_float2(b1.x - te.x, b1.y - te.y);
            }
            __syncthreads();
        }
        float* y = P.out + (long long)s * P.out_pitch;
        for (int k = tid; k <= m; k += SPEC_THREADS) {
            const float2 zk = spec_x[spec_pad(k == m ? 0 : k)], zr = spec_x[spec_pad(k == 0 ? 0 : m - k)];
            // e = (Z[k] + conj Z[m-k]) / 2,  o = (Z[k] - conj Z[m-k]) / 2,  X[k] = e - i W_n^k o   (W_n^m = -1)
            const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
            const float2 o = make_float2(0.5f * (zk.x - zr.x), 0.5f * (zk.y + zr.y));
            const float2 w = k == m ? make_float2(-1.f, 0.f) : __ldg(P.twiddle + k);
            const float2 t = spec_mul(o, w);                                   // W o;  -i (t.x + i t.y) = t.y - i t.x
            const float2 v = make_float2(e.x + t.y, e.y - t.x);
            // 20 log10(|v|) = 20 log10(2) * log2(|v|); MUFU.LG2 is good to ~1e-6 in the result, far inside 0.01 dB
            y[k] = 6.02059991327962f * __log2f(sqrtf(fmaf(v.x, v.x, v.y * v.y)) + 1e-12f);
        }
        __syncthreads();       // spec_x is rewritten by the next stream
    }
}


// ---- n = 1024 (the Waterfall GUI's default block, stream_process_GUI_Waterfall.py:115): one WARP per row, no CTA barrier.
// The 512 packed complex points are transformed as 8 x 8 x 8, decimation in frequency: a lane owns 16 points and does two
// 8-point DFTs in registers per pass; the first pass runs straight on the loaded (windowed) samples - lane l holds points
// l + 32 r, which are exactly the stride-64 groups of a = l and a = l + 32 - and only the two regroupings between the passes,
// the natural-order write of the last pass and the mirrored read of the unpacking go through the warp's own 4 KB of
// shared memory (__syncwarp only).  All pass twiddles are lane constants and live in registers for the whole launch
// (W_512^{a q} for the lane's two a, W_64^{a' q'} for its a'); window and unpacking twiddles come from shared-memory
// copies.  Shared-memory slot of point i: i ^ ((i >> 4) & 15) - 16 consecutive points of an aligned group stay a
// permutation of one 128-byte line (pass 1 stores, unpacking), the stride-8 groups of pass 3 and the two runs of pass 2
// (lanes 0-7 / 8-15 take blocks q and q + 2, which differ in bit 3 of i >> 4) fall into distinct 8-byte banks.
constexpr int SPW_N = 1024, SPW_M = 512, SPW_WARPS = 8;
constexpr int SPW_SMEM = SPW_WARPS * SPW_M * 8 + SPW_N * 4 + SPW_M * 8;   // buffers + window + unpacking twiddles

__device__ __forceinline__ int spw_slot(int i) { return i ^ ((i >> 4) & 15); }
__device__ __forceinline__ float2 spw_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 spw_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 spw_mi(float2 a) { return make_float2(a.y, -a.x); }          // times -i
// forward 8-point DFT in place, natural order in and out (exp(-2 pi i b q / 8))
__device__ __forceinline__ void spw_dft8(float2 (&x)[8]) {
    const float h = 0.70710678118654752f;
    const float2 a0 = spw_add(x[0], x[4]), a1 = spw_add(x[1], x[5]), a2 = spw_add(x[2], x[6]), a3 = spw_add(x[3], x[7]);
    const float2 b0 = spw_sub(x[0], x[4]), t1 = spw_sub(x[1], x[5]), t2 = spw_sub(x[2], x[6]), t3 = spw_sub(x[3], x[7]);
    const float2 b1 = make_float2((t1.x + t1.y) * h, (t1.y - t1.x) * h);       // times (1 - i) / sqrt 2
    const float2 b2 = spw_mi(t2);
    const float2 b3 = make_float2((t3.y - t3.x) * h, -(t3.x + t3.y) * h);      // times (-1 - i) / sqrt 2
    const float2 d0 = spw_add(a0, a2), d1 = spw_sub(a0, a2), d2 = spw_add(a1, a3), d3 = spw_mi(spw_sub(a1, a3));
    x[0] = spw_add(d0, d2); x[4] = spw_sub(d0, d2); x[2] = spw_add(d1, d3); x[6] = spw_sub(d1, d3);
    const float2 e0 = spw_add(b0, b2), e1 = spw_sub(b0, b2), e2 = spw_add(b1, b3), e3 = spw_mi(spw_sub(b1, b3));
    x[1] = spw_add(e0, e2); x[5] = spw_sub(e0, e2); x[3] = spw_add(e1, e3); x[7] = spw_sub(e1, e3);
}
// W_1024^k for k in 0..1023 from the half table (W^{k+512} = -W^k)
__device__ __forceinline__ float2 spw_tw(const float2* tab, int k) {
    const float2 w = __ldg(tab + (k & 511));
    return (k & 512) ? make_float2(-w.x, -w.y) : w;
}

__global__ void __launch_bounds__(32 * SPW_WARPS, 2) spectrum_db_warp_kernel(const __grid_constant__ SpectrumParams P) {
    extern __shared__ __align__(16) unsigned char spw_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float2* buf = reinterpret_cast<float2*>(spw_smem) + warp * SPW_M;
    float2* win2 = reinterpret_cast<float2*>(spw_smem + SPW_WARPS * SPW_M * 8);     // (window[2j], window[2j+1])
    float2* utw = win2 + SPW_M;                                                      // W_1024^k, k < 512
    for (int j = tid; j < SPW_M; j += 32 * SPW_WARPS) {
        win2[j] = __ldg(reinterpret_cast<const float2*>(P.window) + j);
        utw[j] = __ldg(P.twiddle + j);
    }
    // lane constants: pass 1 W_512^{a q} = W_1024^{2 a q} for a = lane, lane + 32; pass 2 W_64^{a' q'} = W_1024^{16 a' q'}, a' = lane & 7
    float2 tw1a[7], tw1b[7], tw2[7];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        tw1a[q - 1] = spw_tw(P.twiddle, 2 * lane * q);
        tw1b[q - 1] = spw_tw(P.twiddle, 2 * (lane + 32) * q);
        tw2[q - 1] = spw_tw(P.twiddle, 16 * (lane & 7) * q);
    }
    const int ap = lane & 7;
    const int q2lane = (((lane >> 3) & 1) << 1) + (lane >> 4);      // pass 2: block q = q2lane + 4 t
    __syncthreads();

    for (int row = blockIdx.x * SPW_WARPS + warp; row < P.n_streams; row += gridDim.x * SPW_WARPS) {
        const float* x = P.in + (long long)row * P.in_pitch;
        float2 z[16];
        if ((reinterpret_cast<uintptr_t>(x) & 7u) == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = __ldg(reinterpret_cast<const float2*>(x) + lane + 32 * r);
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) { z[r].x = __ldg(x + 2 * (lane + 32 * r)); z[r].y = __ldg(x + 2 * (lane + 32 * r) + 1); }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float2 w = win2[lane + 32 * r];
            z[r].x *= w.x; z[r].y *= w.y;
        }
        // ---- pass 1: points a + 64 b, a = lane (even r) and a = lane + 32 (odd r); result q of a goes to 64 q + a
        {
            float2 A[8], B[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) { A[b] = z[2 * b]; B[b] = z[2 * b + 1]; }
            spw_dft8(A);
            spw_dft8(B);
            buf[spw_slot(lane)] = A[0];
            buf[spw_slot(lane + 32)] = B[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                buf[spw_slot(64 * q + lane)] = spec_mul(A[q], tw1a[q - 1]);
                buf[spw_slot(64 * q + lane + 32)] = spec_mul(B[q], tw1b[q - 1]);
            }
        }
        __syncwarp();
        // ---- pass 2: inside block q, points a' + 8 b'; result q' goes to 8 q' + a' (in place)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int base = 64 * (q2lane + 4 * t) + ap;
            float2 C[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) C[b] = buf[spw_slot(base + 8 * b)];
            spw_dft8(C);
            buf[spw_slot(base)] = C[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) buf[spw_slot(base + 8 * q)] = spec_mul(C[q], tw2[q - 1]);
        }
        __syncwarp();
        // ---- pass 3: the 8 points of block g = 8 q + q'; result q'' is Z[q + 8 q' + 64 q''], written in natural order
        {
            float2 D0[8], D1[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) { D0[a] = buf[spw_slot(8 * lane + a)]; D1[a] = buf[spw_slot(8 * (lane + 32) + a)]; }
            __syncwarp();                                   // every point has been read before the first natural-order write
            spw_dft8(D0);
            spw_dft8(D1);
            const int k0 = (lane >> 3) + 8 * (lane & 7);    // g = lane: q = lane >> 3; g = lane + 32: q = (lane >> 3) + 4
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                buf[spw_slot(k0 + 64 * q)] = D0[q];
                buf[spw_slot(k0 + 4 + 64 * q)] = D1[q];
            }
        }
        __syncwarp();
        // ---- unpacking, magnitude, dB:  X[k] = (Z[k] + conj Z[m-k]) / 2 - i W_n^k (Z[k] - conj Z[m-k]) / 2
        float* y = P.out + (long long)row * P.out_pitch;
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const int k = lane + 32 * r;
            const float2 zk = buf[spw_slot(k)], zr = buf[spw_slot((SPW_M - k) & (SPW_M - 1))];
            const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
            const float2 o = make_float2(0.5f * (zk.x - zr.x), 0.5f * (zk.y + zr.y));
            const float2 t = spec_mul(o, utw[k]);
            const float2 v = make_float2(e.x + t.y, e.y - t.x);
            y[k] = 6.02059991327962f * __log2f(sqrtf(fmaf(v.x, v.x, v.y * v.y)) + 1e-12f);
        }
        if (lane == 0) {                                    // k = m: W_n^m = -1, Z[m] = Z[0]
            const float2 z0 = buf[0];
            y[SPW_M] = 6.02059991327962f * __log2f(fabsf(z0.x - z0.y) + 1e-12f);
        }
        __syncwarp();                                       // buf is rewritten by the next row
    }
}

}  // namespace afp
