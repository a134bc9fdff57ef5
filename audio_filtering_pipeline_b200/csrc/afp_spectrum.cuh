// Waterfall row of one block per stream: out[k] = 20*log10(|rfft(x * hanning(n))[k]| + 1e-12), k = 0..n/2
// (stream_process_GUI_Waterfall.py:471-476).  One CTA per stream at a time (persistent over streams): the
// windowed block goes to shared memory in bit-reversed order as complex numbers, log2(n) radix-2 stages in
// place, two per pass (twiddles and window from host-built float64 -> float32 tables), then magnitude and dB
// of the first n/2 + 1 bins.  n is a power of two, 64..4096 (12.3 KB of shared memory at 1024, 49 KB at 4096).
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

// Two radix-2 stages per pass: a thread takes the four elements {i, i+h, i+2h, i+3h} (h = 2^st, i with
// bits st and st+1 clear), does the stage-st butterflies on (i, i+h), (i+2h, i+3h) and the stage-(st+1)
// butterflies on (i, i+2h), (i+h, i+3h) in registers: half the barriers and shared-memory round trips.
// Twiddles come from a shared-memory copy of the table.
__global__ void __launch_bounds__(SPEC_THREADS) spectrum_db_kernel(const __grid_constant__ SpectrumParams P) {
    extern __shared__ float2 spec_smem[];
    const int n = P.n, tid = threadIdx.x;
    float2* spec_x = spec_smem;                    // [spec_pad(n)]
    float2* tw = spec_smem + spec_pad(n);          // [n/2]
    for (int k = tid; k < (n >> 1); k += SPEC_THREADS) tw[k] = __ldg(P.twiddle + k);
    for (int s = blockIdx.x; s < P.n_streams; s += gridDim.x) {
        const float* x = P.in + (long long)s * P.in_pitch;
        for (int i = tid; i < n; i += SPEC_THREADS)
            spec_x[spec_pad((int)(__brev((unsigned)i) >> (32 - P.log2n)))] = make_float2(__ldg(x + i) * __ldg(P.window + i), 0.f);
        __syncthreads();
        int st = 0;
        if (P.log2n & 1) {                         // odd number of stages: one plain radix-2 stage first (all twiddles 1)
            for (int j = tid; j < (n >> 1); j += SPEC_THREADS) {
                const int p0 = spec_pad(2 * j), p1 = spec_pad(2 * j + 1);
                const float2 a = spec_x[p0], b = spec_x[p1];
                spec_x[p0] = make_float2(a.x + b.x, a.y + b.y);
                spec_x[p1] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
            st = 1;
        }
        for (; st < P.log2n; st += 2) {
            const int h = 1 << st;
            for (int j = tid; j < (n >> 2); j += SPEC_THREADS) {
                const int k = j & (h - 1);
                const int i0 = ((j >> st) << (st + 2)) + k;
                const float2 w1 = tw[(size_t)k * (n >> (st + 1))];            // stage st:     W_{2h}^k
                const float2 w2a = tw[(size_t)k * (n >> (st + 2))];           // stage st + 1: W_{4h}^k
                const float2 w2b = tw[(size_t)(k + h) * (n >> (st + 2))];     //               W_{4h}^{k+h}
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
        for (int k = tid; k <= (n >> 1); k += SPEC_THREADS) {
            const float2 v = spec_x[spec_pad(k)];
            // 20 log10(m) = 20 log10(2) * log2(m); MUFU.LG2 is good to ~1e-6 in the result, far inside 0.01 dB
            y[k] = 6.02059991327962f * __log2f(sqrtf(fmaf(v.x, v.x, v.y * v.y)) + 1e-12f);
        }
        __syncthreads();       // spec_x is rewritten by the next stream
    }
}

}  // namespace afp
