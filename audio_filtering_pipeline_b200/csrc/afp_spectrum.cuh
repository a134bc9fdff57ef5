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

}  // namespace afp
