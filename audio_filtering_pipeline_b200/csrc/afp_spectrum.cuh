// Waterfall row of one block per stream: out[k] = 20*log10(|rfft(x * hanning(n))[k]| + 1e-12), k = 0..n/2
// (stream_process_GUI_Waterfall.py:471-476).  One CTA per stream at a time (persistent over streams): the
// windowed block goes to shared memory in bit-reversed order as complex numbers, log2(n) radix-2 stages in
// place (twiddles and window from host-built float64 -> float32 tables), then magnitude and dB of the first
// n/2 + 1 bins.  n is a power of two, 64..4096 (16 KB of shared memory at 2048, 32 KB at 4096).
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

__global__ void __launch_bounds__(SPEC_THREADS) spectrum_db_kernel(const __grid_constant__ SpectrumParams P) {
    extern __shared__ float2 spec_x[];
    const int n = P.n, tid = threadIdx.x;
    for (int s = blockIdx.x; s < P.n_streams; s += gridDim.x) {
        const float* x = P.in + (long long)s * P.in_pitch;
        for (int i = tid; i < n; i += SPEC_THREADS)
            spec_x[__brev((unsigned)i) >> (32 - P.log2n)] = make_float2(__ldg(x + i) * __ldg(P.window + i), 0.f);
        __syncthreads();
        for (int st = 0; st < P.log2n; ++st) {
            const int half = 1 << st;
            for (int j = tid; j < (n >> 1); j += SPEC_THREADS) {
                const int k = j & (half - 1);
                const int i0 = ((j >> st) << (st + 1)) + k, i1 = i0 + half;
                const float2 w = __ldg(P.twiddle + (size_t)k * (n >> (st + 1)));
                const float2 a = spec_x[i0], b = spec_x[i1];
                const float2 t = make_float2(fmaf(b.x, w.x, -b.y * w.y), fmaf(b.x, w.y, b.y * w.x));
                spec_x[i0] = make_float2(a.x + t.x, a.y + t.y);
                spec_x[i1] = make_float2(a.x - t.x, a.y - t.y);
            }
            __syncthreads();
        }
        float* y = P.out + (long long)s * P.out_pitch;
        for (int k = tid; k <= (n >> 1); k += SPEC_THREADS) {
            const float2 v = spec_x[k];
            y[k] = 20.f * log10f(sqrtf(fmaf(v.x, v.x, v.y * v.y)) + 1e-12f);
        }
        __syncthreads();       // spec_x is rewritten by the next stream
    }
}

}  // namespace afp
