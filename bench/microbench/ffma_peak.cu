// FP32 FMA issue-rate microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json has no FP32 entry, so the FMA roof the FIR kernels are
// judged against has to be measured on the box.  Each variant runs a long unrolled stream of
// independent FMAs (16 or 32 accumulators per thread) in the operand form one of the FIR inner
// loops would use, and reports FMA/clk/SM (from clock64) and TFLOP/s (from CUDA events).
//
//   rrr    FFMA  acc, x(reg), h(reg), acc          - taps held in ordinary registers
//   cst    FFMA  acc, x(reg), c[bank][imm], acc    - taps as constant-bank operands
//   ur     FFMA  acc, x(reg), h(uniform reg), acc  - taps loaded through the uniform datapath
//   x2     FFMA2 (fma.rn.f32x2) on 64-bit register pairs - two FMAs per lane per instruction
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ffma_peak ffma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NACC = 16;
constexpr int NTAP = 16;
__constant__ float c_taps[1024];

__global__ void __launch_bounds__(256) k_rrr(float* out, const float* in, int iters, long long* cyc) {
    float acc[NACC], h[NTAP];
    float x = in[threadIdx.x];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = x + i;
#pragma unroll
    for (int j = 0; j < NTAP; ++j) h[j] = in[j + 32 + (threadIdx.x & 1)];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NTAP; ++j)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[(i + 1) % NACC] , h[j], acc[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// x operand distinct from the accumulators: window registers w[], like the FIR loop.
__global__ void __launch_bounds__(256) k_rrr_win(float* out, const float* in, int iters, long long* cyc) {
    float acc[NACC], h[NTAP], w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = in[threadIdx.x + i];
#pragma unroll
    for (int j = 0; j < NTAP; ++j) h[j] = in[j + 32 + (threadIdx.x & 1)];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NTAP; ++j)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(w[i + NTAP - j], h[j], acc[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) k_cst(float* out, const float* in, int iters, long long* cyc) {
    float acc[NACC], w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = in[threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NTAP; ++j)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(w[i + NTAP - j], c_taps[j], acc[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// taps indexed by the (uniform) loop counter -> ULDC + FFMA with a uniform-register operand
__global__ void __launch_bounds__(256) k_ur(float* out, const float* in, int iters, long long* cyc) {
    float acc[NACC], w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = in[threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float* hp = c_taps + (it & 31) * NTAP;
#pragma unroll
        for (int j = 0; j < NTAP; ++j) {
            float hj = hp[j];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(w[i + NTAP - j], hj, acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

// two interleaved sequences per thread: acc pair, window pair, tap pair (h,h)
__global__ void __launch_bounds__(256) k_x2(float* out, const float* in, int iters, long long* cyc) {
    float2 acc[NACC], h[NTAP], w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 64]);
#pragma unroll
    for (int j = 0; j < NTAP; ++j) h[j] = make_float2(in[j + 32], in[j + 32]);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NTAP; ++j)
#pragma unroll
            for (int i = 0; i < NACC; ++i) ffma2(acc[i], w[i + NTAP - j], h[j]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 with lane-dependent tap pairs (ordinary 64-bit register operands)
__global__ void __launch_bounds__(256) k_x2_rr(float* out, const float* in, int iters, long long* cyc) {
    float2 acc[8], h[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 64]);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = make_float2(in[j + 32 + (threadIdx.x & 1)], in[j + 40 + (threadIdx.x & 1)]);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(acc[i], w[i + 8 - j], h[j]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// taps in shared memory at a warp-uniform address, FIR-like: 16 taps per chunk, chunk index runs
__global__ void __launch_bounds__(256) k_smem_taps(float* out, const float* in, int iters, long long* cyc) {
    __shared__ float hs[1024];
    __shared__ float xs[256 * 20 + 64];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) hs[i] = in[i];
    for (int i = threadIdx.x; i < 256 * 20 + 64; i += blockDim.x) xs[i] = in[i & 1023];
    __syncthreads();
    float acc[NACC], w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = xs[threadIdx.x * 16 + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float4* hp = reinterpret_cast<const float4*>(hs + (it & 63) * NTAP);
        const float4* xp = reinterpret_cast<const float4*>(xs + ((threadIdx.x + it) & 255) * 20);
        float hh[NTAP];
#pragma unroll
        for (int j = 0; j < NTAP / 4; ++j) { float4 v = hp[j]; hh[4*j]=v.x; hh[4*j+1]=v.y; hh[4*j+2]=v.z; hh[4*j+3]=v.w; }
#pragma unroll
        for (int j = 0; j < NTAP / 4; ++j) { float4 v = xp[j]; w[4*j]=v.x; w[4*j+1]=v.y; w[4*j+2]=v.z; w[4*j+3]=v.w; }
#pragma unroll
        for (int j = 0; j < NTAP; ++j)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(w[i + NTAP - j], hh[j], acc[i]);
#pragma unroll
        for (int i = 0; i < NACC; ++i) w[i + NACC] = w[i];
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 with 8 tap pairs only (fewer registers), to see whether occupancy matters
__global__ void __launch_bounds__(256) k_x2_small(float* out, const float* in, int iters, long long* cyc) {
    float2 acc[8], h[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 64]);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = make_float2(in[j + 32], in[j + 32]);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(acc[i], w[i + 8 - j], h[j]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 in the form  acc(y[r],y[r+1]) += x[n](scalar broadcast, ordinary register) * (h[k],h[k+1])
// (tap PAIR in a 64-bit uniform register): SASS "FFMA2 R, R.F32, UR.F32x2.HI_LO, R.F32x2.HI_LO".
// One row per thread, no row interleave: the inner loop of the single-row FIR kernel for U = 1, T = 16.
__global__ void __launch_bounds__(256) k_x2_bc(float* out, const float* in, int iters, long long* cyc) {
    unsigned long long acc[NACC / 2];
    float w[2 * NACC];
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) acc[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) w[i] = in[threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float* ta = c_taps + (it & 15) * NTAP;          // even-start pairs
        const float* tb = c_taps + 512 + (it & 15) * NTAP;    // odd-start pairs (table shifted by one)
#pragma unroll
        for (int m = 1; m < 2 * NACC; ++m) {
#pragma unroll
            for (int r = 0; r < NACC; r += 2) {
                const int j = NACC + r - m;
                if (j < 0 || j >= NACC) continue;
                const float* tp = (m & 1) ? tb + (j - 1) : ta + j;
                const unsigned long long h = *reinterpret_cast<const unsigned long long*>(tp);
                const unsigned long long xx = ((unsigned long long)__float_as_uint(w[m]) << 32) | __float_as_uint(w[m]);
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r / 2]) : "l"(xx), "l"(h));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(float*, const float*, int, long long*);

static void run(const char* name, kern_t k, int fma_per_iter_per_thread, int nsm, int ctas_per_sm, int threads) {
    int grid = nsm * ctas_per_sm;
    float *out, *in; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * grid * threads));
    CK(cudaMalloc(&in, sizeof(float) * 4096));
    CK(cudaMemset(in, 0, sizeof(float) * 4096));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    int iters = 4000;
    k<<<grid, threads>>>(out, in, 10, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<grid, threads>>>(out, in, iters, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long* hc = (long long*)malloc(sizeof(long long) * grid);
    CK(cudaMemcpy(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double cmax = 0; for (int i = 0; i < grid; ++i) if (hc[i] > cmax) cmax = (double)hc[i];
    double fma = (double)fma_per_iter_per_thread * iters * threads * grid;
    double per_sm_clk = (double)fma_per_iter_per_thread * iters * threads * ctas_per_sm / cmax;
    printf("{\"variant\":\"%s\",\"ctas_per_sm\":%d,\"threads\":%d,\"ms\":%.4f,\"tflops\":%.2f,"
           "\"fma_per_clk_sm\":%.1f,\"eff_mhz\":%.0f}\n",
           name, ctas_per_sm, threads, best, 2.0 * fma / (best * 1e-3) / 1e12, per_sm_clk,
           cmax / (best * 1e-3) / 1e6);
    free(hc); cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"smem_per_sm\":%zu,\"regs_per_sm\":%d}\n",
           p.name, p.multiProcessorCount, p.clockRate, p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor);
    float hz[1024]; for (int i = 0; i < 1024; ++i) hz[i] = 1e-3f * i;
    CK(cudaMemcpyToSymbol(c_taps, hz, sizeof(hz)));
    int nsm = p.multiProcessorCount;
    for (int c = 1; c <= 4; c *= 2) {
        run("rrr", k_rrr, NACC * NTAP, nsm, c, 256);
        run("rrr_win", k_rrr_win, NACC * NTAP, nsm, c, 256);
        run("cst", k_cst, NACC * NTAP, nsm, c, 256);
        run("ur", k_ur, NACC * NTAP, nsm, c, 256);
        run("x2", k_x2, 2 * NACC * NTAP, nsm, c, 256);
        run("x2_small", k_x2_small, 2 * 4 * 8 * 8, nsm, c, 256);
        run("x2_rr", k_x2_rr, 2 * 4 * 8 * 8, nsm, c, 256);
        run("x2_bc", k_x2_bc, NACC * NTAP, nsm, c, 256);
        run("smem_taps", k_smem_taps, NACC * NTAP, nsm, c, 256);
    }
    return 0;
}
