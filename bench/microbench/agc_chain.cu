// Dependent-chain cost of the AGC smoother step  s' = max(cA*s + aA*d, cR*s + aR*d)  in several
// instruction forms, one warp per CTA (the recurrence phase of afp_agc.cuh runs on one warp per CTA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o agc_chain agc_chain.cu && ./agc_chain
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int V>
__global__ void chain(const float* __restrict__ d, float* out, long long* cyc, int n, float cA, float aA, float cR, float aR) {
    __shared__ float row[32 * 257];
    for (int i = threadIdx.x; i < 32 * 257; i += 32) row[i] = d[i % 1024];
    __syncwarp();
    const float* r = row + threadIdx.x * 257;
    float s = r[0];
    const unsigned long long cc = pack(cA, cR), aa = pack(aA, aR);
    const long long t0 = clock64();
    for (int rep = 0; rep < n; ++rep) {
#pragma unroll 16
        for (int i = 0; i < 256; ++i) {
            const float x = r[i];
            if (V == 0) {            // packed: FFMA2 (products) + FFMA2 + FMNMX
                unsigned long long t;
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(pack(x, x)), "l"(aa), "l"(0ULL));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(pack(s, s)), "l"(cc));
                float2 b = *reinterpret_cast<float2*>(&t);
                s = fmaxf(b.x, b.y);
            } else if (V == 1) {     // scalar: 2 FMUL off chain, 2 FFMA, FMNMX
                s = fmaxf(fmaf(cA, s, aA * x), fmaf(cR, s, aR * x));
            } else if (V == 2) {     // u = s - d;  s' = d + max(cA*u, cR*u): FADD, 2 FMUL, FMNMX, FADD
                const float u = s - x;
                s = x + fmaxf(cA * u, cR * u);
            } else if (V == 3) {     // two-step look-ahead: 4 affine maps of s, max tree
                // (uses pairs of samples; the intermediate value is computed off the chain)
                if (i & 1) continue;
                const float y = r[i + 1];
                const float pA = aA * x, pR = aR * x, qA = aA * y, qR = aR * y;
                const float mAA = fmaf(cA * cA, s, fmaf(cA, pA, qA)), mRA = fmaf(cA * cR, s, fmaf(cA, pR, qA));
                const float mAR = fmaf(cR * cA, s, fmaf(cR, pA, qR)), mRR = fmaf(cR * cR, s, fmaf(cR, pR, qR));
                s = fmaxf(fmaxf(mAA, mRA), fmaxf(mAR, mRR));
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* d; float* out; long long* cyc;
    cudaMalloc(&d, 1024 * 4); cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.4f * ((i * 37 % 101) / 101.f);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    const int n = 64;
    const char* names[] = {"packed FFMA2+FFMA2+FMNMX", "scalar 2xFFMA+FMNMX", "FADD,2xFMUL,FMNMX,FADD", "two-step look-ahead (per 2 steps)"};
    for (int v = 0; v < 4; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            if (v == 0) chain<0><<<148, 32>>>(d, out, cyc, n, 0.82f, 0.18f, 0.98f, 0.02f);
            if (v == 1) chain<1><<<148, 32>>>(d, out, cyc, n, 0.82f, 0.18f, 0.98f, 0.02f);
            if (v == 2) chain<2><<<148, 32>>>(d, out, cyc, n, 0.82f, 0.18f, 0.98f, 0.02f);
            if (v == 3) chain<3><<<148, 32>>>(d, out, cyc, n, 0.82f, 0.18f, 0.98f, 0.02f);
            cudaDeviceSynchronize();
        }
        long long c[148];
        cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
        printf("%-36s %.2f cycles per step\n", names[v], (double)c[0] / (n * 256.0));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
