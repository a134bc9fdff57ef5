// Write-bandwidth microbenchmark: does the per-thread-contiguous store pattern of the FIR epilogue
// (each lane writes 64 contiguous bytes as 4 x STG.128, lane stride 64 B) cost HBM/L2 throughput
// against fully coalesced STG.128 (a warp instruction covers 512 contiguous bytes)?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void st_coalesced(float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(i, 1.f, 2.f, 3.f);
    for (; i + 3 * stride < n4; i += 4 * stride) { out[i] = v; out[i + stride] = v; out[i + 2 * stride] = v; out[i + 3 * stride] = v; }
}
__global__ void st_thread64(float4* out, size_t n4) {   // lane writes 4 consecutive float4
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(t, 1.f, 2.f, 3.f);
    for (size_t i = t * 4; i + 3 < n4; i += 4 * stride) { out[i] = v; out[i + 1] = v; out[i + 2] = v; out[i + 3] = v; }
}
__global__ void st_thread64_rd(float4* out, const float4* in, size_t n4) {   // + 1/4 read traffic
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = t * 4; i + 3 < n4; i += 4 * stride) { float4 v = in[i / 4]; out[i] = v; out[i + 1] = v; out[i + 2] = v; out[i + 3] = v; }
}
__global__ void st_coalesced_rd(float4* out, const float4* in, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) { float4 v = in[i / 4]; out[i] = v; out[i + stride] = v; out[i + 2 * stride] = v; out[i + 3 * stride] = v; }
}
__global__ void copy_k(float4* out, const float4* in, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) out[i] = in[i];
}
int main() {
    size_t bytes = 1ull << 30, n4 = bytes / 16;
    float4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 0, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 8, blk = 256; float ms;
    for (int k = 0; k < 5; ++k) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (k == 0) st_coalesced<<<grid, blk>>>(a, n4);
            if (k == 1) st_thread64<<<grid, blk>>>(a, n4);
            if (k == 2) st_coalesced_rd<<<grid, blk>>>(a, b, n4);
            if (k == 3) st_thread64_rd<<<grid, blk>>>(a, b, n4);
            if (k == 4) copy_k<<<grid, blk>>>(a, b, n4);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const char* nm[] = {"write coalesced", "write thread-64B", "write coalesced + 1/4 read", "write thread-64B + 1/4 read", "copy"};
        double moved = k == 4 ? 2.0 * bytes : (k >= 2 ? 1.25 * bytes : 1.0 * bytes);
        printf("%-30s %.3f ms  %.0f GB/s\n", nm[k], ms, moved / ms / 1e6);
    }
    return 0;
}
