// Write-pattern microbenchmark, second round: which SHAPE of store does a FIR epilogue need to reach
// coalesced-store bandwidth?  A thread owns RUN contiguous bytes of output (its T*U samples); groups of
// G lanes exchange data (shuffles / shared memory in the real kernel, free here) so that one STG.128
// instruction writes G*16 contiguous bytes per group.  Also: the bulk-copy engine fed from shared memory
// with per-thread and per-warp copies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern2 store_pattern2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// RUN bytes per thread, G lanes cooperate: a group of G lanes owns G*RUN contiguous bytes and writes it
// in RUN/16 instructions, each covering G*16 contiguous bytes.
template <int RUN, int G>
__global__ void st_group(float4* out, size_t n4, int extra_alu) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, grp = lane / G, lg = lane % G;
    constexpr int GR = RUN / 16;                 // granules per thread
    float4 v = make_float4(lane, 1.f, 2.f, 3.f);
    for (size_t base = warp * 32 * GR; base + 32 * GR <= n4; base += nwarp * 32 * GR) {
        float4* p = out + base + (size_t)grp * G * GR + lg;
#pragma unroll
        for (int m = 0; m < GR; ++m) p[m * G] = v;
    }
}

// per-thread bulk copy of RUN bytes from a shared-memory staging row
template <int RUN>
__global__ void st_bulk_thread(float4* out, size_t n4) {
    extern __shared__ float4 sm[];
    constexpr int GR = RUN / 16;
    float4* stage = sm + threadIdx.x * (GR + 1);
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(t, 1.f, 2.f, 3.f);
    for (size_t base = t * GR; base + GR <= n4; base += nt * GR) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
        for (int m = 0; m < GR; ++m) stage[m] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(out + base), "r"((unsigned)__cvta_generic_to_shared(stage)), "r"(RUN) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// per-warp bulk copy: 32 lanes stage 32*RUN contiguous bytes (linear, lane-interleaved writes: conflict
// free), lane 0 issues one copy
template <int RUN>
__global__ void st_bulk_warp(float4* out, size_t n4) {
    extern __shared__ float4 sm[];
    constexpr int GR = RUN / 16;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* stage = sm + wib * 32 * GR;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
    float4 v = make_float4(lane, 1.f, 2.f, 3.f);
    for (size_t base = warp * 32 * GR; base + 32 * GR <= n4; base += nwarp * 32 * GR) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int m = 0; m < GR; ++m) stage[m * 32 + lane] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(out + base), "r"((unsigned)__cvta_generic_to_shared(stage)), "r"(32 * RUN) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <class F> static float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, ms;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    size_t bytes = 1ull << 30, n4 = bytes / 16;
    float4* a; CK(cudaMalloc(&a, bytes));
    const int blk = 256;
    for (int cta = 2; cta <= 8; cta *= 2) {
        const int grid = 148 * cta;
        printf("-- %d CTAs/SM x 256 threads, 1 GiB written\n", cta);
#define RUNG(RUN, G) { float ms = time_it([&] { st_group<RUN, G><<<grid, blk>>>(a, n4, 0); }); CK(cudaGetLastError()); \
        printf("run %3d B/thread, %2d lanes x 16 B contiguous per instr (%4d B): %.3f ms %5.0f GB/s\n", RUN, G, G * 16, ms, bytes / ms / 1e6); }
        RUNG(64, 1) RUNG(64, 2) RUNG(64, 4) RUNG(64, 8) RUNG(64, 32)
        RUNG(128, 1) RUNG(128, 2) RUNG(128, 4) RUNG(128, 8) RUNG(128, 32)
        RUNG(32, 1) RUNG(32, 2)
        { CK(cudaFuncSetAttribute(st_bulk_thread<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, blk * 9 * 16));
          float ms = time_it([&] { st_bulk_thread<128><<<grid, blk, blk * 9 * 16>>>(a, n4); }); CK(cudaGetLastError());
          printf("bulk per thread 128 B: %.3f ms %5.0f GB/s\n", ms, bytes / ms / 1e6); }
        { float ms = time_it([&] { st_bulk_thread<64><<<grid, blk, blk * 5 * 16>>>(a, n4); }); CK(cudaGetLastError());
          printf("bulk per thread  64 B: %.3f ms %5.0f GB/s\n", ms, bytes / ms / 1e6); }
        { float ms = time_it([&] { st_bulk_warp<128><<<grid, blk, blk * 128>>>(a, n4); }); CK(cudaGetLastError());
          printf("bulk per warp  4096 B: %.3f ms %5.0f GB/s\n", ms, bytes / ms / 1e6); }
        { float ms = time_it([&] { st_bulk_warp<64><<<grid, blk, blk * 64>>>(a, n4); }); CK(cudaGetLastError());
          printf("bulk per warp  2048 B: %.3f ms %5.0f GB/s\n", ms, bytes / ms / 1e6); }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
