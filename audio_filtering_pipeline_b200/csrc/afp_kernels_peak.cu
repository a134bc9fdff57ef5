// Measured tensor-pipe roof for the roofline of the split-FP16 FIR kernel (SURVEY.md 8d asks for measured peaks;
// MEASURED_PEAKS.json holds a cuBLAS bf16 figure, this is the same pipe driven by a bare tcgen05.mma loop in the
// un-swizzled K-major layout the FIR kernel uses).  M = 128, K = 16 (kind::f16), N = n_cols: N = 256 is the pipe's
// own rate (128 clk per MMA), N = 64 the FIR kernel's instruction shape, where fetching the 4 KB A tile + 2 KB B tile
// from shared memory at 128 B/clk (48 clk) is slower than the pipe (32 clk) - bench/microbench/umma_peak.cu has the
// whole table.
#include <vector>
#include "../../include/afp.h"
#include "afp_fir_tc2.cuh"

namespace {

constexpr int PK_A_BYTES = 64 * 1024, PK_B_BYTES = 64 * 1024;

__global__ void __launch_bounds__(128, 1) tensor_peak_kernel(uint32_t idesc, uint32_t b_lbo, int n_mma, long long* cycles) {
    using namespace afp;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + PK_A_BYTES + PK_B_BYTES);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (PK_A_BYTES + PK_B_BYTES) / 4; i += blockDim.x) {   // two fp16 of 0.06..0.12 per word
        const uint32_t r = (uint32_t)i * 2654435761u;
        reinterpret_cast<uint32_t*>(smem)[i] = (0x2C00u | (r & 0x3FFu)) | ((0x2C00u | ((r >> 10) & 0x3FFu)) << 16);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
        const uint64_t a_d = umma_desc(smem_u32(smem), 2048, 128), b_d = umma_desc(smem_u32(smem + PK_A_BYTES), b_lbo, 128);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < n_mma; ++i)
                umma_f16(tmem_base, (uint32_t)a_d, (uint32_t)(a_d >> 32), (uint32_t)b_d, (uint32_t)(b_d >> 32), idesc, i == 0 ? 0u : 1u);
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

}  // namespace

extern "C" int afp_measure_tensor_peak(int32_t device, int32_t n_cols, double* tflops, double* sm_mhz, double* clk_per_mma) {
    if (!tflops || !sm_mhz || !clk_per_mma) return AFP_ERR_INVALID;
    if (n_cols < 16 || n_cols > 256 || n_cols % 16 != 0) return AFP_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return AFP_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AFP_ERR_CUDA;
    if (prop.major != 10) return AFP_ERR_INVALID;
    const int grid = prop.multiProcessorCount, n_mma = 8192, smem = PK_A_BYTES + PK_B_BYTES + 64;
    if (cudaFuncSetAttribute(tensor_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AFP_ERR_CUDA;
    long long* cyc = nullptr;
    if (cudaMalloc(&cyc, sizeof(long long) * grid) != cudaSuccess) return AFP_ERR_NOMEM;
    const uint32_t idesc = afp::tc2_idesc(n_cols), b_lbo = (uint32_t)n_cols * 16u;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    tensor_peak_kernel<<<grid, 128, smem>>>(idesc, b_lbo, 256, cyc);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        tensor_peak_kernel<<<grid, 128, smem>>>(idesc, b_lbo, n_mma, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const cudaError_t e = cudaGetLastError();
    std::vector<long long> hc(grid, 0);
    cudaMemcpy(hc.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long v : hc) cmax = v > cmax ? v : cmax;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(cyc);
    if (e != cudaSuccess) return AFP_ERR_CUDA;
    *tflops = 2.0 * 128 * n_cols * 16 * (double)n_mma * grid / (best * 1e-3) / 1e12;
    *sm_mhz = (double)cmax / (best * 1e-3) / 1e6;
    *clk_per_mma = (double)cmax / n_mma;
    return AFP_OK;
}
