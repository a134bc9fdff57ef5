// tcgen05.mma issue-rate probe: cycles per MMA for kind::f16 (K = 16) and kind::tf32 (K = 8), M = 128, both
// operands from shared memory in the un-swizzled K-major core-matrix layout the FIR kernels use
// (afp_fir_tc.cuh / afp_fir_tc2.cuh).  It answers, with numbers measured on the box:
//   * the tensor-pipe roof for the roofline (SURVEY.md 8d asks for a measured one): FLOP/s of a plain
//     128 x N x K loop at N = 256, per kind, burst and over a seconds-long run;
//   * what an MMA of the FIR's shapes costs (N = 64 / 128, A descriptors that move by one 16-byte row per
//     chunk, B descriptors that walk a Hankel core array in 128-byte steps), i.e. whether the kernel is
//     bound by the tensor pipe or by operand fetch from shared memory;
//   * how many issuing warps it takes to keep the pipe busy.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench/microbench/umma_peak bench/microbench/umma_peak.cu
// Run:    bench/microbench/umma_peak            (prints one JSON line per configuration)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Params {
    uint32_t idesc;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;     // bytes
    uint32_t a_step, b_step;                 // descriptor start-address increment per MMA (16-byte units)
    uint32_t a_mod, b_mod;                   // ... modulo this many MMAs (keeps the addresses inside the buffers)
    int n_mma;                               // MMAs per issuing warp
    int issuers;                             // warps that issue (each into its own accumulator columns)
    int acc_cols;                            // TMEM columns per issuer
    int commit_every;                        // tcgen05.commit after this many MMAs (0 = only at the end)
    long long* cycles;                       // [grid]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                     uint32_t accumulate) {
    if (KIND == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int A_BYTES = 128 * 1024, B_BYTES = 64 * 1024;

template <int KIND>
__global__ void __launch_bounds__(128, 1) umma_loop(const Params P) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* a_buf = smem;
    unsigned char* b_buf = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    // operands: small finite values of the kind's element type (fp16 1.0 = 0x3C00; any fp32 pattern is a tf32)
    for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        const uint32_t r = (uint32_t)i * 2654435761u;
        uint32_t v;
        if (KIND == 0) v = 0x3C000000u | (r & 0x007FE000u);                  // ~0.0078..0.0156
        else v = (0x2C00u | (r & 0x3FFu)) | ((0x2C00u | ((r >> 10) & 0x3FFu)) << 16);   // two fp16 of ~0.06..0.12
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    long long t0 = 0, t1 = 0;
    if (warp_u < P.issuers) {
        const uint64_t a_d = umma_desc(smem_u32(a_buf), P.a_lbo, P.a_sbo), b_d = umma_desc(smem_u32(b_buf), P.b_lbo, P.b_sbo);
        const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32);
        const uint32_t acc = tmem_base + (uint32_t)(warp_u * P.acc_cols);
        t0 = clock64();
        uint32_t parity = 0;
        if (elect_one()) {
            uint32_t ia = 0, ib = 0;
            int since = 0;
            for (int i = 0; i < P.n_mma; ++i) {
                umma<KIND>(acc, (uint32_t)a_d + ia * P.a_step, a_hi, (uint32_t)b_d + ib * P.b_step, b_hi, P.idesc, i == 0 ? 0u : 1u);
                if (++ia == P.a_mod) ia = 0;
                if (++ib == P.b_mod) ib = 0;
                if (P.commit_every && ++since == P.commit_every) {   // what a ring of operand slots does: free the slot
                    since = 0;
                    umma_commit(&bars[4 + (warp_u & 3)]);             // nobody waits on these
                }
            }
            umma_commit(&bars[warp_u]);
        }
        __syncwarp();
        mbar_wait(&bars[warp_u], parity);
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) P.cycles[blockIdx.x] = t1 - t0;
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

struct Case { const char* name; int kind, n; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, a_mod, b_mod; int issuers, commit_every; };

int main(int argc, char** argv) {
    const double sustain_s = argc > 1 ? atof(argv[1]) : 2.0;
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int grid = prop.multiProcessorCount;
    const int smem = A_BYTES + B_BYTES + 256;
    CK(cudaFuncSetAttribute(umma_loop<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(umma_loop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * grid));
    // FIR-like A operand: 16 stream tiles of 40 rows (R = 64: 8 column chunks of 8 fp16), lbo = rows*16, sbo = tile pitch
    const uint32_t rows = 40, a_lbo = rows * 16, a_pitch = 8 * a_lbo;      // 5120 B per stream tile, 80 KB for 16
    std::vector<Case> cases = {
        // plain GEMM-like loops: contiguous core matrices (lbo = 128 rows * 16 B, sbo = 128 B)
        {"f16 N=256 plain", 1, 256, 2048, 128, 4096, 128, 0, 0, 1, 1, 1, 0},
        {"f16 N=128 plain", 1, 128, 2048, 128, 2048, 128, 0, 0, 1, 1, 1, 0},
        {"f16 N=64 plain", 1, 64, 2048, 128, 1024, 128, 0, 0, 1, 1, 1, 0},
        {"f16 N=32 plain", 1, 32, 2048, 128, 512, 128, 0, 0, 1, 1, 1, 0},
        {"tf32 N=256 plain", 0, 256, 2048, 128, 4096, 128, 0, 0, 1, 1, 1, 0},
        {"tf32 N=128 plain", 0, 128, 2048, 128, 2048, 128, 0, 0, 1, 1, 1, 0},
        {"tf32 N=64 plain", 0, 64, 2048, 128, 1024, 128, 0, 0, 1, 1, 1, 0},
        // two and four issuing warps on separate accumulators
        {"f16 N=128 plain x2 issuers", 1, 128, 2048, 128, 2048, 128, 0, 0, 1, 1, 2, 0},
        {"f16 N=64 plain x4 issuers", 1, 64, 2048, 128, 1024, 128, 0, 0, 1, 1, 4, 0},
        // FIR shapes: A = 16 stream tiles (sbo = tile pitch), start moves one row (16 B) per MMA; B = Hankel cores (lbo = sbo = 128 B)
        {"f16 N=64 FIR (A rows shifted, B Hankel)", 1, 64, a_lbo, a_pitch, 128, 128, 1, 8, 24, 64, 1, 0},
        {"f16 N=128 FIR (A rows shifted, B Hankel)", 1, 128, a_lbo, a_pitch, 128, 128, 1, 8, 24, 64, 1, 0},
        {"f16 N=64 FIR, commit every 12", 1, 64, a_lbo, a_pitch, 128, 128, 1, 8, 24, 64, 1, 12},
        {"f16 N=64 FIR x2 issuers", 1, 64, a_lbo, a_pitch, 128, 128, 1, 8, 24, 64, 2, 0},
        {"tf32 N=64 FIR (A rows shifted, B Hankel 256/128)", 0, 64, a_lbo, a_pitch, 128, 256, 1, 8, 24, 64, 1, 0},
        {"tf32 N=128 FIR", 0, 128, a_lbo, a_pitch, 128, 256, 1, 8, 24, 32, 1, 0},
        // same operands for every MMA (does moving the start address cost anything?)
        {"f16 N=64 FIR layout, fixed addresses", 1, 64, a_lbo, a_pitch, 128, 128, 0, 0, 1, 1, 1, 0},
    };
    for (const Case& c : cases) {
        Params P{};
        P.idesc = (1u << 4) | (c.kind == 0 ? ((2u << 7) | (2u << 10)) : 0u) | ((uint32_t)(c.n >> 3) << 17) | ((128u >> 4) << 24);
        P.a_lbo = c.a_lbo; P.a_sbo = c.a_sbo; P.b_lbo = c.b_lbo; P.b_sbo = c.b_sbo;
        P.a_step = c.a_step; P.b_step = c.b_step; P.a_mod = c.a_mod; P.b_mod = c.b_mod;
        P.n_mma = 4096; P.issuers = c.issuers; P.acc_cols = 512 / c.issuers < c.n ? c.n : (c.n <= 128 ? 128 : 256);
        if (P.acc_cols * c.issuers > 512) P.acc_cols = 512 / c.issuers;
        P.commit_every = c.commit_every; P.cycles = d_cyc;
        auto launch = [&]() {
            if (c.kind == 0) umma_loop<0><<<grid, 128, smem>>>(P); else umma_loop<1><<<grid, 128, smem>>>(P);
        };
        launch(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        std::vector<long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long cmax = 0; for (long long v : cyc) cmax = v > cmax ? v : cmax;
        const int K = c.kind == 0 ? 8 : 16;
        const double flop = 2.0 * 128 * c.n * K * (double)P.n_mma * c.issuers * grid;
        printf("{\"case\": \"%s\", \"clk_per_mma\": %.1f, \"tflops_burst\": %.1f, \"ms\": %.4f, \"operand_bytes_per_clk\": %.1f}\n", c.name,
               (double)cmax / P.n_mma / c.issuers, flop / (best * 1e-3) / 1e12, best,
               (double)(128 * 32 + c.n * 32) * P.n_mma * c.issuers / (double)cmax);
        fflush(stdout);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    // sustained tensor roof: N = 256 loops back to back for sustain_s seconds, per kind
    for (int kind = 1; kind >= 0; --kind) {
        Params P{};
        P.idesc = (1u << 4) | (kind == 0 ? ((2u << 7) | (2u << 10)) : 0u) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        P.a_lbo = 2048; P.a_sbo = 128; P.b_lbo = 4096; P.b_sbo = 128; P.a_mod = P.b_mod = 1;
        P.n_mma = 1 << 16; P.issuers = 1; P.acc_cols = 256; P.cycles = d_cyc;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0; float ms = 0;
        do {
            for (int i = 0; i < 8; ++i) { if (kind == 0) umma_loop<0><<<grid, 128, smem>>>(P); else umma_loop<1><<<grid, 128, smem>>>(P); }
            launches += 8;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        } while (ms < sustain_s * 1e3);
        const int K = kind == 0 ? 8 : 16;
        const double flop = 2.0 * 128 * 256 * K * (double)P.n_mma * grid * launches;
        printf("{\"case\": \"%s N=256 sustained %.1f s\", \"tflops_sustained\": %.1f}\n", kind == 0 ? "tf32" : "f16", ms * 1e-3, flop / (ms * 1e-3) / 1e12);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return 0;
}
