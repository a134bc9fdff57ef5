// Stand-alone speed / stall probe of the split-FP16 tensor FIR kernel (afp_fir_tc2.cuh): runs one launch on
// synthetic data, checks a few streams against a float64 convolution and prints where the MMA warp's cycles go
// (waiting for accumulators / window rows / shared rows, issuing) - the evidence behind DESIGN.md section 5b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I audio_filtering_pipeline_b200/csrc \
//             -o bench/microbench/tc2_probe bench/microbench/tc2_probe.cu
// Run:   bench/microbench/tc2_probe [taps 1001] [streams 16384] [frames 512] [blocks 16] [iters 10] [upsample 1 | 2 | 4]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "afp_fir_tc2.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1001, S = argc > 2 ? atoi(argv[2]) : 16384;
    const int frames = argc > 3 ? atoi(argv[3]) : 512, NB = argc > 4 ? atoi(argv[4]) : 16, iters = argc > 5 ? atoi(argv[5]) : 10;
    const int U = argc > 6 ? atoi(argv[6]) : 1;           // 2 / 4: polyphase interpolator, U * frames output samples per block
    std::vector<double> h(N);
    srand(1);
    for (int k = 0; k < N; ++k) h[k] = ((rand() / (double)RAND_MAX) - 0.5) * 4.0 / N * (1.0 + 3.0 * exp(-k / 50.0));
    std::vector<__half> w;
    const int n_chunks = afp::tc2_build_cores(h.data(), N, w, U);
    const int H = 16, hist_len = ((N + U - 1) / U + 63) / 64 * 64;
    std::vector<float> hist((size_t)S * hist_len), in((size_t)NB * S * frames);
    for (auto& v : hist) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    for (auto& v : in) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    float *d_hist, *d_hout, *d_in, *d_out; __half* d_w; unsigned int* d_flag; long long* d_prof;
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int n_groups = (S + 15) / 16, grid = std::min(n_groups, prop.multiProcessorCount);
    CK(cudaMalloc(&d_hist, hist.size() * 4)); CK(cudaMalloc(&d_hout, hist.size() * 4));
    CK(cudaMalloc(&d_in, in.size() * 4)); CK(cudaMalloc(&d_out, in.size() * 4 * U));
    CK(cudaMalloc(&d_w, w.size() * 2)); CK(cudaMalloc(&d_flag, 4)); CK(cudaMalloc(&d_prof, sizeof(long long) * 16 * grid));
    CK(cudaMemcpy(d_hist, hist.data(), hist.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_flag, 0, 4)); CK(cudaMemset(d_prof, 0, sizeof(long long) * 16 * grid));
    afp::Tc2Params P{};
    P.hist_in = d_hist; P.hist_out = d_hout; P.hist_pitch = hist_len; P.hist_len = hist_len;
    P.in = d_in; P.in_pitch = frames; P.in_block_stride = (long long)S * frames;
    P.out = d_out; P.out_pitch = (long long)U * frames; P.out_block_stride = (long long)S * frames * U;
    P.w = d_w; P.n_cores = afp::tc2_n_cores(n_chunks, U); P.n_streams = S; P.frames = frames; P.n_blocks = NB;
    P.n_chunks = n_chunks; P.n_groups = n_groups; P.range_flag = d_flag; P.prof = d_prof;
    P.fmt = afp::FMT_F32; P.dither = afp::DITHER_NONE; P.decim = 1; P.fast_out = 1; P.gain = 1.f;
    if (getenv("I24")) { P.fmt = afp::FMT_I24_IN_I32; P.dither = afp::DITHER_TPDF; }
    if (getenv("P24")) { P.fmt = afp::FMT_I24_PACKED; P.dither = afp::DITHER_TPDF; }    // the bench headline's epilogue
    const bool nostore = getenv("NOSTORE") != nullptr;
    P.epi_stage = (!getenv("NOSTAGE") && afp::tc2_smem_bytes(n_chunks, true, U) <= 227 * 1024) ? 1 : 0;
    const int smem = afp::tc2_smem_bytes(n_chunks, P.epi_stage != 0, U);
    auto kern = U == 2 ? afp::fir_tc2_kernel<afp::epi_code(afp::FMT_F32, afp::DITHER_NONE), false, 2>
                : U == 3 ? afp::fir_tc2_kernel<afp::epi_code(afp::FMT_F32, afp::DITHER_NONE), false, 3>
                : U == 4 ? afp::fir_tc2_kernel<afp::epi_code(afp::FMT_F32, afp::DITHER_NONE), false, 4>
                : getenv("P24") ? afp::fir_tc2_kernel<afp::epi_code(afp::FMT_I24_PACKED, afp::DITHER_TPDF)>
                : getenv("I24") ? afp::fir_tc2_kernel<afp::epi_code(afp::FMT_I24_IN_I32, afp::DITHER_TPDF)> : afp::fir_tc2_kernel<afp::epi_code(afp::FMT_F32, afp::DITHER_NONE)>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    printf("taps %d streams %d frames %d blocks %d: chunks %d, H %d, smem %d, grid %d\n", N, S, frames, NB, n_chunks, H, smem, grid);
    kern<<<grid, afp::TC2_THREADS, smem>>>(P);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    if (nostore) P.gain = 0.f;        // timing runs without the epilogue's stores (the check above used them)
    if (!getenv("I24") && !getenv("P24")) {
        std::vector<float> out(in.size() * U);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int s : {0, 1, 15, 16, S / 2, S - 1}) {
            if (s >= S) continue;
            for (int on = 0; on < frames * NB * U; on += 7) {
                const int n = on / U, ph = on % U;                 // out[U n + ph] = sum_j h[U j + ph] x[n - j]
                double acc = 0;
                for (int j = 0; U * j + ph < N; ++j) {
                    const int k = U * j + ph;
                    const int t = n - j;
                    double xv;
                    if (t >= 0) xv = in[((size_t)(t / frames) * S + s) * frames + t % frames];
                    else if (t >= -hist_len) xv = hist[(size_t)s * hist_len + hist_len + t];
                    else break;
                    acc += h[k] * xv;
                }
                const double got = out[((size_t)(n / frames) * S + s) * frames * U + (size_t)(n % frames) * U + ph];
                const double e = fabs(acc - got);
                if (!(e <= 1e-5)) { if (bad < 5) printf("  mismatch s=%d out=%d ref=%g got=%g\n", s, on, acc, got); ++bad; }
                maxerr = std::max(maxerr, e);
            }
        }
        printf("max abs err %.3e, out of tolerance %d\n", maxerr, bad);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) kern<<<grid, afp::TC2_THREADS, smem>>>(P);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    std::vector<long long> prof((size_t)16 * grid);
    CK(cudaMemcpy(prof.data(), d_prof, prof.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double tot = 0, acc = 0, dir = 0, stg = 0, iss = 0, steps = 0, ld = 0, slot = 0, place = 0, ew = 0, et = 0, es = 0; long long tmax = 0;
    for (int b = 0; b < grid; ++b) {
        const long long* o = &prof[(size_t)b * 16];
        tot += o[0]; acc += o[1]; dir += o[2]; stg += o[3]; iss += o[4]; steps += o[5];
        ld += o[6]; slot += o[7]; place += o[8]; ew += o[9]; et += o[10]; es += o[11];
        tmax = std::max(tmax, o[0]);
    }
    const double mma_clk = U == 1 ? (64.0 + 48.0) * n_chunks * 4 : 3 * 64.0 * n_chunks * 4 * (U / 2) + (U == 3 ? 3 * 48.0 * n_chunks * 4 : 0);   // U = 1: N = 128 + N = 64 per chunk and k-step; else 3 x N = 128 per pass
    printf("{\"taps\": %d, \"streams\": %d, \"frames\": %d, \"blocks\": %d, \"ms\": %.4f, \"gsamples_s\": %.1f, \"sm_mhz_from_clk\": %.0f, "
           "\"clk_per_step\": %.0f, \"model_mma_clk_per_step\": %.0f, \"wait_acc\": %.0f, \"wait_window\": %.0f, \"wait_shared_rows\": %.0f, \"issue\": %.0f, "
           "\"loader_load\": %.0f, \"loader_wait_slot\": %.0f, \"loader_split_store\": %.0f, \"epi_wait_mma\": %.0f, \"epi_tmem\": %.0f, \"epi_store\": %.0f}\n",
           N, S, frames, NB, ms, (double)S * frames * NB * U / ms / 1e6, tmax / (ms * 1e-3) / 1e6, tot / steps, mma_clk, acc / steps, dir / steps, stg / steps, iss / steps,
           2 * ld / steps, 2 * slot / steps, 2 * place / steps, ew / steps, et / steps, es / steps);
    return 0;
}
