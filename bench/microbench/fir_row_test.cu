// Stand-alone correctness / speed probe of the single-row FIR kernel (afp_fir_row.cuh) in its streaming
// form (history | block -> outputs + next history) against a float64 CPU convolution.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I audio_filtering_pipeline_b200/csrc \
//        -o bench/microbench/fir_row_test bench/microbench/fir_row_test.cu
// Run:  fir_row_test N U T S frames iters tg_log2 store(0 direct,1 bulk,2 transposed) [ctas_per_sm] [dbg] [pad]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "afp_fir_row.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

using afp::FirParams;
typedef void (*kern_t)(const FirParams);
constexpr int EPI = afp::epi_code(afp::FMT_F32, afp::DITHER_NONE);

template <int U, int T, int PAD> kern_t pick(int store) {
    return store == 2 ? (kern_t)afp::fir_row_kernel<U, T, 256, EPI, 2, PAD>
         : store == 1 ? (kern_t)afp::fir_row_kernel<U, T, 256, EPI, 1, PAD> : (kern_t)afp::fir_row_kernel<U, T, 256, EPI, 0, PAD>;
}
template <int PAD> kern_t lookup_p(int U, int T, int store) {
    if (U == 1 && T == 16) return pick<1, 16, PAD>(store);
    if (U == 1 && T == 8) return pick<1, 8, PAD>(store);
    if (U == 2 && T == 8) return pick<2, 8, PAD>(store);
    if (U == 2 && T == 16) return pick<2, 16, PAD>(store);
    if (U == 4 && T == 8) return pick<4, 8, PAD>(store);
    if (U == 4 && T == 4) return pick<4, 4, PAD>(store);
    return nullptr;
}
static kern_t lookup(int U, int T, int store, int pad) { return pad ? lookup_p<1>(U, T, store) : lookup_p<0>(U, T, store); }
// specialised instantiations (compile-time group size, streaming form) of the shapes the library uses
static kern_t lookup_spec(int U, int T, int store, int pad, int tgl) {
    if (U == 1 && T == 16 && store == 0 && pad == 1 && tgl == 5) return (kern_t)afp::fir_row_kernel<1, 16, 256, EPI, 0, 1, 5, true>;
    if (U == 1 && T == 16 && store == 0 && pad == 1 && tgl == -5) return (kern_t)afp::fir_row_kernel<1, 16, 128, EPI, 0, 1, 5, true>;   // 128-thread CTAs
    if (U == 2 && T == 8 && store == 0 && pad == 0 && tgl == -6) return (kern_t)afp::fir_row_kernel<2, 8, 128, EPI, 0, 0, 6, true>;
    if (U == 2 && T == 8 && store == 0 && pad == 0 && tgl == 6) return (kern_t)afp::fir_row_kernel<2, 8, 256, EPI, 0, 0, 6, true>;
    if (U == 2 && T == 16 && store == 2 && pad == 0 && tgl == 5) return (kern_t)afp::fir_row_kernel<2, 16, 256, EPI, 2, 0, 5, true>;
    if (U == 4 && T == 4 && store == 2 && pad == 0 && tgl == 7) return (kern_t)afp::fir_row_kernel<4, 4, 256, EPI, 2, 0, 7, true>;
    if (U == 4 && T == 8 && store == 2 && pad == 0 && tgl == 5) return (kern_t)afp::fir_row_kernel<4, 8, 256, EPI, 2, 0, 5, true>;
    return nullptr;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 51;
    const int U = argc > 2 ? atoi(argv[2]) : 4;
    const int T = argc > 3 ? atoi(argv[3]) : 8;
    const int S = argc > 4 ? atoi(argv[4]) : 1024;
    const int frames = argc > 5 ? atoi(argv[5]) : 1024;
    const int iters = argc > 6 ? atoi(argv[6]) : 10;
    const int tg_log2 = argc > 7 ? atoi(argv[7]) : 5;
    const int store = argc > 8 ? atoi(argv[8]) : 0;
    int ctas = argc > 9 ? atoi(argv[9]) : 0;
    const int dbg = argc > 10 ? atoi(argv[10]) : 0;   // needs -DAFP_ROW_DBG: 1 no FMA, 2 no stores, 4 no loads
    const int NT = (argc > 13 && atoi(argv[13]) == 128) ? 128 : 256, NBUF = 3;   // argv[13]: CTA size
    const int pad = argc > 11 ? atoi(argv[11]) : 0;   // 1: padded layout + cp.async loader, 0: linear + bulk copies
    const int spec = argc > 12 ? atoi(argv[12]) : 0;   // 1: compile-time group size + streaming form
    kern_t k = spec ? lookup_spec(U, T, store, pad, NT == 128 ? -tg_log2 : tg_log2) : lookup(U, T, store, pad);
    if (!k) { printf("no kernel for U=%d T=%d\n", U, T); return 1; }
    const int step = afp::row_tail_step(U, T);
    const int slots_needed = (((N + U - 1) / U + (U == 1 ? 1 : 0)) + step - 1) / step * step;   // tap slots per phase
    const int n_chunks = (slots_needed + T - 1) / T, kq = n_chunks * T;
    int tail = slots_needed - (n_chunks - 1) * T;
    if (tail == T || getenv("NOTAIL")) tail = 0;
    const int tg = 1 << tg_log2, lq = tg * T, slots = NT / tg;
    const int segs = (frames + lq - 1) / lq;
    const int smem = afp::row_smem_bytes(U, T, NT, tg_log2, kq, store, pad);
    CK(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, NT, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (const void*)k));
    if (ctas <= 0 || ctas > per_sm) ctas = per_sm;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const long long n_items = (long long)S * segs;
    long long grid = (n_items + slots - 1) / slots;
    if (grid > (long long)ctas * prop.multiProcessorCount) grid = (long long)ctas * prop.multiProcessorCount;

    std::vector<double> h(N);
    srand(1);
    for (int i = 0; i < N; ++i) h[i] = ((rand() / (double)RAND_MAX) - 0.5) * 4.0 / N * U;
    std::vector<float> hist((size_t)S * kq), in((size_t)S * frames);
    for (auto& v : hist) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    for (auto& v : in) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    const size_t n_out = (size_t)S * frames * U;
    float *d_hist, *d_hout, *d_in[NBUF], *d_out[NBUF];
    CK(cudaMalloc(&d_hist, hist.size() * 4)); CK(cudaMalloc(&d_hout, hist.size() * 4));
    for (int b = 0; b < NBUF; ++b) {
        CK(cudaMalloc(&d_in[b], in.size() * 4)); CK(cudaMalloc(&d_out[b], n_out * 4));
        CK(cudaMemcpy(d_in[b], in.data(), in.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_out[b], 0xFF, n_out * 4));
    }
    CK(cudaMemcpy(d_hist, hist.data(), hist.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_hout, 0xFF, hist.size() * 4));

    static FirParams P;
    memset(&P, 0, sizeof(P));
    P.src0 = d_hist; P.pitch0 = kq; P.len0 = kq;
    P.src1 = d_in[0]; P.pitch1 = frames; P.len1 = frames;
    P.q_begin = kq; P.n_q = frames;
    P.tg_log2 = tg_log2; P.segs_per_stream = segs; P.segs_shift = -1;
    for (int b = 0; b < 30; ++b) if ((1 << b) == segs) P.segs_shift = b;
    P.n_seg_total = (int)n_items; P.n_pairs = (int)((n_items + 1) / 2);
    P.n_chunks = n_chunks; P.n_sub = dbg ? dbg : tail; P.sub_stride = n_chunks;
    P.out = d_out[0]; P.out_pitch = (long long)frames * U;
    P.decim = 1; P.fmt = afp::FMT_F32; P.dither = afp::DITHER_NONE;
    P.hist_out = d_hout; P.hist_pitch = kq; P.hist_len = kq; P.hist_from = frames;
    P.fast0 = P.fast1 = P.fast_hist = P.fast_out = P.fast_all = 1;
    P.gain = 1.0f; P.dither_lsb = 0.f;
    for (int i = 0; i < N; ++i) P.taps[i] = (float)h[i];
    if (U == 1) for (int i = 0; i < N; ++i) P.taps[afp::row_tabb_offset(kq) + 1 + i] = (float)h[i];

    printf("N=%d U=%d T=%d S=%d frames=%d tg=%d store=%d: kq=%d chunks=%d tail=%d segs=%d smem=%d regs=%d ctas/SM=%d(max %d) grid=%lld\n",
           N, U, T, S, frames, tg, store, kq, n_chunks, tail, segs, smem, fa.numRegs, ctas, per_sm, grid);
    k<<<(int)grid, NT, smem>>>(P);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    // ---- check a few streams against float64
    std::vector<float> out(n_out), hout(hist.size());
    CK(cudaMemcpy(out.data(), d_out[0], n_out * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hout.data(), d_hout, hist.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0; long long bad = 0;
    const int check[] = {0, 1, S / 2, S - 1};
    for (int ci = 0; ci < 4; ++ci) {
        const int s = check[ci];
        auto X = [&](int n) -> double {   // logical sequence [hist | in]
            if (n < 0) return 0.0;
            return n < kq ? hist[(size_t)s * kq + n] : in[(size_t)s * frames + (n - kq)];
        };
        for (int q = 0; q < frames; ++q)
            for (int p = 0; p < U; ++p) {
                double a = 0;
                for (int j = 0; j * U + p < N; ++j) a += (double)(float)h[j * U + p] * X(kq + q - j);
                const double e = fabs(a - out[(size_t)s * frames * U + (size_t)q * U + p]);
                if (e > max_err) max_err = e;
                if (!(e < 2e-6)) { if (bad < 5) printf("  mismatch s=%d q=%d p=%d want %g got %g\n", s, q, p, a, out[(size_t)s * frames * U + (size_t)q * U + p]); ++bad; }
            }
        for (int i = 0; i < kq; ++i) {
            const float want = (float)X(frames + i);
            if (hout[(size_t)s * kq + i] != want) { if (bad < 5) printf("  hist mismatch s=%d i=%d\n", s, i); ++bad; }
        }
    }
    if (dbg) bad = 0;
    printf("  check: max_err %.3g, bad %lld -> %s\n", max_err, bad, bad ? "FAIL" : "ok");

    // ---- timing: rotate over NBUF input/output buffers (working set >> L2 for the large cases)
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 3; ++it) { P.src1 = d_in[it % NBUF]; P.out = d_out[it % NBUF]; k<<<(int)grid, NT, smem>>>(P); }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it) { P.src1 = d_in[it % NBUF]; P.out = d_out[it % NBUF]; k<<<(int)grid, NT, smem>>>(P); }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    const double gs = (double)n_out / ms / 1e6;
    printf("  RESULT N=%d U=%d T=%d tg=%d store=%d pad=%d spec=%d ctas=%d dbg=%d: %.4f ms  %.1f Gsamples/s  %.1f TFLOP/s (2N/U)  %.0f GB/s alg\n",
           N, U, T, tg, store, pad, spec, ctas, dbg, ms, gs, gs * 2.0 * ((N + U - 1) / U) / 1e3, gs * (4.0 / U + 4.0));
    return bad ? 2 : 0;
}
