// Stand-alone correctness / speed probe of the tensor-core FIR kernel (afp_fir_tc.cuh) against a
// float64 CPU convolution.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I audio_filtering_pipeline_b200/csrc \
//        -o bench/microbench/tc_fir_test bench/microbench/tc_fir_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "afp_fir_tc.cuh"
#ifndef AFP_TC_ROW
#define AFP_TC_ROW 64
#endif
using Cfg = afp::TcCfg<AFP_TC_ROW>;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

static float trunc13(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1001;
    const int S = argc > 2 ? atoi(argv[2]) : 64;
    const int frames = argc > 3 ? atoi(argv[3]) : 512;
    const int iters = argc > 4 ? atoi(argv[4]) : 1;
    const int R = AFP_TC_ROW;
    const int n_chunks = (N + R - 2) / R + 1;          // d = 0 .. D, D = ceil((N-1+R-1)/R)
    const int maxd = getenv("MAXD") ? atoi(getenv("MAXD")) : 1000;   // debug: keep only chunks d <= MAXD
    const int mind = getenv("MIND") ? atoi(getenv("MIND")) : 0;
    const int hist_rows = n_chunks - 1;
    const int hist_len = hist_rows * R;
    printf("N=%d S=%d frames=%d n_chunks=%d hist_rows=%d\n", N, S, frames, n_chunks, hist_rows);
    std::vector<double> h(N);
    srand(1);
    for (int k = 0; k < N; ++k) h[k] = ((rand() / (double)RAND_MAX) - 0.5) * 4.0 / N * (1.0 + 3.0 * exp(-k / 50.0));
    std::vector<float> hist((size_t)S * hist_len), in((size_t)S * frames);
    for (auto& v : hist) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    for (auto& v : in) v = 0.2f * ((rand() / (float)RAND_MAX) - 0.5f);
    // W slices: chunk d, k-step ks, column chunk cc (2), row n (0..63 hi: j = n, 64..127 lo: j = n-64), 4 floats:
    // B[j][i] = h[R d + j - i] with i = 8 ks + 4 cc + e
    std::vector<float> w((size_t)n_chunks * Cfg::W_FLOATS, 0.f);
    for (int d = 0; d < n_chunks; ++d)
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks)
            for (int cc = 0; cc < 2; ++cc)
                for (int j = 0; j < R; ++j)
                    for (int e = 0; e < 4; ++e) {
                        const int i = 8 * ks + 4 * cc + e;
                        const int k = R * d + j - i;
                        const float hv = (k >= 0 && k < N && d <= maxd && d >= mind) ? (float)h[k] : 0.f;
                        const float hi = trunc13(hv), lo = hv - hi;
                        const size_t base = (size_t)d * Cfg::W_FLOATS + (size_t)ks * Cfg::W_SLICE + (size_t)cc * (2 * R * 4);
                        w[base + (size_t)j * 4 + e] = hi;
                        w[base + (size_t)(R + j) * 4 + e] = lo;
                    }
    float *d_hist, *d_in, *d_hout, *d_out, *d_w;
    CK(cudaMalloc(&d_hist, hist.size() * 4)); CK(cudaMalloc(&d_in, in.size() * 4));
    CK(cudaMalloc(&d_hout, hist.size() * 4)); CK(cudaMalloc(&d_out, in.size() * 4)); CK(cudaMalloc(&d_w, w.size() * 4));
    CK(cudaMemcpy(d_hist, hist.data(), hist.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, in.size() * 4));
    afp::TcParams P{};
    P.hist_in = d_hist; P.in = d_in; P.in_pitch = frames; P.hist_out = d_hout; P.out = d_out; P.out_pitch = frames;
    P.w = d_w; P.n_streams = S; P.frames = frames; P.hist_rows = hist_rows; P.n_chunks = n_chunks;
    P.n_items = (S + 15) / 16; P.fmt = afp::FMT_F32; P.dither = afp::DITHER_NONE; P.decim = 1; P.fast_out = 1;
    P.seed = 0; P.sample_base = 0; P.stream_id0 = 0; P.gain = 1.0f; P.dither_lsb = 0.f;
    const int rows = hist_rows + frames / R;
    const int data_bytes = 2 * 16 * rows * R * 4;
    int slots = (232448 - data_bytes - 512) / (Cfg::SLOT_FLOATS * 4);
    if (slots > 8) slots = 8;
    if (getenv("SLOTS")) slots = atoi(getenv("SLOTS"));
    if (slots < 2) { printf("does not fit in shared memory\n"); return 1; }
    P.w_slots = slots;
    if (getenv("MMA_N") || getenv("MMA_M")) {   // timing experiment: other instruction shapes (results are garbage)
        const unsigned n = getenv("MMA_N") ? atoi(getenv("MMA_N")) : 32, m = getenv("MMA_M") ? atoi(getenv("MMA_M")) : 128;
        P.idesc_override = (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
        (void)m;
    }
    const int smem = data_bytes + slots * Cfg::SLOT_FLOATS * 4 + 512;
    printf("W ring slots %d\n", slots);
    auto kern = frames / (8 * R) == 1 ? afp::fir_tc_kernel<afp::EPI_GENERIC, 1, AFP_TC_ROW> : (frames / (8 * R) == 2 ? afp::fir_tc_kernel<afp::EPI_GENERIC, 2, AFP_TC_ROW> : afp::fir_tc_kernel<afp::EPI_GENERIC, 4, AFP_TC_ROW>);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = P.n_items < 148 ? P.n_items : 148;
    printf("smem %d bytes, grid %d\n", smem, grid);
    kern<<<grid, afp::TC_THREADS, smem>>>(P);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(in.size()), hout(hist.size());
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hout.data(), d_hout, hout.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int s : {0, 1, 7, 8, 15, 16, S - 1}) {
        if (s >= S) continue;
        for (int n = 0; n < frames; ++n) {
            double acc = 0;
            for (int k = 0; k < N; ++k) {
                const int pos = hist_len + n - k;      // index into [hist | in]
                if (pos < 0) break;
                const int dd = (hist_len + n) / R - pos / R;
                if (dd > maxd || dd < mind) continue;
                const double xv = pos < hist_len ? hist[(size_t)s * hist_len + pos] : in[(size_t)s * frames + pos - hist_len];
                acc += h[k] * xv;
            }
            const double e = fabs(acc - out[(size_t)s * frames + n]);
            if (!(e <= 1e-5)) { if (bad < 5) printf("  mismatch s=%d n=%d ref=%g got=%g\n", s, n, acc, out[(size_t)s * frames + n]); ++bad; }
            if (e > maxerr) maxerr = e;
            if (fabs(acc) > maxref) maxref = fabs(acc);
        }
    }
    int hbad = 0;
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < hist_len; ++i) {
            const int pos = frames + i;
            const float ref = pos < hist_len ? hist[(size_t)s * hist_len + pos] : in[(size_t)s * frames + pos - hist_len];
            if (hout[(size_t)s * hist_len + i] != ref) ++hbad;
        }
    printf("max abs err %.3e (max |ref| %.3f), out-of-tolerance %d, history mismatches %d\n", maxerr, maxref, bad, hbad);
    if (iters > 1) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) kern<<<grid, afp::TC_THREADS, smem>>>(P);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
        printf("%.4f ms per launch, %.2f Gsamples/s, %.1f TFLOP/s (2N flop per sample)\n", ms, (double)S * frames / ms / 1e6,
               2.0 * N * S * frames / ms / 1e9);
    }
    return bad || hbad;
}
