// C-ABI of the streaming block-FIR engine (include/afp.h): context, stage planning, launches.
//
// A context owns, per GPU: the stage list of one processing chain, the ping-pong history buffers of
// every stage (the device-side replacement of the reference's sliding `input_buffer`,
// stream_process.py:45-46,97-98), f32 intermediates between stages, and the dither sample counter.
// Device I/O buffers belong to the caller.  Nothing here falls back to the CPU: every failure is a
// status code + message.
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/afp.h"
#include "afp_kernel_table.cuh"
#include "afp_fir_row.cuh"
#include "afp_fir_tc.cuh"
#include "afp_fir_tc2.cuh"
#include "afp_agc.cuh"
#include "afp_spectrum.cuh"
#include <map>

afp::tc2_kern_t afp_kernel_tc2(int epi, int up);                                    // afp_kernels_tc2.cu
afp::tc2_kern_t afp_kernel_tc2_grouped();
afp::agc_kern_t afp_kernel_agc(int variant);   // afp_kernels_agc.cu: variant AGC_*

using afp::FirParams;

namespace {

// message of the last failed context-free call (afp_create, afp_host_alloc, ...) of THIS thread: contexts are created one
// per caller thread (one per GPU), so the slot is per thread too
thread_local std::string g_create_error;

constexpr int kMaxSmem = 227 * 1024;
constexpr uint32_t kStateMagic = 0x41465031u;  // "AFP1"

enum StageKind { ST_INTERP = 0, ST_EQ = 1, ST_MAIN = 2 };

struct Geometry {
    int nt = 0, tg_log2 = 0, smem = 0, grid = 0, segs_per_stream = 0, resident = 0;
    int key_nq = -1, key_ns = -1;   // launch shape this geometry was computed for
};

struct Stage {
    int kind = ST_MAIN;
    bool active = false;
    int U = 1;                      // phases this stage computes
    int T = 16;                     // chunk size = base positions per thread
    int n_taps = 0;
    int kq_pad = 0;                 // padded taps per phase (history length in base samples)
    int n_chunks = 0;
    int n_sub = 1;                  // passes over the tap axis (EQ bands in per-band form)
    std::vector<double> taps;       // as given (EQ: the centre-aligned taps-domain sum)
    std::vector<float> table;       // n_sub tables, natural order, each zero padded to U*kq_pad
    float* hist[2] = {nullptr, nullptr};
    float* hist_scratch = nullptr;  // third history buffer: long blocks run on the tensor cores as a chain of sub-blocks
    long long hist_pitch = 0;
    int cur = 0;
    Geometry geo;
    // tensor-core (tcgen05 / TMEM, 3xTF32) path for long single-table U = 1 filters: Toeplitz tap
    // tiles on the device; which kernel ran last (0 = FFMA2, 1 = tcgen05)
    float* tc_w[2] = {nullptr, nullptr};   // tap tiles for row length 32 and 64
    int tc_chunks[2] = {0, 0};
    // split-FP16 tensor kernel (afp_fir_tc2.cuh): Hankel tap cores [2][n_cores][8][8] halves, chunk count D + 1
    void* tc2_w = nullptr;
    int tc2_cores = 0, tc2_chunks = 0;     // (U = 2 .. 4: the polyphase form of the same kernel)
    int last_kernel = 0;           // 0 = row-pair FFMA2, 1 = tcgen05 (3xTF32), 2 = single-row FFMA2, 3 = tcgen05 (split FP16)
    int last_T = 0;                // tile of the last launch (the single-row kernel picks its own)
    // per-launch host work of the single-row kernel that only depends on the stage: its tap-table image
    // and the occupancy of the kernel it last used
    std::vector<float> row_taps; int row_taps_T = 0;
    const void* occ_kernel = nullptr; int occ_smem = 0, occ_per_sm = 0;
};

// Kernels live in one translation unit per up-sampling factor (afp_kernels_u{1..4}.cu, compiled in
// parallel); each exports a lookup by (T, CTA size, multi-table, epilogue code) - afp_kernel_table.cuh.
using afp::kern_t;
using afp::kAllEpi;

kern_t kernel_for(int U, int T, int nt, bool multi, int epi) {
    switch (U) {
        case 1: return afp_kernel_u1(T, nt, multi, epi);
        case 2: return afp_kernel_u2(T, nt, multi, epi);
        case 3: return afp_kernel_u3(T, nt, multi, epi);
        case 4: return afp_kernel_u4(T, nt, multi, epi);
    }
    return nullptr;
}

kern_t row_kernel_for(int U, int T, int pad, int store, int epi, int spec_tgl) {
    switch (U) {
        case 1: return afp_kernel_row_u1(T, pad, store, epi, spec_tgl);
        case 2: return afp_kernel_row_u2(T, pad, store, epi, spec_tgl);
        case 3: return afp_kernel_row_u3(T, pad, store, epi, spec_tgl);
        case 4: return afp_kernel_row_u4(T, pad, store, epi, spec_tgl);
    }
    return nullptr;
}

// Shape of the single-row kernel (afp_fir_row.cuh) for an up-sampling factor and phase length, from the
// sweeps in bench/microbench/fir_row_test.cu on B200 (profiles/r01_experiments/row_probe*.txt):
//   U = 1: 16-sample tiles; the FMA pipe is the bound, so the window loads must be conflict free: padded
//          layout + cp.async; a thread's run is 64 B and transposing the stores does not pay.
//   U = 2: <= 32 taps per phase: 8-sample tiles, direct stores, 64-thread groups (741 vs 700 Gsamples/s
//          at 51 taps); longer: 16-sample tiles, transposed stores (514 vs 484 at 101 taps).
//   U = 4: <= 16 taps per phase: 4-sample tiles, 128-thread groups (1168 vs 1045 at 51 taps); longer:
//          8-sample tiles, one warp per group (925 vs 835 at 101 taps); transposed stores throughout.
struct RowShape { int T = 0, pad = 0, store = 0, tg_log2 = 5; };
RowShape pick_row_shape(int U, int kq) {
    RowShape r;
    if (U == 1) { r.T = 16; r.pad = 1; r.store = 0; r.tg_log2 = 5; }
    else if (U == 2) {
        if (kq <= 32) { r.T = 8; r.pad = 0; r.store = 0; r.tg_log2 = 6; }
        else { r.T = 16; r.pad = 0; r.store = 2; r.tg_log2 = 5; }
    } else if (U == 3) { r.T = 8; r.pad = 0; r.store = 2; r.tg_log2 = 5; }   // phase-major form, not swept
    else if (U == 4) {
        if (kq <= 16) { r.T = 4; r.pad = 0; r.store = 2; r.tg_log2 = 7; }
        else { r.T = 8; r.pad = 0; r.store = 2; r.tg_log2 = 5; }
    }
    return r;
}
const RowShape kRowShapes[] = {{16, 1, 0, 5}, {8, 0, 0, 6}, {16, 0, 2, 5}, {4, 0, 2, 7}, {8, 0, 2, 5}};

// Tile (= chunk) sizes compiled for each U, largest first.
const int kTiles[5][3] = {{0, 0, 0}, {16, 0, 0}, {8, 0, 0}, {4, 0, 0}, {8, 4, 0}};

// Chunk size T for a filter of n taps at up-sampling U (measured, bench/sweep_tiles.sh on B200): the
// largest tile wins for U = 1 (16) and U = 2 (8) at every length tried (51..501 taps); for U = 4 the
// 8-tile wins from 63 taps per phase up (fewer shared-memory / constant-bank words per FMA) and the
// 4-tile below (less zero padding of a 13..26-tap phase, 80 registers -> a third resident CTA).
// AFP_FORCE_T overrides (tuning experiments).
int tile_for(int U, int n_taps) {
    const int kq = (n_taps + U - 1) / U;
    if (const char* f = std::getenv("AFP_FORCE_T")) {
        const int t = std::atoi(f);
        for (int i = 0; i < 3; ++i) if (kTiles[U][i] && kTiles[U][i] == t) return t;
    }
    if (U == 4) return kq <= 32 ? 4 : 8;
    return kTiles[U][0];
}

}  // namespace

struct AgcPlanEntry { int group = 0, grid = 0; size_t smem = 0; };

struct afp_ctx {
    afp_config cfg{};
    std::string err;
    Stage up, eq, mainf;
    // AFP_CHAIN_REFERENCE without EQ: interpolate -> FIR -> keep every U-th sample is ONE base-rate FIR with
    // taps (h_up (*) h)[jU] (see rebuild_fused); when active it replaces the up / mainf launches
    Stage fused;
    std::vector<std::vector<double>> bands;
    std::vector<double> gains;
    // AGC ahead of the chain (afp_set_agc; stream_process_AGC.py:131-143): parameters and its output block
    bool agc_on = false;
    afp_agc_params agc{};
    float* tmp_agc = nullptr;        // [S][pitch_agc]
    long long pitch_agc = 0;
    float* tmp_a = nullptr;          // interp output  [S][max_frames*U]
    float* tmp_b = nullptr;          // EQ output      [S][pitch_b]
    long long pitch_a = 0, pitch_b = 0;
    // AFP_EQ_BUFFER_SAME*: the reference's sliding input_buffer minus the new block, i.e. the last
    // N_main-1 processing-rate samples (stream_process_EQ.py:44-45,119-121), ping-pong
    float* xhist[2] = {nullptr, nullptr};
    long long xhist_pitch = 0;
    int xhist_len = -1, xhist_cur = 0;
    bool xhist_migrate = false, xhist_reset = false;   // next ensure_tmp: keep (afp_swap_taps) / clear (afp_set_taps) the sliding buffer
    int eq_centre = 0;               // (N_eq-1)//2 of the widest band: 'same' centring
    unsigned long long out_samples = 0;   // kept output samples per stream since reset
    long long launches = 0;
    int n_sm = 148;
    int decim = 1;                   // keep every decim-th sample of the main stage's output
    int dither_bits = 24;            // LSB of the float-output dither
    // afp_process_host staging: a ring of device slots fed by three streams (H2D, compute, D2H)
    static constexpr int kSlots = 3;
    float* d_in[kSlots] = {nullptr, nullptr, nullptr};
    unsigned char* d_out[kSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[kSlots] = {}, ev_k[kSlots] = {}, ev_out[kSlots] = {};
    int slot_streams = 0;            // streams per staging slot
    std::map<std::pair<int, int>, AgcPlanEntry> agc_plans;   // launch shape of the AGC kernel per (variant, streams)
    // afp_spectrum_db: per block length n, device tables [window n floats | twiddles n/2 float2]
    std::map<int, float*> spec_tables;
    int spw_per_sm = 0;              // resident CTAs per SM of the warp-per-row spectrum kernel
    // afp_process_host_interleaved: frame-major device copies of one block in / out
    float* d_il_in = nullptr;
    unsigned char* d_il_out = nullptr;
    cudaStream_t own_stream = nullptr, s_in = nullptr, s_out = nullptr;
    FirParams params{};              // scratch for the kernel-parameter block of the launch being prepared
    unsigned int* range_flag = nullptr;   // device word: set by the split-FP16 tensor kernel when a sample exceeds FP16's range
    bool range_dirty = false;             // a split-FP16 launch has run since the flag was last read
    // per-stream-group filters (afp_assign_groups): every session of a batch with its own EQ gains / main taps, the way
    // each reference process reads its own sliders per block (stream_process_EQ_GUI.py:86-90).  Empty = one filter set.
    struct GroupDef { std::vector<double> gains, taps; };   // empty = the context-level gains / main taps
    std::vector<int> group_of_stream;
    std::vector<GroupDef> groups;
    bool groups_dirty = false, groups_pair_aligned = false;
    int* d_group_of_stream = nullptr;
    float* d_gtaps = nullptr;        // [G][gtaps_stride] per-group tables of the chain's single stage (layout of Stage::table)
    long long gtaps_stride = 0;
    void* d_gcores = nullptr;        // [G][gcores_stride] halves: per-group Hankel tap cores for the split-FP16 tensor kernel
    long long gcores_stride = 0;
};

namespace {

// tuning / debugging switches from the environment: set and neither empty nor "0"
bool env_flag(const char* name) {
    const char* v = std::getenv(name);
    return v && *v && std::strcmp(v, "0") != 0;
}

int fail(afp_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(afp_ctx* c, cudaError_t e, const char* what) {
    return fail(c, AFP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define AFP_CUDA(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail((c), e__, #call); } while (0)

int rate_of(const afp_ctx* c) {   // processing rate / base rate
    return c->cfg.upsample;
}

void free_stage(Stage& s) {
    for (int i = 0; i < 2; ++i) { if (s.hist[i]) cudaFree(s.hist[i]); s.hist[i] = nullptr; }
    if (s.hist_scratch) { cudaFree(s.hist_scratch); s.hist_scratch = nullptr; }
    for (int i = 0; i < 2; ++i) { if (s.tc_w[i]) cudaFree(s.tc_w[i]); s.tc_w[i] = nullptr; s.tc_chunks[i] = 0; }
    if (s.tc2_w) { cudaFree(s.tc2_w); s.tc2_w = nullptr; s.tc2_cores = s.tc2_chunks = 0; }
    s.active = false;
}

// (re)build one stage: n_sub tables of n taps each (natural order, double), pad, convert to float32,
// allocate + zero the carried history (kq_pad base-rate samples per stream)
int build_stage(afp_ctx* c, Stage& s, int kind, int U, const double* taps, int n, int n_sub = 1,
                bool keep_state = false) {
    if (n < 1) return fail(c, AFP_ERR_INVALID, "filter needs at least one tap");
    for (long long i = 0; i < (long long)n * n_sub; ++i)
        if (!std::isfinite(taps[i])) return fail(c, AFP_ERR_INVALID, "non-finite filter tap");
    const int T = tile_for(U, n);
    const int kq = (n + U - 1) / U;
    // Long single-table U = 1 filters can also run on the tensor cores (afp_fir_tc.cuh).  Both kernels
    // then share one history format: the last kq_pad samples with kq_pad a multiple of 32 (the tensor
    // kernel's row length), so a stage may switch kernel from one block to the next.
    // Crossovers against the single-row FFMA2 kernel, measured at 16384 streams (bench/quick_perf.py and
    // bench/microbench/tc2_probe with AFP_TC_MIN_TAPS 32 vs 5000, bench/gpu_calls/r02_tcmin*.sh, r02_tc2_ring.sh): the
    // split-FP16 kernel (afp_fir_tc2.cuh) is bound by its loaders and epilogue, not by the MMAs, below ~400 taps - one
    // 1024-frame block per launch: 295 / 286 / 243 / 208 Gsamples/s at 51 / 101 / 251 / 385 taps (FFMA2: 388 / 243 / 123 /
    // 83); 16 blocks of 512 frames per launch: 494 / 424 / 325 at 51 / 251 / 449 taps.  So cores are built from 48 taps
    // up and launch_stage_tc2 decides per call (>= 8 steps of 512 frames: from 48 taps; >= 2 steps: from 60; one step: from 112).  The 3xTF32
    // kernel (afp_fir_tc.cuh, 256-frame blocks) only wins from 385 taps up.
    const int tc_min = std::getenv("AFP_TC_MIN_TAPS") ? std::atoi(std::getenv("AFP_TC_MIN_TAPS")) : 400;
    const int tc2_min = std::getenv("AFP_TC_MIN_TAPS") ? std::atoi(std::getenv("AFP_TC_MIN_TAPS")) : 48;
    const bool tc_any = U == 1 && n_sub == 1 && !env_flag("AFP_NO_TC");
    // Polyphase interpolators (U = 2 .. 4) run on the same kernel with wider Toeplitz rows (tc2_build_cores).  Crossovers
    // against the FFMA2 row kernel, one 1024-frame block of 65536 streams per launch (bench/quick_perf.py): U = 2: 63 taps
    // 638 vs 719 Gsamples/s, 79 taps 635 vs 566; U = 4: 101 taps 958 vs 916, 129 taps 959 vs 768; U = 3: 51 taps 834 vs 978, 101 taps 828 vs 684.
    const bool tc2_up = U >= 2 && U <= 4 && n_sub == 1 && !env_flag("AFP_NO_TC") && !env_flag("AFP_NO_TC2_UP") &&
                        n >= (std::getenv("AFP_TC_MIN_TAPS") ? std::atoi(std::getenv("AFP_TC_MIN_TAPS")) : (U == 2 ? 72 : (U == 3 ? 84 : 96)));
    const bool tc1_ok = tc_any && n >= tc_min, tc2_ok = (tc_any && n >= tc2_min || tc2_up) && !env_flag("AFP_NO_TC2");
    const bool tc_ok = tc1_ok || tc2_ok;
    // History length = padded taps per phase.  A multiple of 16 (every tile size divides it, so the two
    // FFMA2 kernels share one history format and a stage may switch kernel from block to block), with room
    // for the extra tap slot the single-row kernel's U = 1 form needs (afp_fir_row.cuh).
    const int n_chunks = tc_ok ? ((kq + 63) / 64) * (64 / T) : ((kq + (U % 2 == 1 ? 1 : 0) + 15) / 16) * (16 / T);
    if ((long long)n_chunks * T * U * n_sub > afp::MAX_TAPS)
        return fail(c, AFP_ERR_INVALID, "filter (x bands) too long for the per-launch tap table (max 6144 padded taps)");
    const bool same_shape = s.active && s.kind == kind && s.U == U && s.n_chunks == n_chunks;
    if (!(keep_state && same_shape)) {
        // keep_state with a different padded length: the carried samples move over right-aligned (a history
        // row is the last hist_pitch input samples, oldest first); what the old row did not hold is zero
        float* old_hist = nullptr;
        long long old_pitch = 0;
        if (keep_state && s.active && s.kind == kind && s.U == U && s.hist[s.cur]) {
            old_hist = s.hist[s.cur]; old_pitch = s.hist_pitch;
            s.hist[s.cur] = nullptr;                      // survives free_stage
        }
        free_stage(s);
        s.hist_pitch = (long long)n_chunks * T;
        const size_t bytes = (size_t)c->cfg.n_streams * s.hist_pitch * sizeof(float);
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaMalloc(&s.hist[i], bytes);
            if (e == cudaSuccess) e = cudaMemset(s.hist[i], 0, bytes);
            if (e != cudaSuccess) { if (old_hist) cudaFree(old_hist); return cuda_fail(c, e, "history allocation"); }
        }
        s.cur = 0;
        s.geo = Geometry();
        if (old_hist) {
            const long long k = old_pitch < s.hist_pitch ? old_pitch : s.hist_pitch;
            cudaError_t e = cudaMemcpy2D(s.hist[0] + (s.hist_pitch - k), (size_t)s.hist_pitch * sizeof(float),
                                         old_hist + (old_pitch - k), (size_t)old_pitch * sizeof(float),
                                         (size_t)k * sizeof(float), (size_t)c->cfg.n_streams, cudaMemcpyDeviceToDevice);
            cudaFree(old_hist);
            if (e != cudaSuccess) return cuda_fail(c, e, "history migration");
        }
    }
    s.kind = kind; s.U = U; s.T = T; s.n_taps = n; s.n_sub = n_sub;
    s.taps.assign(taps, taps + (size_t)n * n_sub);
    s.row_taps.clear(); s.row_taps_T = 0;
    s.n_chunks = n_chunks; s.kq_pad = n_chunks * T;
    s.table.assign((size_t)s.kq_pad * U * n_sub, 0.f);
    for (int b = 0; b < n_sub; ++b)
        for (int i = 0; i < n; ++i) s.table[(size_t)b * s.kq_pad * U + i] = (float)taps[(size_t)b * n + i];
    for (int i = 0; i < 2; ++i) { if (s.tc_w[i]) cudaFree(s.tc_w[i]); s.tc_w[i] = nullptr; s.tc_chunks[i] = 0; }
    if (tc1_ok) {
        // W slices for both row lengths R: chunk d, k-step ks, column chunk cc, row (0..R-1: hi, j = row;
        // R..2R-1: lo, j = row-R), 4 floats:  W[j][i] = h[R d + j - i],  i = 8 ks + 4 cc + e;
        // hi = the value with its 13 low mantissa bits cleared, lo = value - hi
        for (int v = 0; v < 2; ++v) {
            const int R = v == 0 ? 32 : 64;
            const int ksteps = R / 8, w_slice = 2 * 2 * R * 4, w_floats = ksteps * w_slice;
            const int nc = (n + R - 2) / R + 1;
            std::vector<float> w((size_t)nc * w_floats, 0.f);
            for (int d = 0; d < nc; ++d)
                for (int ks = 0; ks < ksteps; ++ks)
                    for (int cc = 0; cc < 2; ++cc)
                        for (int j = 0; j < R; ++j)
                            for (int e = 0; e < 4; ++e) {
                                const int k = R * d + j - (8 * ks + 4 * cc + e);
                                if (k < 0 || k >= n) continue;
                                const float hv = (float)taps[k];
                                uint32_t u; std::memcpy(&u, &hv, 4); u &= 0xFFFFE000u;
                                float hi; std::memcpy(&hi, &u, 4);
                                const size_t base = (size_t)d * w_floats + (size_t)ks * w_slice + (size_t)cc * (2 * R * 4);
                                w[base + (size_t)j * 4 + e] = hi;
                                w[base + (size_t)(R + j) * 4 + e] = hv - hi;
                            }
            AFP_CUDA(c, cudaMalloc(&s.tc_w[v], w.size() * sizeof(float)));
            AFP_CUDA(c, cudaMemcpy(s.tc_w[v], w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
            s.tc_chunks[v] = nc;
        }
    }
    if (s.tc2_w) { cudaFree(s.tc2_w); s.tc2_w = nullptr; s.tc2_cores = s.tc2_chunks = 0; }
    if (tc2_ok) {
        // split-FP16 kernel (afp_fir_tc2.cuh): Hankel tap cores, all high parts then all low parts
        if ((n - 1 + (afp::TC2_R - 1) * U) / (afp::TC2_R * U) + 1 <= 17) {
            std::vector<__half> w;
            s.tc2_chunks = afp::tc2_build_cores(taps, n, w, U);
            s.tc2_cores = afp::tc2_n_cores(s.tc2_chunks, U);
            AFP_CUDA(c, cudaMalloc(&s.tc2_w, w.size() * sizeof(__half)));
            AFP_CUDA(c, cudaMemcpy(s.tc2_w, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
        }
    }
    // the memsets / copies above ran on the legacy stream; the kernels run on non-blocking streams that do
    // not order themselves after it
    AFP_CUDA(c, cudaDeviceSynchronize());
    s.active = true;
    return AFP_OK;
}

// CTA geometry for n_q base positions per stream with a halo of kq_pad.
// A thread group (tg threads) owns one segment of tg*T positions.  Small groups mean many independent
// groups per SM (one group's load/epilogue hides behind the others' FMAs) but re-load the halo more
// often; the halo is kept <= 1/4 of a segment unless the block itself is shorter.
Geometry pick_geometry(int T, int kq_pad, int n_q, int n_streams, int nt_min = 128) {
    Geometry best;
    double best_score = -1;
    for (int nt = 256; nt >= nt_min; nt >>= 1) {
        int cover = 3;                                   // smallest group that covers the whole block
        while ((1 << cover) * T < n_q && (1 << cover) < nt) ++cover;
        int halo = 5;                                    // smallest warp-or-larger group with halo <= 25 %
        while ((1 << halo) * T < 4 * kq_pad && (1 << halo) < nt) ++halo;
        int tg_log2 = cover < halo ? cover : halo;
        for (;; ++tg_log2) {
            const int tg = 1 << tg_log2, slots = nt / tg;
            const int smem = slots * afp::region_granules(kq_pad + tg * T, T) * 16;
            if (smem <= kMaxSmem) {
                const int max_cta = nt == 256 ? 2 : 4;
                int cta = kMaxSmem / (smem + 1024);
                if (cta > max_cta) cta = max_cta;
                if (cta < 1) cta = 1;
                const int segs = (n_q + tg * T - 1) / (tg * T);
                const double useful = (double)n_q / ((double)segs * tg * T);   // masked-lane waste
                const double score = cta * nt * useful + (nt == 256 ? 1 : 0);
                if (score > best_score) {
                    best_score = score;
                    best.nt = nt; best.tg_log2 = tg_log2; best.smem = smem; best.segs_per_stream = segs;
                    const long long pairs = ((long long)n_streams * segs + 1) / 2;
                    best.grid = (int)((pairs + slots - 1) / slots);
                }
                break;
            }
            if (tg >= nt) break;   // cannot fit even one segment per CTA
        }
    }
    return best;
}

// ---- AGC (afp_agc.cuh).  A work item is `group` streams x the whole block, one 256-sample row per stream in shared
// memory, at most 32 streams per item (one lane each in the recurrence phase).  Items are dealt round-robin over a
// persistent grid, every item takes the same time, and the recurrence runs at 13.5 clk per sample and lane whatever the
// group size - so the group is chosen for the fullest last wave (65536 streams: 32 per item are 2.3 waves of 888 CTAs,
// 22 per item 2.9 waves of 1036), and a small batch is spread over every SM.
using AgcPlan = AgcPlanEntry;
int plan_agc(afp_ctx* c, afp::agc_kern_t k, int n_streams, AgcPlan& pl) {
    double best = -1.0;
    for (int g = afp::AGC_MAX_GROUP; g >= 8; --g) {
        const size_t smem = afp::agc_smem_bytes(g);
        int per_sm = 0;
        AFP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, afp::AGC_THREADS, smem));
        if (per_sm < 1) continue;
        const long long slots = (long long)c->n_sm * per_sm, items = (n_streams + g - 1) / g;
        const long long waves = (items + slots - 1) / slots;
        // time ~ waves x (phases of one item, nearly independent of g); streams per SM in flight = per_sm x g
        const double eff = (double)items / (double)(waves * slots);
        const double score = (items >= slots ? eff : eff * 0.5 + 0.5 * (double)std::min<long long>(items, c->n_sm) / c->n_sm) + 1e-4 * g;
        if (score > best) { best = score; pl.group = g; pl.grid = (int)std::min(items, slots); pl.smem = smem; }
    }
    return best < 0 ? fail(c, AFP_ERR_INVALID, "AGC kernel does not fit on an SM") : AFP_OK;
}

int launch_agc(afp_ctx* c, const afp_agc_params& a, const float* in, long long in_pitch, float* out, long long out_pitch,
               int ns, int frames, cudaStream_t stream) {
    if (ns < 1 || frames < 1) return AFP_OK;
    afp::AgcParams P{};
    P.in = in; P.in_pitch = in_pitch; P.out = out; P.out_pitch = out_pitch;
    P.n_streams = ns; P.frames = frames;
    const int W = a.window_size;
    // window < 1: np.ones(0)/0 -> np.convolve raises; window > frames: the gains cannot be applied
    // (shape mismatch) - both end in the reference's except branch, block returned unchanged (:87-89)
    P.passthrough = (W < 1 || W > frames) ? 1 : 0;
    P.window = W < 1 ? 1 : W; P.left = P.window / 2; P.right = (P.window - 1) / 2;   // np.convolve 'same' centring
    P.inv_window = (float)(1.0 / P.window);
    P.target = (float)a.target_level; P.max_gain = (float)a.max_gain; P.min_gain = 0.1f;   // :83
    P.clip = (float)a.clip;
    const int att_n = (int)(a.attack * (double)W), rel_n = (int)(a.release * (double)W);   // :56-57
    P.att_ok = att_n != 0; P.rel_ok = rel_n != 0;
    const double aa = P.att_ok ? 1.0 - std::exp(-1.0 / att_n) : 0.0, ar = P.rel_ok ? 1.0 - std::exp(-1.0 / rel_n) : 0.0;   // :73-78
    P.a_att = (float)aa; P.c_att = (float)(1.0 - aa); P.a_rel = (float)ar; P.c_rel = (float)(1.0 - ar);
    // The reference divides by (rms + 1e-10).  Wherever the gain is not clamped to max_gain, rms exceeds
    // target/max_gain; if that is >= 1e-4 the epsilon moves the gain by < 1e-6 relative and the kernel may
    // use target*rsqrt(mean) (a zero window then gives +inf -> max_gain, as the reference's 1e9 clip does).
    const bool fast = a.target_level > 0 && a.max_gain > 0 && a.target_level / a.max_gain >= 1e-4;
    P.fast_scale = (float)(a.target_level * std::sqrt((double)P.window));
    int variant = !(P.att_ok && P.rel_ok) ? afp::AGC_CHECKED : (fast ? afp::AGC_FAST : afp::AGC_EXACT);
    P.probe = std::getenv("AFP_AGC_PROBE") ? std::atoi(std::getenv("AFP_AGC_PROBE")) : 0;   // timing probes, not a product path
    if (variant == afp::AGC_FAST && (P.probe & 1)) variant = afp::AGC_PROBE_F32;
    afp::agc_kern_t k = afp_kernel_agc(variant);
    AgcPlan& pl = c->agc_plans[{variant, ns}];
    if (!pl.group) {
        if (int rc = plan_agc(c, k, ns, pl)) return rc;
        if (const char* v = std::getenv("AFP_AGC_GROUP")) {   // tuning experiments
            const int g = std::atoi(v);
            if (g >= 1 && g <= afp::AGC_MAX_GROUP) {
                int per_sm = 1;
                pl.group = g; pl.smem = afp::agc_smem_bytes(g);
                AFP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, afp::AGC_THREADS, pl.smem));
                pl.grid = (int)std::min<long long>((ns + g - 1) / g, (long long)c->n_sm * std::max(per_sm, 1));
            }
        }
    }
    P.group = pl.group;
    P.n_items = (ns + P.group - 1) / P.group;
    k<<<pl.grid, afp::AGC_THREADS, pl.smem, stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
    return AFP_OK;
}

struct StageIO {
    const float* src0 = nullptr; long long pitch0 = 0; int len0 = 0;
    const float* src1 = nullptr; long long pitch1 = 0; int len1 = 0;
    int q_begin = 0, n_q = 0;
    void* out = nullptr; long long out_pitch = 0;
    int fmt = AFP_FMT_F32, dither = AFP_DITHER_NONE, decim = 1;
    unsigned long long sample_base = 0;
    float* hist_out = nullptr; long long hist_pitch = 0; int hist_len = 0, hist_from = 0;
    int s0 = 0, ns = 0;              // stream sub-range [s0, s0+ns) this launch covers
};

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Tensor-core launch for a streaming block (history | new block) of a TC-eligible stage.  Returns
// AFP_OK and sets *done when the block shape qualifies; otherwise leaves *done false.
int launch_stage_tc(afp_ctx* c, Stage& s, const StageIO& io, cudaStream_t stream, bool* done) {
    *done = false;
    if (!s.tc_w[0] || io.decim != 1 || s.U != 1) return AFP_OK;
    const int frames = io.len1;
    if (io.len0 != s.kq_pad || io.q_begin != io.len0 || io.n_q != frames || !io.hist_out ||
        io.hist_len != s.kq_pad || io.hist_from != frames || io.pitch0 != s.kq_pad || io.hist_pitch != s.kq_pad) return AFP_OK;
    if (!aligned16(io.src0) || !aligned16(io.src1) || !aligned16(io.hist_out) || io.pitch1 % 4 != 0) return AFP_OK;
    const int epi = afp::epi_code(io.fmt, io.dither);
    // row length 64 first (fewer, larger MMAs), then 32; each needs whole 8-row groups and room in
    // shared memory for the 16-stream hi/lo tiles plus at least two 8 KB ring slots
    for (int v = 1; v >= 0; --v) {
        const int R = v == 0 ? 32 : 64;
        if (std::getenv("AFP_TC_ROW") && std::atoi(std::getenv("AFP_TC_ROW")) != R) continue;   // tuning experiments
        if (frames % (8 * R) != 0) continue;
        const int nts = frames / (8 * R);
        afp::tc_kern_t k = afp_kernel_tc(R, nts, epi);
        if (!k) continue;
        const int hist_rows = s.kq_pad / R, rows = hist_rows + frames / R;
        const int data_bytes = 2 * afp::TC_STREAMS * rows * R * 4;
        const int slot_bytes = 8192;
        int slots = (kMaxSmem - data_bytes - 512) / slot_bytes;
        if (slots < 2) continue;
        if (slots > 8) slots = 8;
        afp::TcParams P{};
        P.hist_in = io.src0; P.in = io.src1; P.in_pitch = io.pitch1; P.hist_out = io.hist_out;
        P.out = io.out; P.out_pitch = io.out_pitch; P.w = s.tc_w[v];
        P.n_streams = io.ns; P.frames = frames; P.hist_rows = hist_rows; P.n_chunks = s.tc_chunks[v];
        P.n_items = (io.ns + afp::TC_STREAMS - 1) / afp::TC_STREAMS;
        P.idesc_override = 0; P.w_slots = slots;
        P.fmt = io.fmt; P.dither = io.dither; P.decim = 1;
        const int eb = io.fmt == AFP_FMT_I16 ? 2 : (io.fmt == AFP_FMT_I24_PACKED ? 3 : 4);
        P.fast_out = (aligned16(io.out) && (io.out_pitch * eb) % 16 == 0) ? 1 : 0;
        P.seed = c->cfg.seed; P.sample_base = io.sample_base; P.stream_id0 = c->cfg.first_stream_id + io.s0;
        P.gain = 1.0f; P.dither_lsb = std::ldexp(1.0f, -(c->dither_bits - 1));
        const int smem = data_bytes + slots * slot_bytes + 512;
        const int grid = P.n_items < c->n_sm ? P.n_items : c->n_sm;
        k<<<grid, afp::TC_THREADS, smem, stream>>>(P);
        AFP_CUDA(c, cudaGetLastError());
        c->launches += 1;
        s.last_kernel = 1;
        s.geo.nt = afp::TC_THREADS; s.geo.tg_log2 = 4; s.geo.smem = smem; s.geo.grid = grid; s.geo.key_nq = -1;
        *done = true;
        return AFP_OK;
    }
    return AFP_OK;
}

// Split-FP16 tensor kernel (afp_fir_tc2.cuh): n_blocks consecutive blocks of a streaming stage in ONE launch,
// history read from s.hist[cur] and written to s.hist[cur ^ 1] (the caller flips cur once).  *done stays
// false when the stage / shape does not qualify.
int launch_stage_tc2(afp_ctx* c, Stage& s, const float* in, long long in_pitch, long long in_bstride, void* out,
                     long long out_pitch, long long out_bstride, int frames, int n_blocks, int fmt, int dither,
                     unsigned long long sample_base, int s0, int ns, cudaStream_t stream, bool* done) {
    *done = false;
    if (!s.tc2_w || s.U < 1 || s.U > 4 || env_flag("AFP_NO_TC2")) return AFP_OK;
    if (!c->group_of_stream.empty()) return AFP_OK;   // per-group filters: launch_stage_tc2_grouped / the row-pair kernel
    if (frames < afp::TC2_STEP || frames % afp::TC2_STEP != 0 || n_blocks < 1) return AFP_OK;
    if ((long long)frames * n_blocks > (1 << 30)) return AFP_OK;
    // short filters: only when a stream group's window start is amortised over enough steps (see build_stage)
    // (one 1024-frame block of 65536 streams per launch, tensor vs FFMA2 kernel: 51 taps 430 vs 464 Gsamples/s, 63 taps 429 vs 410,
    // 79 taps 414 vs 351, 101 taps 410 vs 284; a single 512-frame step pays the whole window start)
    {
        const long long steps = (long long)(frames / afp::TC2_STEP) * n_blocks;
        if (!std::getenv("AFP_TC_MIN_TAPS") && s.U == 1 && ((s.n_taps < 60 && steps < 8) || (s.n_taps < 112 && steps < 2))) return AFP_OK;
    }
    const int eb = fmt == AFP_FMT_I16 ? 2 : (fmt == AFP_FMT_I24_PACKED ? 3 : 4);
    if (!aligned16(in) || in_pitch % 4 != 0 || in_bstride % 4 != 0 || s.hist_pitch % 4 != 0 || s.kq_pad % 8 != 0) return AFP_OK;
    if (s.tc2_chunks > 17 || s.kq_pad > 24 * afp::TC2_R) return AFP_OK;
    afp::Tc2Params P{};
    P.hist_in = s.hist[s.cur] + (long long)s0 * s.hist_pitch; P.hist_out = s.hist[s.cur ^ 1] + (long long)s0 * s.hist_pitch;
    P.hist_pitch = s.hist_pitch; P.hist_len = s.kq_pad;
    P.in = in; P.in_pitch = in_pitch; P.in_block_stride = in_bstride;
    P.out = out; P.out_pitch = out_pitch; P.out_block_stride = out_bstride;
    P.w = static_cast<const __half*>(s.tc2_w); P.n_cores = s.tc2_cores;
    P.n_streams = ns; P.frames = frames; P.n_blocks = n_blocks; P.n_chunks = s.tc2_chunks;
    P.n_groups = (ns + afp::TC2_STREAMS - 1) / afp::TC2_STREAMS;
    P.range_flag = c->range_flag;
    P.fmt = fmt; P.dither = dither; P.decim = 1;
    P.fast_out = (aligned16(out) && (out_pitch * eb) % 16 == 0 && (out_bstride * eb) % 16 == 0) ? 1 : 0;
    P.seed = c->cfg.seed; P.sample_base = sample_base; P.stream_id0 = c->cfg.first_stream_id + s0;
    P.gain = 1.0f; P.dither_lsb = std::ldexp(1.0f, -(c->dither_bits - 1));
    // transposed epilogue stores when the staging tiles fit beside the window tile and the tap cores (filters up to ~640 taps)
    P.epi_stage = afp::tc2_smem_bytes(s.tc2_chunks, true, s.U) <= kMaxSmem && !env_flag("AFP_TC2_NO_STAGE") ? 1 : 0;
    const int smem = afp::tc2_smem_bytes(s.tc2_chunks, P.epi_stage != 0, s.U);
    if (smem > kMaxSmem) return AFP_OK;
    const int grid = P.n_groups < c->n_sm ? P.n_groups : c->n_sm;
    afp_kernel_tc2(afp::epi_code(fmt, dither), s.U)<<<grid, afp::TC2_THREADS, smem, stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->range_dirty = true;
    s.last_kernel = 3;
    s.geo.nt = afp::TC2_THREADS; s.geo.tg_log2 = 4; s.geo.smem = smem; s.geo.grid = grid; s.geo.key_nq = -1;
    *done = true;
    return AFP_OK;
}

// The same kernel with per-stream tap cores (afp_assign_groups): a work item is one stream, whose blocks fill the 16 tile
// slots as 16 consecutive runs - so the call must bring a multiple of 16 x 512 frames per stream.
int launch_stage_tc2_grouped(afp_ctx* c, Stage& s, const float* in, long long in_pitch, long long in_bstride, void* out,
                             long long out_pitch, long long out_bstride, int frames, int n_blocks, int fmt, int dither,
                             unsigned long long sample_base, int s0, int ns, cudaStream_t stream, bool* done) {
    *done = false;
    if (!s.tc2_w || !c->d_gcores || s.U != 1 || env_flag("AFP_NO_TC2")) return AFP_OK;
    if (frames < afp::TC2_STEP || frames % afp::TC2_STEP != 0 || n_blocks < 1) return AFP_OK;
    const long long steps = (long long)(frames / afp::TC2_STEP) * n_blocks;
    if (steps % afp::TC2_STREAMS != 0 || (long long)frames * n_blocks > (1 << 30)) return AFP_OK;
    const int eb = fmt == AFP_FMT_I16 ? 2 : (fmt == AFP_FMT_I24_PACKED ? 3 : 4);
    if (!aligned16(in) || in_pitch % 4 != 0 || in_bstride % 4 != 0 || s.hist_pitch % 4 != 0 || s.kq_pad % 8 != 0) return AFP_OK;
    if (s.tc2_chunks > 17 || s.kq_pad > 24 * afp::TC2_R) return AFP_OK;
    afp::Tc2Params P{};
    P.hist_in = s.hist[s.cur] + (long long)s0 * s.hist_pitch; P.hist_out = s.hist[s.cur ^ 1] + (long long)s0 * s.hist_pitch;
    P.hist_pitch = s.hist_pitch; P.hist_len = s.kq_pad;
    P.in = in; P.in_pitch = in_pitch; P.in_block_stride = in_bstride;
    P.out = out; P.out_pitch = out_pitch; P.out_block_stride = out_bstride;
    P.w = static_cast<const __half*>(s.tc2_w); P.n_cores = s.tc2_cores;
    P.n_streams = ns; P.frames = frames; P.n_blocks = n_blocks; P.n_chunks = s.tc2_chunks;
    P.n_groups = ns;                                   // work items = streams
    P.steps_per_slot = (int)(steps / afp::TC2_STREAMS);
    P.gw = static_cast<const __half*>(c->d_gcores); P.gw_stride = c->gcores_stride; P.group_of_stream = c->d_group_of_stream + s0;
    P.range_flag = c->range_flag;
    P.fmt = fmt; P.dither = dither; P.decim = 1;
    P.fast_out = (aligned16(out) && (out_pitch * eb) % 16 == 0 && (out_bstride * eb) % 16 == 0) ? 1 : 0;
    P.seed = c->cfg.seed; P.sample_base = sample_base; P.stream_id0 = c->cfg.first_stream_id + s0;
    P.gain = 1.0f; P.dither_lsb = std::ldexp(1.0f, -(c->dither_bits - 1));
    P.epi_stage = afp::tc2_smem_bytes(s.tc2_chunks, true, 1, true) <= kMaxSmem && !env_flag("AFP_TC2_NO_STAGE") ? 1 : 0;
    const int smem = afp::tc2_smem_bytes(s.tc2_chunks, P.epi_stage != 0, 1, true);
    if (smem > kMaxSmem) return AFP_OK;
    const int grid = ns < c->n_sm ? ns : c->n_sm;
    afp_kernel_tc2_grouped()<<<grid, afp::TC2_THREADS, smem, stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->range_dirty = true;
    s.last_kernel = 3;
    s.geo.nt = afp::TC2_THREADS; s.geo.tg_log2 = 4; s.geo.smem = smem; s.geo.grid = grid; s.geo.key_nq = -1;
    *done = true;
    return AFP_OK;
}

// 16-byte alignment facts of one launch (the loaders' fast paths)
struct Align { int fast0, fast1, fast_hist, fast_all, fast_out; };
Align alignment_of(const Stage& s, const StageIO& io) {
    Align a;
    const int eb = io.fmt == AFP_FMT_I16 ? 2 : (io.fmt == AFP_FMT_I24_PACKED ? 3 : 4);
    const int start0 = io.q_begin - s.kq_pad;   // logical index of the first sample of segment 0's region
    const bool n_al = start0 % 4 == 0;          // segment lengths are multiples of 4 in every geometry
    a.fast0 = (io.len0 > 0 && n_al && aligned16(io.src0) && io.pitch0 % 4 == 0) ? 1 : 0;
    a.fast1 = (io.len1 > 0 && n_al && aligned16(io.src1) && io.pitch1 % 4 == 0 && io.len0 % 4 == 0) ? 1 : 0;
    a.fast_hist = (io.hist_out && n_al && aligned16(io.hist_out) && io.hist_pitch % 4 == 0 && io.hist_from % 4 == 0) ? 1 : 0;
    // aligned fast loader: both sources vector-aligned, both lengths multiples of 4, history (if any) too
    a.fast_all = (n_al && (io.len0 == 0 || a.fast0) && (io.len1 == 0 || (aligned16(io.src1) && io.pitch1 % 4 == 0))
                  && io.len0 % 4 == 0 && io.len1 % 4 == 0
                  && (!io.hist_out || (a.fast_hist && io.hist_len % 4 == 0))) ? 1 : 0;
    a.fast_out = (aligned16(io.out) && (io.out_pitch * eb) % 16 == 0) ? 1 : 0;
    return a;
}

void fill_common(afp_ctx* c, const Stage& s, const StageIO& io, const Align& a, FirParams& P) {
    P.src0 = io.src0; P.src1 = io.src1;     // all pointers arrive offset to stream s0
    P.pitch0 = io.pitch0; P.pitch1 = io.pitch1;
    P.len0 = io.len0; P.len1 = io.len1;
    P.q_begin = io.q_begin; P.n_q = io.n_q;
    P.out = io.out; P.out_pitch = io.out_pitch;
    P.decim = io.decim; P.fmt = io.fmt; P.dither = io.dither;
    P.seed = c->cfg.seed; P.sample_base = io.sample_base; P.stream_id0 = c->cfg.first_stream_id + io.s0;
    P.hist_out = io.hist_out; P.hist_pitch = io.hist_pitch; P.hist_len = io.hist_len; P.hist_from = io.hist_from;
    P.gain = 1.0f;
    P.dither_lsb = std::ldexp(1.0f, -(c->dither_bits - 1));
    P.fast0 = a.fast0; P.fast1 = a.fast1; P.fast_hist = a.fast_hist; P.fast_all = a.fast_all; P.fast_out = a.fast_out;
}

// 25 KB of launch parameters: kept in the context (off the stack; contexts are single-owner, so one
// scratch copy per context is race free even with one host thread per GPU)
FirParams& launch_params(afp_ctx* c) { return c->params; }

// Single-row kernel (afp_fir_row.cuh) for single-table stages (U = 1..4) whose buffers are all
// 16-byte aligned.  Sets *done when it launched; otherwise the row-pair kernel takes the block.
int launch_stage_row(afp_ctx* c, Stage& s, const StageIO& io, const Align& al, cudaStream_t stream, bool* done) {
    *done = false;
    if (env_flag("AFP_NO_ROW") || s.n_sub != 1 || !al.fast_all) return AFP_OK;
    if (s.U < 1 || s.U > 4) return AFP_OK;
    const int U = s.U, kq = (s.n_taps + U - 1) / U;
    RowShape sh = pick_row_shape(U, kq);
    const int T = sh.T;
    const int step = afp::row_tail_step(U, T);
    const int slots = (kq + (U % 2 == 1 ? 1 : 0) + step - 1) / step * step;     // tap slots per phase the FMA loop visits
    if (slots > s.kq_pad || s.kq_pad % T != 0) return AFP_OK;
    if (U % 2 == 1 && U * afp::row_phase_stride(s.kq_pad) > afp::MAX_TAPS) return AFP_OK;
    const int n_loop = (slots + T - 1) / T;
    int tail = slots - (n_loop - 1) * T;
    if (tail == T) tail = 0;
    int tg_log2 = sh.tg_log2;
    while (tg_log2 > 5 && (T << (tg_log2 - 1)) >= io.n_q) --tg_log2;      // short blocks: no idle warps
    const int NT = 256;
    int smem = afp::row_smem_bytes(U, T, NT, tg_log2, s.kq_pad, sh.store, sh.pad);
    while (smem > kMaxSmem && tg_log2 > 5) { --tg_log2; smem = afp::row_smem_bytes(U, T, NT, tg_log2, s.kq_pad, sh.store, sh.pad); }
    if (smem > kMaxSmem) return AFP_OK;
    const int epi = io.decim == 1 ? afp::epi_code(io.fmt, io.dither) : afp::EPI_GENERIC;
    // streaming form (history | block, outputs for the block): the instantiation with the group size and the
    // unclamped copy ranges compiled in, when the group size is the shape's own
    const bool streaming = io.q_begin == s.kq_pad && io.len0 == s.kq_pad && io.len1 >= 4;
    kern_t k = row_kernel_for(U, T, sh.pad, sh.store, epi, streaming ? tg_log2 : 0);
    if (!k) return AFP_OK;
    const int S = io.ns;
    const int lq = T << tg_log2, slots_cta = NT >> tg_log2;
    const int segs = (io.n_q + lq - 1) / lq;
    const long long n_items = (long long)S * segs;
    if (n_items <= 0 || n_items > INT_MAX) return AFP_OK;
    if (s.occ_kernel != (const void*)k || s.occ_smem != smem) {
        int per_sm_q = 0;
        AFP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_q, (const void*)k, NT, smem));
        s.occ_kernel = (const void*)k; s.occ_smem = smem; s.occ_per_sm = per_sm_q;
    }
    const int per_sm = s.occ_per_sm;
    if (per_sm < 1) return AFP_OK;
    const long long want = (n_items + slots_cta - 1) / slots_cta, resident = (long long)per_sm * c->n_sm;
    const int grid = (int)(want < resident ? want : resident);

    FirParams& P = launch_params(c);
    fill_common(c, s, io, al, P);
    P.tg_log2 = tg_log2; P.segs_per_stream = segs; P.segs_shift = -1;
    for (int b = 0; b < 30; ++b) if ((1 << b) == segs) P.segs_shift = b;
    P.n_seg_total = (int)n_items; P.n_pairs = (int)((n_items + 1) / 2);
    P.n_chunks = s.kq_pad / T; P.n_sub = tail; P.sub_stride = n_loop;
    if (s.row_taps_T != T || s.row_taps.size() != (size_t)afp::MAX_TAPS) {
        s.row_taps.assign(afp::MAX_TAPS, 0.f);
        if (U % 2 == 0) {
            std::memcpy(s.row_taps.data(), s.table.data(), (size_t)s.kq_pad * U * sizeof(float));
        } else {
            // odd U: per phase p the sub-filter h_p[j] = h[jU+p] in natural order, then the same table moved up
            // by one sample (pairs that start at an odd tap index)
            const int ps = afp::row_phase_stride(s.kq_pad), ob = afp::row_tabb_offset(s.kq_pad);
            for (int p = 0; p < U; ++p)
                for (int j = 0; j < s.kq_pad; ++j) {
                    const float v = s.table[(size_t)j * U + p];
                    s.row_taps[p * ps + j] = v;
                    s.row_taps[p * ps + ob + 1 + j] = v;
                }
        }
        s.row_taps_T = T;
    }
    std::memcpy(P.taps, s.row_taps.data(), sizeof(P.taps));
    k<<<grid, NT, smem, stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches += 1;
    s.last_kernel = 2; s.last_T = T;
    s.geo.nt = NT; s.geo.tg_log2 = tg_log2; s.geo.smem = smem; s.geo.grid = grid; s.geo.segs_per_stream = segs;
    s.geo.key_nq = -1;   // the row-pair kernel recomputes its own geometry if it runs next
    *done = true;
    return AFP_OK;
}

int launch_stage(afp_ctx* c, Stage& s, const StageIO& io, cudaStream_t stream) {
    const bool grouped = !c->group_of_stream.empty();   // per-group tap tables: the row-pair kernel reads them per work item
    const Align al = alignment_of(s, io);
    if (!grouped) {
        bool done = false;
        int rc = launch_stage_tc(c, s, io, stream, &done);
        if (rc || done) return rc;
        rc = launch_stage_row(c, s, io, al, stream, &done);
        if (rc || done) return rc;
    }
    s.last_kernel = 0; s.last_T = s.T;
    const int S = io.ns;
    if (s.geo.key_nq != io.n_q || s.geo.key_ns != S) {
        Geometry g = pick_geometry(s.T, s.kq_pad, io.n_q, S, grouped ? 256 : 128);
        if (g.nt == 0) return fail(c, AFP_ERR_INVALID, "filter/block combination does not fit in shared memory");
        // persistent grid: no more CTAs than fit on the device at once (items are dealt round-robin)
        int per_sm = 0;
        AFP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, (const void*)kernel_for(s.U, s.T, g.nt, s.n_sub > 1, afp::EPI_GENERIC), g.nt, g.smem));
        if (per_sm < 1) return fail(c, AFP_ERR_INVALID, "kernel does not fit on an SM with this geometry");
        // ... and never fewer than min(resident, items): items are dealt CTA-first, so a small launch
        // spreads over all SMs instead of filling every group of a few CTAs
        const long long n_pairs = ((long long)S * g.segs_per_stream + 1) / 2;
        const int resident = per_sm * c->n_sm;
        g.grid = (int)(n_pairs < resident ? n_pairs : resident);
        g.resident = resident;
        g.key_nq = io.n_q; g.key_ns = S;
        s.geo = g;
    }
    const Geometry& g = s.geo;
    FirParams& P = launch_params(c);
    fill_common(c, s, io, al, P);
    P.tg_log2 = g.tg_log2; P.segs_per_stream = g.segs_per_stream;
    P.segs_shift = -1;
    for (int b = 0; b < 30; ++b) if ((1 << b) == g.segs_per_stream) P.segs_shift = b;
    P.n_seg_total = S * g.segs_per_stream;
    P.n_pairs = (P.n_seg_total + 1) / 2;
    P.n_chunks = s.n_chunks; P.n_sub = s.n_sub; P.sub_stride = s.kq_pad * s.U;
    std::memset(P.taps, 0, sizeof(P.taps));
    std::memcpy(P.taps, s.table.data(), s.table.size() * sizeof(float));
    const int epi = io.decim == 1 ? afp::epi_code(io.fmt, io.dither) : afp::EPI_GENERIC;
    kern_t k = nullptr;
    int grid = g.grid;
    if (grouped) {
        if (s.n_sub != 1 || !c->d_gtaps || c->gtaps_stride != (long long)s.kq_pad * s.U)
            return fail(c, AFP_ERR_STATE, "per-group filters are not built for this stage");
        k = afp_kernel_grouped(s.U, s.T);
        P.gtaps = c->d_gtaps; P.gstride = c->gtaps_stride; P.group_of_stream = c->d_group_of_stream + io.s0;
        // the two rows of a work item share their taps: pairs of segments of one stream, or of two streams of one group
        P.pair_single = (g.segs_per_stream % 2 == 1 && !c->groups_pair_aligned) ? 1 : 0;
        if (P.pair_single) {
            P.n_pairs = P.n_seg_total;
            grid = P.n_pairs < g.resident ? P.n_pairs : g.resident;
        }
    } else {
        k = kernel_for(s.U, s.T, g.nt, s.n_sub > 1, epi);
    }
    if (!k) return fail(c, AFP_ERR_INVALID, "no kernel for this stage shape");
    k<<<grid, g.nt, g.smem, stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return AFP_OK;
}

// streaming launch: history | new block, outputs for the new block, history rolled forward
// Largest block (512 or 256 frames... up to 1024) one tensor-kernel launch of this stage can take: the
// 16-stream hi/lo tiles of [history | block] plus two 8 KB tap-ring slots must fit in shared memory.
int tc_max_block(const Stage& s) {
    if (!s.tc_w[0] || s.U != 1) return 0;
    for (int f : {1024, 512, 256}) {
        const int R = f % 512 == 0 ? 64 : 32;
        if (2 * afp::TC_STREAMS * (s.kq_pad + f) * 4 + 2 * 8192 + 512 <= kMaxSmem && afp_kernel_tc(R, f / (8 * R), afp::EPI_GENERIC)) return f;
    }
    return 0;
}

int run_streaming(afp_ctx* c, Stage& s, const float* in, long long in_pitch, int len,
                  void* out, long long out_pitch, int fmt, int dither, int decim,
                  unsigned long long sample_base, int s0, int ns, cudaStream_t stream) {
    // A block too long for one tensor-kernel launch (the reference's default blocksize is 4096,
    // stream_process.py:23) runs as a chain of sub-blocks, each a complete streaming launch: history A ->
    // scratch -> A -> ... -> B, so the block still ends with its history in the "other" buffer like any
    // single launch (+30 % over the FFMA2 kernel for >= 400-tap filters).
    if (decim == 1) {
        bool done = false;
        int rc = launch_stage_tc2(c, s, in, in_pitch, 0, out, out_pitch, 0, len, 1, fmt, dither, sample_base, s0, ns, stream, &done);
        if (rc || done) return rc;
    }
    const int sub = (decim == 1 && c->group_of_stream.empty() && !env_flag("AFP_NO_TC_CHAIN")) ? tc_max_block(s) : 0;
    if (sub > 0 && len > sub && len % sub == 0 && aligned16(in) && in_pitch % 4 == 0) {
        if (!s.hist_scratch) {
            const size_t bytes = (size_t)c->cfg.n_streams * s.hist_pitch * sizeof(float);
            AFP_CUDA(c, cudaMalloc(&s.hist_scratch, bytes));
        }
        const int k = len / sub;
        const int eb = fmt == AFP_FMT_I16 ? 2 : (fmt == AFP_FMT_I24_PACKED ? 3 : 4);
        float* const A = s.hist[s.cur];
        float* const B = s.hist[s.cur ^ 1];
        for (int j = 0; j < k; ++j) {
            float* hin = j == 0 ? A : ((j & 1) ? s.hist_scratch : A);
            float* hout = j == k - 1 ? B : (((j + 1) & 1) ? s.hist_scratch : A);
            StageIO io;
            io.s0 = s0; io.ns = ns;
            io.src0 = hin + (long long)s0 * s.hist_pitch; io.pitch0 = s.hist_pitch; io.len0 = s.kq_pad;
            io.src1 = in + (long long)j * sub; io.pitch1 = in_pitch; io.len1 = sub;
            io.q_begin = s.kq_pad; io.n_q = sub;
            io.out = static_cast<unsigned char*>(out) + (size_t)j * sub * eb; io.out_pitch = out_pitch;
            io.fmt = fmt; io.dither = dither; io.decim = 1;
            io.sample_base = sample_base + (unsigned long long)j * sub;
            io.hist_out = hout + (long long)s0 * s.hist_pitch; io.hist_pitch = s.hist_pitch; io.hist_len = s.kq_pad; io.hist_from = sub;
            int rc = launch_stage(c, s, io, stream);
            if (rc) return rc;
        }
        return AFP_OK;
    }
    StageIO io;
    io.s0 = s0; io.ns = ns;
    io.src0 = s.hist[s.cur] + (long long)s0 * s.hist_pitch; io.pitch0 = s.hist_pitch; io.len0 = s.kq_pad;
    io.src1 = in; io.pitch1 = in_pitch; io.len1 = len;
    io.q_begin = s.kq_pad; io.n_q = len;
    io.out = out; io.out_pitch = out_pitch; io.fmt = fmt; io.dither = dither; io.decim = decim;
    io.sample_base = sample_base;
    io.hist_out = s.hist[s.cur ^ 1] + (long long)s0 * s.hist_pitch; io.hist_pitch = s.hist_pitch; io.hist_len = s.kq_pad; io.hist_from = len;
    return launch_stage(c, s, io, stream);   // the caller flips s.cur once the whole block is launched
}

bool eq_is_per_band(const afp_ctx* c) {
    return c->cfg.eq_mode == AFP_EQ_PER_BAND || c->cfg.eq_mode == AFP_EQ_BUFFER_SAME_PER_BAND;
}
bool eq_is_buffer_same(const afp_ctx* c) {
    return c->cfg.eq_mode == AFP_EQ_BUFFER_SAME || c->cfg.eq_mode == AFP_EQ_BUFFER_SAME_PER_BAND;
}

// EQ tables.  Bands are aligned on their 'same' centres (N_b-1)//2, the way scipy centres each band's
// output in stream_process_EQ.py:128; fused form: one table sum_b g_b*h_b (stream_process_EQ_GUI.py:
// 86-90); per-band form: one table g_b*h_b per band, accumulated in the kernel.
int rebuild_eq(afp_ctx* c, bool reset_state) {
    if (c->bands.empty()) { free_stage(c->eq); return AFP_OK; }
    int cmax = 0;
    for (auto& b : c->bands) cmax = std::max(cmax, ((int)b.size() - 1) / 2);
    int nall = 0;
    for (auto& b : c->bands) nall = std::max(nall, cmax - ((int)b.size() - 1) / 2 + (int)b.size());
    const int nb = (int)c->bands.size();
    const int n_sub = eq_is_per_band(c) ? nb : 1;
    std::vector<double> tab((size_t)n_sub * nall, 0.0);
    for (int i = 0; i < nb; ++i) {
        const int off = cmax - ((int)c->bands[i].size() - 1) / 2;
        double* dst = tab.data() + (size_t)(n_sub == 1 ? 0 : i) * nall;
        for (size_t k = 0; k < c->bands[i].size(); ++k) dst[off + k] += c->gains[i] * c->bands[i][k];
    }
    c->eq_centre = cmax;
    return build_stage(c, c->eq, ST_EQ, 1, tab.data(), nall, n_sub, !reset_state);
}

int ensure_tmp(afp_ctx* c) {
    const int S = c->cfg.n_streams, U = c->cfg.upsample;
    const bool ref = c->cfg.chain == AFP_CHAIN_REFERENCE;
    const long long rate_len = (long long)c->cfg.max_frames * (ref ? U : 1);   // EQ-rate samples per block
    if (c->agc_on && !c->tmp_agc) {
        c->pitch_agc = ((long long)c->cfg.max_frames + 3) / 4 * 4;
        AFP_CUDA(c, cudaMalloc(&c->tmp_agc, (size_t)S * c->pitch_agc * sizeof(float)));
    }
    if (ref && U > 1 && !c->tmp_a && !c->fused.active) {   // the collapsed chain has no processing-rate intermediate
        c->pitch_a = (rate_len + 3) / 4 * 4;
        AFP_CUDA(c, cudaMalloc(&c->tmp_a, (size_t)S * c->pitch_a * sizeof(float)));
    }
    const bool stateless = c->cfg.eq_mode == AFP_EQ_STATELESS_SAME;
    if (eq_is_buffer_same(c)) {
        // carried samples: N_main-1, or (fixed buffer) everything but the newest full-size block
        const int hm = c->cfg.same_buffer_len > 0 ? c->cfg.same_buffer_len - (int)rate_len : c->mainf.n_taps - 1;
        if (hm < c->mainf.n_taps - 1)
            return fail(c, AFP_ERR_INVALID, "same_buffer_len is shorter than max_frames*rate + N_main - 1");
        if (c->xhist_len != hm || c->xhist_reset) {
            // afp_swap_taps with another filter length: the carried samples move over right-aligned (a row is the last
            // xhist_len processing-rate input samples, oldest first); afp_set_taps: cleared like every other state
            float* old = (c->xhist_migrate && !c->xhist_reset && c->xhist[c->xhist_cur]) ? c->xhist[c->xhist_cur] : nullptr;
            const long long old_pitch = c->xhist_pitch;
            const int old_len = c->xhist_len;
            for (int i = 0; i < 2; ++i) { if (c->xhist[i] && c->xhist[i] != old) cudaFree(c->xhist[i]); c->xhist[i] = nullptr; }
            c->xhist_pitch = (hm + 3) / 4 * 4 + 4;
            for (int i = 0; i < 2; ++i) {
                cudaError_t e = cudaMalloc(&c->xhist[i], (size_t)S * c->xhist_pitch * sizeof(float));
                if (e == cudaSuccess) e = cudaMemset(c->xhist[i], 0, (size_t)S * c->xhist_pitch * sizeof(float));
                if (e != cudaSuccess) { if (old) cudaFree(old); return cuda_fail(c, e, "sliding-buffer allocation"); }
            }
            if (old) {
                const int k = old_len < hm ? old_len : hm;
                cudaError_t e = k > 0 ? cudaMemcpy2D(c->xhist[0] + (hm - k), (size_t)c->xhist_pitch * sizeof(float), old + (old_len - k),
                                                     (size_t)old_pitch * sizeof(float), (size_t)k * sizeof(float), (size_t)S,
                                                     cudaMemcpyDeviceToDevice) : cudaSuccess;
                cudaFree(old);
                if (e != cudaSuccess) return cuda_fail(c, e, "sliding-buffer migration");
            }
            c->xhist_migrate = c->xhist_reset = false;
            AFP_CUDA(c, cudaDeviceSynchronize());   // (legacy-stream memsets vs. the non-blocking streams of the kernels)
            c->xhist_len = hm; c->xhist_cur = 0;
            if (c->tmp_b) { cudaFree(c->tmp_b); c->tmp_b = nullptr; }
        }
    }
    if ((c->eq.active || stateless) && !c->tmp_b && !c->fused.active) {
        const long long need = rate_len + (eq_is_buffer_same(c) ? c->xhist_len : 0);
        c->pitch_b = (need + 3) / 4 * 4;
        AFP_CUDA(c, cudaMalloc(&c->tmp_b, (size_t)S * c->pitch_b * sizeof(float)));
    }
    return AFP_OK;
}

// ---- wire format of a sounddevice stream (stream_process.py:128-136: [frames, channels] float32 buffers)
// <-> the planar stream-major layout of the kernels.  in[(session*frames + f)*C + ch] <-> row session*C + ch.
__global__ void __launch_bounds__(256) deinterleave_kernel(const float* __restrict__ in, float* __restrict__ out, int sessions,
                                                           int frames, int C, long long out_pitch) {
    const long long total = (long long)sessions * frames;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int se = (int)(t / frames), f = (int)(t - (long long)se * frames);
        const float* src = in + t * C;
        for (int ch = 0; ch < C; ++ch) out[((long long)se * C + ch) * out_pitch + f] = src[ch];
    }
}
// planar rows of EB-byte samples -> frame-major [session][sample][channel]
template <int EB>
__global__ void __launch_bounds__(256) interleave_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                         int sessions, int n, int C, long long in_pitch_bytes) {
    const long long total = (long long)sessions * n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int se = (int)(t / n), j = (int)(t - (long long)se * n);
        unsigned char* dst = out + t * C * EB;
        for (int ch = 0; ch < C; ++ch) {
            const unsigned char* src = in + ((long long)se * C + ch) * in_pitch_bytes + (long long)j * EB;
            if (EB == 4) *reinterpret_cast<uint32_t*>(dst + ch * 4) = *reinterpret_cast<const uint32_t*>(src);
            else if (EB == 2) *reinterpret_cast<uint16_t*>(dst + ch * 2) = *reinterpret_cast<const uint16_t*>(src);
            else { dst[ch * 3] = src[0]; dst[ch * 3 + 1] = src[1]; dst[ch * 3 + 2] = src[2]; }
        }
    }
}

template <int N>
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, const float* in, int iters, long long* cyc) {
    float2 acc[N], w[2 * N];
    float h[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) w[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 64]);
#pragma unroll
    for (int j = 0; j < N; ++j) h[j] = in[j + 32];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) afp::ffma2(acc[i], w[i + N - j], h[j]);
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" {

const char* afp_version(void) { return "afp-b200 0.1 (sm_100a)"; }

const char* afp_last_error(const afp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int afp_create(const afp_config* cfg, afp_ctx** out_ctx) {
    if (!cfg || !out_ctx) return fail(nullptr, AFP_ERR_INVALID, "null argument");
    *out_ctx = nullptr;
    if (cfg->n_streams < 1) return fail(nullptr, AFP_ERR_INVALID, "n_streams must be >= 1");
    if (cfg->max_frames < 1) return fail(nullptr, AFP_ERR_INVALID, "max_frames must be >= 1");
    if (cfg->upsample < 1 || cfg->upsample > 4) return fail(nullptr, AFP_ERR_INVALID, "upsample must be 1..4");
    if (cfg->chain != AFP_CHAIN_DAC && cfg->chain != AFP_CHAIN_REFERENCE) return fail(nullptr, AFP_ERR_INVALID, "unknown chain");
    if (cfg->out_format < 0 || cfg->out_format > 3) return fail(nullptr, AFP_ERR_INVALID, "unknown out_format");
    if (cfg->dither < 0 || cfg->dither > 2) return fail(nullptr, AFP_ERR_INVALID, "unknown dither mode");
    if (cfg->eq_mode < 0 || cfg->eq_mode > 5) return fail(nullptr, AFP_ERR_INVALID, "unknown eq_mode");
    if (cfg->decimate < 0 || cfg->decimate > 64) return fail(nullptr, AFP_ERR_INVALID, "decimate must be 0..64");
    if (cfg->dither_bits != 0 && (cfg->dither_bits < 2 || cfg->dither_bits > 32))
        return fail(nullptr, AFP_ERR_INVALID, "dither_bits must be 0 or 2..32");
    if (cfg->eq_mode != AFP_EQ_NONE && cfg->chain == AFP_CHAIN_DAC && cfg->upsample > 1)
        return fail(nullptr, AFP_ERR_INVALID, "EQ runs at the processing rate: use AFP_CHAIN_REFERENCE with upsample > 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceCount");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, AFP_ERR_INVALID, "no such CUDA device");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10) return fail(nullptr, AFP_ERR_INVALID, "this library is built for sm_100a (B200) only");
    for (int U = 1; U <= 4; ++U)
        for (int ti = 0; ti < 3 && kTiles[U][ti]; ++ti)
            for (int nt = 128; nt <= 256; nt <<= 1)
                for (int multi = 0; multi <= 1; ++multi)
                    for (int epi : kAllEpi) {
                        kern_t k = kernel_for(U, kTiles[U][ti], nt, multi != 0, epi);
                        if (!k) continue;
                        e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
                        if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem)");
                    }
    for (int U : {1, 2, 3, 4})
        for (const RowShape& sh : kRowShapes)
            for (int epi : kAllEpi) {
                for (int spec : {0, sh.tg_log2}) {
                    kern_t k = row_kernel_for(U, sh.T, sh.pad, sh.store, epi, spec);
                    if (!k) continue;
                    e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
                    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, single-row kernel)");
                }
            }
    for (int row : {32, 64})
        for (int nts : {1, 2, 4})
            for (int epi : kAllEpi) {
                afp::tc_kern_t k = afp_kernel_tc(row, nts, epi);
                if (!k) continue;
                e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
                if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, tensor path)");
            }
    for (int U = 1; U <= 4; ++U)
        for (int ti = 0; ti < 3 && kTiles[U][ti]; ++ti) {
            kern_t k = afp_kernel_grouped(U, kTiles[U][ti]);
            if (!k) continue;
            e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, grouped kernel)");
        }
    for (int epi : kAllEpi) {
        for (int up : {1, 2, 3, 4}) {
            e = cudaFuncSetAttribute((const void*)afp_kernel_tc2(epi, up), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, split-FP16 tensor path)");
        }
    }
    e = cudaFuncSetAttribute((const void*)afp_kernel_tc2(afp::epi_code(afp::FMT_F32, afp::DITHER_RPDF), 1), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, split-FP16 tensor path, float dither)");
    e = cudaFuncSetAttribute((const void*)afp_kernel_tc2_grouped(), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, split-FP16 tensor path, grouped)");
    for (int variant : {afp::AGC_FAST, afp::AGC_EXACT, afp::AGC_CHECKED, afp::AGC_PROBE_F32}) {
        e = cudaFuncSetAttribute((const void*)afp_kernel_agc(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFuncSetAttribute(max dynamic smem, AGC)");
    }
    afp_ctx* c = new (std::nothrow) afp_ctx();
    if (!c) return fail(nullptr, AFP_ERR_NOMEM, "out of host memory");
    c->cfg = *cfg;
    c->n_sm = prop.multiProcessorCount;
    c->decim = cfg->decimate > 0 ? cfg->decimate : (cfg->chain == AFP_CHAIN_REFERENCE ? cfg->upsample : 1);
    c->dither_bits = cfg->dither_bits > 0 ? cfg->dither_bits : 24;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    e = cudaMalloc(&c->range_flag, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->range_flag, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaStreamDestroy(c->own_stream); delete c; return cuda_fail(nullptr, e, "range flag allocation"); }
    *out_ctx = c;
    return AFP_OK;
}

int afp_destroy(afp_ctx* c) {
    if (!c) return AFP_OK;
    cudaSetDevice(c->cfg.device);
    free_stage(c->up); free_stage(c->eq); free_stage(c->mainf); free_stage(c->fused);
    if (c->tmp_agc) cudaFree(c->tmp_agc);
    if (c->range_flag) cudaFree(c->range_flag);
    if (c->d_group_of_stream) cudaFree(c->d_group_of_stream);
    if (c->d_gtaps) cudaFree(c->d_gtaps);
    if (c->d_gcores) cudaFree(c->d_gcores);
    for (auto& kv : c->spec_tables) cudaFree(kv.second);
    if (c->d_il_in) cudaFree(c->d_il_in);
    if (c->d_il_out) cudaFree(c->d_il_out);
    if (c->tmp_a) cudaFree(c->tmp_a);
    if (c->tmp_b) cudaFree(c->tmp_b);
    for (int i = 0; i < 2; ++i) if (c->xhist[i]) cudaFree(c->xhist[i]);
    for (int i = 0; i < afp_ctx::kSlots; ++i) {
        if (c->d_in[i]) cudaFree(c->d_in[i]);
        if (c->d_out[i]) cudaFree(c->d_out[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
    return AFP_OK;
}

// The reference callback's chain (stream_process.py:96-107): up-sample by U -> FIR at fs*U -> processed[::U].
// With zs = zero-stuffed input, y[nU] = sum_k h[k] sum_l h_up[l] zs[nU-k-l] = sum_m g[m] zs[nU-m], g = h_up (*) h,
// and zs is non-zero only where m = jU:   y_dec[n] = sum_j g[jU] x[n-j].
// So the whole chain is one base-rate FIR of ceil((N+L-1)/U) taps: one launch instead of two, no
// processing-rate intermediate, ~U^2 fewer multiply-adds than computing every processing-rate sample and
// dropping U-1 of U.  Exact in exact arithmetic; in float32 it skips the rounding of the up-sampled
// intermediate (the staged form stays available: AFP_NO_COLLAPSE=1, and is what runs with an EQ stage or
// a decimation factor other than U).
// The same holds with the streaming fused EQ (one table sum_b g_b h_b, AFP_EQ_FUSED) in the chain: it is one
// more LTI stage at the processing rate, so g = h_up (*) eq (*) h; and with U = 1 the EQ and the main FIR are
// one filter eq (*) h (one launch, no intermediate, and >= 400 combined taps reach the tensor path).  Not
// done for the DAC chain with U > 1: there the EQ runs at the base rate and folding it into the polyphase
// filter would multiply its cost by U.  keep_state: a gain change keeps the carried history (it holds input
// samples only, so it is valid for any taps); a new filter or EQ bank resets it like every other stage.
static std::vector<double> convolve(const std::vector<double>& a, const std::vector<double>& b) {
    std::vector<double> g(a.size() + b.size() - 1, 0.0);
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t k = 0; k < b.size(); ++k) g[i + k] += a[i] * b[k];
    return g;
}

static int rebuild_fused(afp_ctx* c, bool keep_state = false) {
    c->groups_dirty = !c->group_of_stream.empty();   // every filter change goes through here: per-group tables follow lazily
    const int U = c->cfg.upsample;
    const bool ref_up = c->cfg.chain == AFP_CHAIN_REFERENCE && U > 1;
    const bool eq_ok = c->cfg.eq_mode == AFP_EQ_NONE || (c->cfg.eq_mode == AFP_EQ_FUSED && c->eq.active);
    const bool shape_ok = ref_up ? (c->decim == U && c->up.active) : (U == 1 && c->decim == 1 && c->cfg.eq_mode == AFP_EQ_FUSED);
    const bool want = eq_ok && shape_ok && c->mainf.active && !env_flag("AFP_NO_COLLAPSE");
    // The collapsed form carries INPUT samples; so does the first stage of the staged form (interpolator, or the fused EQ
    // at U = 1).  A keep-state rebuild that changes form hands them over (right-aligned): collapsed -> staged seeds the
    // first stage and clears the later ones (their state is intermediate signal that was never computed - a restart
    // transient of the later stages' length); staged -> collapsed is exact.
    Stage* first = ref_up ? &c->up : &c->eq;
    auto leave_collapsed = [&]() -> int {
        if (keep_state && c->fused.active && c->fused.hist[c->fused.cur] && first->active && first->hist[0]) {
            const long long k = std::min(c->fused.hist_pitch, first->hist_pitch);
            AFP_CUDA(c, cudaDeviceSynchronize());
            for (Stage* s : {&c->up, &c->eq, &c->mainf})
                if (s->active)
                    for (int i = 0; i < 2; ++i)
                        if (s->hist[i]) AFP_CUDA(c, cudaMemset(s->hist[i], 0, (size_t)c->cfg.n_streams * s->hist_pitch * sizeof(float)));
            AFP_CUDA(c, cudaMemcpy2D(first->hist[first->cur] + (first->hist_pitch - k), (size_t)first->hist_pitch * sizeof(float),
                                     c->fused.hist[c->fused.cur] + (c->fused.hist_pitch - k), (size_t)c->fused.hist_pitch * sizeof(float),
                                     (size_t)k * sizeof(float), (size_t)c->cfg.n_streams, cudaMemcpyDeviceToDevice));
            AFP_CUDA(c, cudaDeviceSynchronize());
        }
        free_stage(c->fused);
        return AFP_OK;
    };
    if (!want) return leave_collapsed();
    std::vector<double> g = c->mainf.taps;
    if (c->cfg.eq_mode == AFP_EQ_FUSED) g = convolve(c->eq.taps, g);
    if (ref_up) g = convolve(c->up.taps, g);
    const int step = ref_up ? U : 1;
    std::vector<double> g0((g.size() + step - 1) / step);
    for (size_t j = 0; j < g0.size(); ++j) g0[j] = g[j * step];
    if ((long long)g0.size() + 16 > afp::MAX_TAPS) return leave_collapsed();   // too long for one table: stay staged
    // two tensor-kernel stages beat one FFMA2 stage: do not merge EQ and main FIR into a filter that is too
    // long for the tensor kernel's shared memory (16 streams x [history | 256 frames], hi + lo, + tap ring)
    if (!ref_up) {
        const long long kq64 = ((long long)g0.size() + 63) / 64 * 64;
        if (2 * afp::TC_STREAMS * (kq64 + 256) * 4 + 2 * 8192 + 512 > kMaxSmem) return leave_collapsed();
    }
    const bool entering = keep_state && !c->fused.active && first->active && first->hist[first->cur];
    int rc = build_stage(c, c->fused, ST_MAIN, 1, g0.data(), (int)g0.size(), 1, keep_state);
    if (rc || !entering) return rc;
    // staged -> collapsed with the state kept: the first stage's carried input samples become the collapsed filter's
    const long long k = std::min(c->fused.hist_pitch, first->hist_pitch);
    AFP_CUDA(c, cudaDeviceSynchronize());
    AFP_CUDA(c, cudaMemcpy2D(c->fused.hist[c->fused.cur] + (c->fused.hist_pitch - k), (size_t)c->fused.hist_pitch * sizeof(float),
                             first->hist[first->cur] + (first->hist_pitch - k), (size_t)first->hist_pitch * sizeof(float),
                             (size_t)k * sizeof(float), (size_t)c->cfg.n_streams, cudaMemcpyDeviceToDevice));
    AFP_CUDA(c, cudaDeviceSynchronize());
    return AFP_OK;
}

// ---- per-stream-group filters ------------------------------------------------------------------------------------
// The one stage of a chain that is a single filter, AGC or not ahead of it (see single_stage below for the launch side).
static Stage* grouped_stage(afp_ctx* c) {
    if (c->fused.active) return &c->fused;
    if (c->cfg.chain == AFP_CHAIN_REFERENCE && c->cfg.upsample > 1) return nullptr;
    if (c->eq.active || c->cfg.eq_mode != AFP_EQ_NONE) return nullptr;
    return c->mainf.active ? &c->mainf : nullptr;
}

// EQ table sum_b g_b h_b with the bands aligned on their 'same' centres (the layout of rebuild_eq, fused form)
static std::vector<double> eq_table_for(const afp_ctx* c, const double* gains, int only_band = -1) {
    int cmax = 0, nall = 0;
    for (auto& b : c->bands) cmax = std::max(cmax, ((int)b.size() - 1) / 2);
    for (auto& b : c->bands) nall = std::max(nall, cmax - ((int)b.size() - 1) / 2 + (int)b.size());
    std::vector<double> tab(nall, 0.0);
    for (size_t i = 0; i < c->bands.size(); ++i) {
        if (only_band >= 0 && (int)i != only_band) continue;
        const int off = cmax - ((int)c->bands[i].size() - 1) / 2;
        const double g = only_band >= 0 ? 1.0 : gains[i];
        for (size_t k = 0; k < c->bands[i].size(); ++k) tab[off + k] += g * c->bands[i][k];
    }
    return tab;
}

// Taps of the chain's single stage for a main filter and an EQ table: what rebuild_fused computes for the context's own.
static std::vector<double> single_stage_taps(const afp_ctx* c, const std::vector<double>& main, const std::vector<double>* eq) {
    if (!c->fused.active) return main;
    const int U = c->cfg.upsample;
    const bool ref_up = c->cfg.chain == AFP_CHAIN_REFERENCE && U > 1;
    std::vector<double> g = main;
    if (c->cfg.eq_mode == AFP_EQ_FUSED && eq) g = convolve(*eq, g);
    if (ref_up) g = convolve(c->up.taps, g);
    const int step = ref_up ? U : 1;
    std::vector<double> g0((g.size() + step - 1) / step);
    for (size_t j = 0; j < g0.size(); ++j) g0[j] = g[j * step];
    return g0;
}

static int rebuild_groups(afp_ctx* c) {
    c->groups_dirty = false;
    if (c->group_of_stream.empty()) return AFP_OK;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    Stage* s = grouped_stage(c);
    if (!s || s->n_sub != 1)
        return fail(c, AFP_ERR_STATE, "per-group filters need a chain that is ONE filter: eq_mode none or fused (U = 1 or the collapsed "
                                      "reference chain)");
    const int G = (int)c->groups.size(), n = s->n_taps;
    const bool eq = c->fused.active && c->cfg.eq_mode == AFP_EQ_FUSED;
    const size_t stride = (size_t)s->kq_pad * s->U;
    AFP_CUDA(c, cudaDeviceSynchronize());      // launches in flight still read the old tables
    if (c->d_gtaps) { cudaFree(c->d_gtaps); c->d_gtaps = nullptr; }
    if (c->d_gcores) { cudaFree(c->d_gcores); c->d_gcores = nullptr; }
    AFP_CUDA(c, cudaMalloc(&c->d_gtaps, (size_t)G * stride * sizeof(float)));
    c->gtaps_stride = (long long)stride;
    const bool cores = s->tc2_w != nullptr;
    const size_t cstride = cores ? (size_t)2 * s->tc2_cores * 64 : 0;
    if (cores) AFP_CUDA(c, cudaMalloc(&c->d_gcores, (size_t)G * cstride * sizeof(__half)));
    c->gcores_stride = (long long)cstride;
    // sessions that only differ in their gains: the stage filter is linear in them, sum_b g_b F_b with F_b the stage
    // filter of band b alone - ten convolutions for the whole batch instead of one per group
    std::vector<std::vector<double>> band_stage;
    if (eq)
        for (size_t b = 0; b < c->bands.size(); ++b) {
            const std::vector<double> e1 = eq_table_for(c, nullptr, (int)b);
            band_stage.push_back(single_stage_taps(c, c->mainf.taps, &e1));
        }
    const int chunk = 64;
    std::vector<float> tab((size_t)chunk * stride);
    std::vector<__half> cw, cbuf;
    if (cores) cbuf.resize((size_t)chunk * cstride);
    for (int g0 = 0; g0 < G; g0 += chunk) {
        const int gn = std::min(chunk, G - g0);
        std::fill(tab.begin(), tab.end(), 0.f);
        for (int gi = 0; gi < gn; ++gi) {
            const afp_ctx::GroupDef& gd = c->groups[g0 + gi];
            const double* gains = gd.gains.empty() ? c->gains.data() : gd.gains.data();
            std::vector<double> t;
            if (eq && gd.taps.empty()) {
                t.assign(band_stage[0].size(), 0.0);
                for (size_t b = 0; b < band_stage.size(); ++b)
                    for (size_t k = 0; k < t.size(); ++k) t[k] += gains[b] * band_stage[b][k];
            } else {
                const std::vector<double>& main = gd.taps.empty() ? c->mainf.taps : gd.taps;
                if (eq) { const std::vector<double> e = eq_table_for(c, gains); t = single_stage_taps(c, main, &e); }
                else t = single_stage_taps(c, main, nullptr);
            }
            if ((int)t.size() != n)
                return fail(c, AFP_ERR_INVALID, "a group's filter has another length than the context's: all groups share one stage shape");
            for (int k = 0; k < n; ++k) tab[(size_t)gi * stride + k] = (float)t[k];
            if (cores) {
                afp::tc2_build_cores(t.data(), n, cw);
                std::memcpy(cbuf.data() + (size_t)gi * cstride, cw.data(), cstride * sizeof(__half));
            }
        }
        AFP_CUDA(c, cudaMemcpy(c->d_gtaps + (size_t)g0 * stride, tab.data(), (size_t)gn * stride * sizeof(float), cudaMemcpyHostToDevice));
        if (cores)
            AFP_CUDA(c, cudaMemcpy(static_cast<__half*>(c->d_gcores) + (size_t)g0 * cstride, cbuf.data(), (size_t)gn * cstride * sizeof(__half),
                                   cudaMemcpyHostToDevice));
    }
    if (!c->d_group_of_stream) AFP_CUDA(c, cudaMalloc(&c->d_group_of_stream, (size_t)c->cfg.n_streams * sizeof(int)));
    AFP_CUDA(c, cudaMemcpy(c->d_group_of_stream, c->group_of_stream.data(), (size_t)c->cfg.n_streams * sizeof(int), cudaMemcpyHostToDevice));
    AFP_CUDA(c, cudaDeviceSynchronize());
    s->geo = Geometry();
    return AFP_OK;
}

int afp_assign_groups(afp_ctx* c, const int32_t* group_of_stream, int32_t n_streams) {
    if (!c) return AFP_ERR_INVALID;
    if (!group_of_stream) {   // back to one filter set for the whole batch
        c->group_of_stream.clear(); c->groups.clear(); c->groups_dirty = false;
        for (Stage* s : {&c->fused, &c->mainf}) s->geo = Geometry();
        return AFP_OK;
    }
    if (n_streams != c->cfg.n_streams) return fail(c, AFP_ERR_INVALID, "afp_assign_groups needs one entry per stream");
    int gmax = 0;
    for (int i = 0; i < n_streams; ++i) {
        if (group_of_stream[i] < 0 || group_of_stream[i] >= (1 << 20)) return fail(c, AFP_ERR_INVALID, "group index outside 0..2^20-1");
        gmax = std::max(gmax, (int)group_of_stream[i]);
    }
    c->group_of_stream.assign(group_of_stream, group_of_stream + n_streams);
    c->groups.resize((size_t)gmax + 1);
    c->groups_pair_aligned = true;
    for (int i = 0; i + 1 < n_streams; i += 2)
        if (group_of_stream[i] != group_of_stream[i + 1]) c->groups_pair_aligned = false;
    c->groups_dirty = true;
    return AFP_OK;
}

int afp_set_gains_group(afp_ctx* c, int32_t group, const double* gains, int32_t n_bands) {
    if (!c) return AFP_ERR_INVALID;
    if (group < 0 || group >= (int)c->groups.size()) return fail(c, AFP_ERR_INVALID, "no such group (afp_assign_groups first)");
    if (!gains) { c->groups[group].gains.clear(); c->groups_dirty = true; return AFP_OK; }
    if ((size_t)n_bands != c->bands.size()) return fail(c, AFP_ERR_INVALID, "gain count does not match the EQ bank");
    for (int b = 0; b < n_bands; ++b)
        if (!std::isfinite(gains[b])) return fail(c, AFP_ERR_INVALID, "non-finite gain");
    c->groups[group].gains.assign(gains, gains + n_bands);
    c->groups_dirty = true;
    return AFP_OK;
}

int afp_set_taps_group(afp_ctx* c, int32_t group, const double* taps, int32_t n_taps) {
    if (!c) return AFP_ERR_INVALID;
    if (group < 0 || group >= (int)c->groups.size()) return fail(c, AFP_ERR_INVALID, "no such group (afp_assign_groups first)");
    if (!taps) { c->groups[group].taps.clear(); c->groups_dirty = true; return AFP_OK; }
    if (!c->mainf.active || n_taps != c->mainf.n_taps)
        return fail(c, AFP_ERR_INVALID, "a group's main filter must have the length of the context's (afp_set_taps)");
    for (int i = 0; i < n_taps; ++i)
        if (!std::isfinite(taps[i])) return fail(c, AFP_ERR_INVALID, "non-finite filter tap");
    c->groups[group].taps.assign(taps, taps + n_taps);
    c->groups_dirty = true;
    return AFP_OK;
}

int afp_set_taps(afp_ctx* c, const double* taps, int32_t n) {
    if (!c || !taps) return fail(c, AFP_ERR_INVALID, "null argument");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    const int U = c->cfg.chain == AFP_CHAIN_DAC ? c->cfg.upsample : 1;
    int rc = build_stage(c, c->mainf, ST_MAIN, U, taps, n);
    c->xhist_reset = true;
    return rc ? rc : rebuild_fused(c);
}

int afp_swap_taps(afp_ctx* c, const double* taps, int32_t n) {
    if (!c || !taps) return fail(c, AFP_ERR_INVALID, "null argument");
    if (!c->mainf.active) return fail(c, AFP_ERR_STATE, "afp_set_taps has not been called");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    // blocks in flight still read the current history rows - on the context's own streams or on whatever stream the
    // caller passed to afp_process, which the context does not know: hence the device-wide wait (a swap is a GUI event)
    AFP_CUDA(c, cudaDeviceSynchronize());
    const int U = c->cfg.chain == AFP_CHAIN_DAC ? c->cfg.upsample : 1;
    int rc = build_stage(c, c->mainf, ST_MAIN, U, taps, n, 1, true);
    c->xhist_migrate = true;
    return rc ? rc : rebuild_fused(c, true);
}

int afp_set_upsampler(afp_ctx* c, const double* taps, int32_t n) {
    if (!c || !taps) return fail(c, AFP_ERR_INVALID, "null argument");
    if (c->cfg.chain != AFP_CHAIN_REFERENCE || c->cfg.upsample < 2)
        return fail(c, AFP_ERR_STATE, "an up-sampler prototype is only used by AFP_CHAIN_REFERENCE with upsample > 1");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    int rc = build_stage(c, c->up, ST_INTERP, c->cfg.upsample, taps, n);
    return rc ? rc : rebuild_fused(c);
}

int afp_set_eq(afp_ctx* c, const double* band_taps, const int32_t* band_len, int32_t n_bands, const double* gains) {
    if (!c) return fail(c, AFP_ERR_INVALID, "null argument");
    if (c->cfg.eq_mode == AFP_EQ_NONE) return fail(c, AFP_ERR_STATE, "context was created with eq_mode = AFP_EQ_NONE");
    if (n_bands < 0 || (n_bands > 0 && (!band_taps || !band_len))) return fail(c, AFP_ERR_INVALID, "bad EQ bank");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    c->bands.clear(); c->gains.clear();
    const double* p = band_taps;
    for (int b = 0; b < n_bands; ++b) {
        if (band_len[b] < 1) return fail(c, AFP_ERR_INVALID, "EQ band needs at least one tap");
        c->bands.emplace_back(p, p + band_len[b]);
        p += band_len[b];
        c->gains.push_back(gains ? gains[b] : 1.0);
    }
    int rc = rebuild_eq(c, true);
    return rc ? rc : rebuild_fused(c, false);
}

int afp_set_gains(afp_ctx* c, const double* gains, int32_t n_bands) {
    if (!c || !gains) return fail(c, AFP_ERR_INVALID, "null argument");
    if ((size_t)n_bands != c->bands.size()) return fail(c, AFP_ERR_INVALID, "gain count does not match the EQ bank");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    for (int b = 0; b < n_bands; ++b) {
        if (!std::isfinite(gains[b])) return fail(c, AFP_ERR_INVALID, "non-finite gain");
        c->gains[b] = gains[b];
    }
    int rc = rebuild_eq(c, false);
    return rc ? rc : rebuild_fused(c, true);
}

int afp_set_agc(afp_ctx* c, const afp_agc_params* p) {
    if (!c) return AFP_ERR_INVALID;
    if (!p) { c->agc_on = false; return AFP_OK; }
    if (!(p->max_gain == p->max_gain) || !(p->target_level == p->target_level))
        return fail(c, AFP_ERR_INVALID, "AGC: NaN parameter");
    c->agc = *p;
    c->agc_on = true;
    return AFP_OK;
}

int afp_agc_block(afp_ctx* c, const afp_agc_params* p, const float* in, int64_t in_pitch, float* out, int64_t out_pitch,
                  int32_t n_streams, int32_t frames, void* cuda_stream) {
    if (!c) return AFP_ERR_INVALID;
    if (!p || !in || !out) return fail(c, AFP_ERR_INVALID, "null argument");
    if (n_streams < 0 || frames < 0) return fail(c, AFP_ERR_INVALID, "negative size");
    if (in_pitch < frames || out_pitch < frames) return fail(c, AFP_ERR_INVALID, "pitch smaller than frames");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    return launch_agc(c, *p, in, in_pitch, out, out_pitch, n_streams, frames, (cudaStream_t)cuda_stream);
}

int afp_spectrum_db(afp_ctx* c, const float* in, int64_t in_pitch, float* out, int64_t out_pitch, int32_t n_streams,
                    int32_t frames, void* cuda_stream) {
    if (!c) return AFP_ERR_INVALID;
    if (!in || !out) return fail(c, AFP_ERR_INVALID, "null buffer");
    if (n_streams < 0) return fail(c, AFP_ERR_INVALID, "negative size");
    int log2n = 0;
    while ((1 << log2n) < frames) ++log2n;
    if (frames < 64 || frames > 4096 || (1 << log2n) != frames)
        return fail(c, AFP_ERR_INVALID, "spectrum: frames must be a power of two in 64..4096");
    if (in_pitch < frames || out_pitch < frames / 2 + 1) return fail(c, AFP_ERR_INVALID, "pitch smaller than the row");
    if (n_streams == 0) return AFP_OK;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    float*& tab = c->spec_tables[frames];
    if (!tab) {
        // np.hanning(n) = 0.5 - 0.5 cos(2 pi k / (n - 1)); twiddles exp(-2 pi i k / n): float64 on the host
        std::vector<float> h((size_t)2 * frames);
        const double pi = 3.14159265358979323846;
        for (int k = 0; k < frames; ++k) h[k] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * k / (frames - 1)));
        for (int k = 0; k < frames / 2; ++k) {
            h[frames + 2 * k] = (float)std::cos(-2.0 * pi * k / frames);
            h[frames + 2 * k + 1] = (float)std::sin(-2.0 * pi * k / frames);
        }
        AFP_CUDA(c, cudaMalloc(&tab, h.size() * sizeof(float)));
        AFP_CUDA(c, cudaMemcpy(tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    afp::SpectrumParams P{};
    P.in = in; P.in_pitch = in_pitch; P.out = out; P.out_pitch = out_pitch;
    P.window = tab; P.twiddle = reinterpret_cast<const float2*>(tab + frames);
    P.n_streams = n_streams; P.n = frames; P.log2n = log2n;
    if (frames == afp::SPW_N && !env_flag("AFP_SPEC_NO_WARP")) {
        // one warp per row, radix-8 passes in registers (afp_spectrum.cuh): the GUI's default block length
        if (!c->spw_per_sm) {
            int per_sm = 0;
            AFP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)afp::spectrum_db_warp_kernel,
                                                                      32 * afp::SPW_WARPS, afp::SPW_SMEM));
            c->spw_per_sm = per_sm < 1 ? 1 : per_sm;
        }
        long long grid_w = (long long)c->n_sm * c->spw_per_sm;
        const long long want = (n_streams + afp::SPW_WARPS - 1) / afp::SPW_WARPS;
        if (grid_w > want) grid_w = want;
        afp::spectrum_db_warp_kernel<<<(int)grid_w, 32 * afp::SPW_WARPS, afp::SPW_SMEM, (cudaStream_t)cuda_stream>>>(P);
        AFP_CUDA(c, cudaGetLastError());
        c->launches++;
        return AFP_OK;
    }
    const size_t smem = (size_t)(afp::spec_pad(frames / 2) + frames / 4) * sizeof(float2);   // padded packed samples + twiddle copy
    const int per_sm = frames <= 1024 ? 8 : (frames <= 2048 ? 4 : 2);
    int grid = c->n_sm * per_sm;
    if (grid > n_streams) grid = n_streams;
    afp::spectrum_db_kernel<<<grid, afp::SPEC_THREADS, smem, (cudaStream_t)cuda_stream>>>(P);
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
    return AFP_OK;
}

int32_t afp_out_frames(const afp_ctx* c, int32_t frames) {
    if (!c) return 0;
    const long long n = (long long)frames * c->cfg.upsample;   // both chains compute at fs*U
    return (int32_t)((n + c->decim - 1) / c->decim);
}

int32_t afp_out_elem_bytes(const afp_ctx* c) {
    if (!c) return 0;
    return c->cfg.out_format == AFP_FMT_I16 ? 2 : (c->cfg.out_format == AFP_FMT_I24_PACKED ? 3 : 4);
}

// One block for streams [s0, s0+ns); `in` / `out` point at stream s0.  Does not advance the state.
static int process_block_range(afp_ctx* c, const float* in, int64_t in_pitch, void* out, int64_t out_pitch,
                               int32_t frames, int s0, int ns, cudaStream_t st) {
    const int U = c->cfg.upsample;
    const bool ref = c->cfg.chain == AFP_CHAIN_REFERENCE;
    const int mode = c->cfg.eq_mode;
    int rc;
    if (c->agc_on) {   // gain-controlled block replaces the input of everything below (stream_process_AGC.py:131-146)
        float* tg = c->tmp_agc + (long long)s0 * c->pitch_agc;
        rc = launch_agc(c, c->agc, in, in_pitch, tg, c->pitch_agc, ns, frames, st);
        if (rc) return rc;
        in = tg; in_pitch = c->pitch_agc;
    }
    const float* cur = in; long long cur_pitch = in_pitch; int cur_len = frames;
    if (c->fused.active)  // interpolate -> FIR -> decimate collapsed into one base-rate FIR (rebuild_fused)
        return run_streaming(c, c->fused, in, in_pitch, frames, out,
                             c->cfg.out_format == AFP_FMT_I24_PACKED ? out_pitch / 3 : out_pitch,
                             c->cfg.out_format, c->cfg.dither, 1, c->out_samples, s0, ns, st);
    if (ref && U > 1) {   // interpolate to the processing rate (replaces soxr.resample, stream_process.py:66-77)
        float* ta = c->tmp_a + (long long)s0 * c->pitch_a;
        rc = run_streaming(c, c->up, cur, cur_pitch, cur_len, ta, c->pitch_a, AFP_FMT_F32, AFP_DITHER_NONE, 1, 0, s0, ns, st);
        if (rc) return rc;
        cur = ta; cur_pitch = c->pitch_a; cur_len = frames * U;
    }
    const long long opitch = c->cfg.out_format == AFP_FMT_I24_PACKED ? out_pitch / 3 : out_pitch;
    float* tb = c->tmp_b ? c->tmp_b + (long long)s0 * c->pitch_b : nullptr;

    if (eq_is_buffer_same(c)) {
        // stream_process_EQ.py:119-134.  Launch 1: eq[i] = sum_b g_b * same(buffer (*) h_b)[i] over the
        // whole sliding buffer [last N-1 samples | block], zero-extended; the buffer rolls forward.
        const int hm = c->xhist_len, lb = hm + cur_len;
        StageIO io;
        io.s0 = s0; io.ns = ns;
        io.src0 = c->xhist[c->xhist_cur] + (long long)s0 * c->xhist_pitch; io.pitch0 = c->xhist_pitch; io.len0 = hm;
        io.src1 = cur; io.pitch1 = cur_pitch; io.len1 = cur_len;
        io.q_begin = c->eq_centre; io.n_q = lb;
        io.out = tb; io.out_pitch = c->pitch_b;
        io.hist_out = c->xhist[c->xhist_cur ^ 1] + (long long)s0 * c->xhist_pitch; io.hist_pitch = c->xhist_pitch;
        io.hist_len = hm; io.hist_from = cur_len;
        rc = launch_stage(c, c->eq, io, st);
        if (rc) return rc;
        // Launch 2: 'valid' main FIR over eq (:134), decimation and epilogue
        StageIO m;
        m.s0 = s0; m.ns = ns;
        m.src1 = tb; m.pitch1 = c->pitch_b; m.len1 = lb;
        // 'valid' outputs start N-1 samples into the buffer; with the default buffer that is also the
        // start of the new block, with a longer fixed buffer these are the oldest outputs (:450-453)
        m.q_begin = c->mainf.n_taps - 1; m.n_q = cur_len;
        m.out = out; m.out_pitch = opitch; m.fmt = c->cfg.out_format; m.dither = c->cfg.dither; m.decim = c->decim;
        m.sample_base = c->out_samples;
        return launch_stage(c, c->mainf, m, st);
    }
    if (mode == AFP_EQ_STATELESS_SAME) {
        // stream_process_EQ_GUI.py:86-95: both filters 'same' over the block alone, nothing carried
        if (c->eq.active) {
            StageIO io;
            io.s0 = s0; io.ns = ns;
            io.src1 = cur; io.pitch1 = cur_pitch; io.len1 = cur_len;
            io.q_begin = c->eq_centre; io.n_q = cur_len;
            io.out = tb; io.out_pitch = c->pitch_b;
            rc = launch_stage(c, c->eq, io, st);
            if (rc) return rc;
            cur = tb; cur_pitch = c->pitch_b;
        }
        StageIO m;
        m.s0 = s0; m.ns = ns;
        m.src1 = cur; m.pitch1 = cur_pitch; m.len1 = cur_len;
        m.q_begin = (c->mainf.n_taps - 1) / 2; m.n_q = cur_len;
        m.out = out; m.out_pitch = opitch; m.fmt = c->cfg.out_format; m.dither = c->cfg.dither; m.decim = c->decim;
        m.sample_base = c->out_samples;
        return launch_stage(c, c->mainf, m, st);
    }
    if (c->eq.active) {   // streaming EQ (fused or per band), state carried like any other FIR
        rc = run_streaming(c, c->eq, cur, cur_pitch, cur_len, tb, c->pitch_b, AFP_FMT_F32, AFP_DITHER_NONE, 1, 0, s0, ns, st);
        if (rc) return rc;
        cur = tb; cur_pitch = c->pitch_b;
    }
    return run_streaming(c, c->mainf, cur, cur_pitch, cur_len, out, opitch, c->cfg.out_format, c->cfg.dither, c->decim,
                         c->out_samples, s0, ns, st);
}

// the whole block has been launched: roll the ping-pong histories and the dither counter
static void finish_block(afp_ctx* c, int32_t frames) {
    const bool ref = c->cfg.chain == AFP_CHAIN_REFERENCE;
    const int mode = c->cfg.eq_mode;
    if (c->fused.active) {
        c->fused.cur ^= 1;
        c->out_samples += (unsigned long long)afp_out_frames(c, frames);
        return;
    }
    if (ref && c->cfg.upsample > 1) c->up.cur ^= 1;
    if (eq_is_buffer_same(c)) c->xhist_cur ^= 1;
    else if (mode != AFP_EQ_STATELESS_SAME) {
        if (c->eq.active) c->eq.cur ^= 1;
        c->mainf.cur ^= 1;
    }
    c->out_samples += (unsigned long long)afp_out_frames(c, frames);
}

static int check_ready(afp_ctx* c, int32_t frames) {
    if (frames < 0 || frames > c->cfg.max_frames) return fail(c, AFP_ERR_INVALID, "frames outside 0..max_frames");
    if (!c->mainf.active) return fail(c, AFP_ERR_STATE, "afp_set_taps has not been called");
    if (c->cfg.chain == AFP_CHAIN_REFERENCE && c->cfg.upsample > 1 && !c->up.active)
        return fail(c, AFP_ERR_STATE, "afp_set_upsampler has not been called");
    const int mode = c->cfg.eq_mode;
    if (mode != AFP_EQ_NONE && mode != AFP_EQ_STATELESS_SAME && !c->eq.active)
        return fail(c, AFP_ERR_STATE, "afp_set_eq has not been called");
    if (eq_is_buffer_same(c) && c->cfg.same_buffer_len > 0 && frames != c->cfg.max_frames)
        return fail(c, AFP_ERR_INVALID, "a fixed same_buffer_len needs frames == max_frames on every call");
    if (((long long)frames * c->cfg.upsample) % c->decim != 0)
        return fail(c, AFP_ERR_INVALID, "frames * upsample is not a multiple of the decimation factor");
    if (c->groups_dirty) return rebuild_groups(c);
    if (!c->group_of_stream.empty() && !grouped_stage(c))
        return fail(c, AFP_ERR_STATE, "per-group filters need a chain that is ONE filter (eq_mode none or fused)");
    return AFP_OK;
}

int afp_process(afp_ctx* c, const float* in, int64_t in_pitch, void* out, int64_t out_pitch, int32_t frames, void* cuda_stream) {
    if (!c) return AFP_ERR_INVALID;
    if (!in || !out) return fail(c, AFP_ERR_INVALID, "null buffer");
    int rc = check_ready(c, frames);
    if (rc) return rc;
    if (frames == 0) return AFP_OK;
    if (in_pitch < frames) return fail(c, AFP_ERR_INVALID, "in_pitch smaller than frames");
    const int out_n = afp_out_frames(c, frames);
    if (out_pitch < (int64_t)out_n * (c->cfg.out_format == AFP_FMT_I24_PACKED ? 3 : 1))
        return fail(c, AFP_ERR_INVALID, "out_pitch smaller than the output row");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    rc = ensure_tmp(c);
    if (rc) return rc;
    rc = process_block_range(c, in, in_pitch, out, out_pitch, frames, 0, c->cfg.n_streams, (cudaStream_t)cuda_stream);
    if (rc) return rc;
    finish_block(c, frames);
    return AFP_OK;
}

// the one streaming stage of a chain that is a single filter (the collapsed reference chain, or a plain FIR
// with no interpolator / EQ / AGC ahead of it); nullptr for every other chain
static Stage* single_stage(afp_ctx* c) {
    if (c->agc_on) return nullptr;
    if (c->fused.active) return &c->fused;
    if (c->cfg.chain == AFP_CHAIN_REFERENCE && c->cfg.upsample > 1) return nullptr;
    if (c->eq.active || c->cfg.eq_mode != AFP_EQ_NONE) return nullptr;
    return c->mainf.active ? &c->mainf : nullptr;
}

// the split-FP16 tensor kernel saw a sample beyond FP16's range since the last check: report and clear
static int check_range(afp_ctx* c, bool force = false) {
    if (!c->range_dirty && !force) return AFP_OK;      // no split-FP16 launch since the last look: nothing can have set it
    c->range_dirty = false;
    unsigned int flag = 0;
    AFP_CUDA(c, cudaMemcpy(&flag, c->range_flag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (!flag) return AFP_OK;
    AFP_CUDA(c, cudaMemset(c->range_flag, 0, sizeof(flag)));
    AFP_CUDA(c, cudaDeviceSynchronize());
    return fail(c, AFP_ERR_RANGE, "an input sample exceeds +-65504 (the split-FP16 tensor path's range; audio is +-1): "
                                  "the affected outputs are not valid; AFP_NO_TC2=1 selects the 3xTF32 / FFMA2 kernels");
}

int afp_range_status(afp_ctx* c, int32_t* exceeded) {
    if (!c || !exceeded) return AFP_ERR_INVALID;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    const int rc = check_range(c, true);
    *exceeded = rc == AFP_ERR_RANGE ? 1 : 0;
    return rc == AFP_ERR_RANGE ? AFP_OK : rc;
}

int afp_process_blocks(afp_ctx* c, const float* in, int64_t in_pitch, int64_t in_block_stride, void* out, int64_t out_pitch,
                       int64_t out_block_stride, int32_t frames, int32_t n_blocks, void* cuda_stream) {
    if (!c) return AFP_ERR_INVALID;
    if (!in || !out) return fail(c, AFP_ERR_INVALID, "null buffer");
    if (n_blocks < 0) return fail(c, AFP_ERR_INVALID, "negative block count");
    int rc = check_ready(c, frames);
    if (rc) return rc;
    if (frames == 0 || n_blocks == 0) return AFP_OK;
    if (in_pitch < frames) return fail(c, AFP_ERR_INVALID, "in_pitch smaller than frames");
    const int out_n = afp_out_frames(c, frames);
    const int64_t unit = c->cfg.out_format == AFP_FMT_I24_PACKED ? 3 : 1;
    if (out_pitch < (int64_t)out_n * unit) return fail(c, AFP_ERR_INVALID, "out_pitch smaller than the output row");
    if (n_blocks > 1 && (in_block_stride < frames || out_block_stride < (int64_t)out_n * unit))
        return fail(c, AFP_ERR_INVALID, "block stride smaller than a row");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    rc = ensure_tmp(c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int S = c->cfg.n_streams;
    if (Stage* s = single_stage(c)) {
        // one launch for all the blocks: the CTA that owns a stream group walks its blocks with the window in
        // shared memory; the carried history is read once and written once
        bool done = false;
        // (the collapsed reference chain has its decimation folded into the base-rate filter: the stage itself runs with decim 1)
        if ((s->U == 1 || (s->U >= 2 && s->U <= 4 && c->group_of_stream.empty())) && (c->decim == 1 || s == &c->fused)) {
            auto launch = c->group_of_stream.empty() ? launch_stage_tc2 : launch_stage_tc2_grouped;
            rc = launch(c, *s, in, in_pitch, in_block_stride, out, out_pitch / unit, out_block_stride / unit, frames, n_blocks,
                        c->cfg.out_format, c->cfg.dither, c->out_samples, 0, S, st, &done);
            if (rc) return rc;
        }
        if (done) {
            s->cur ^= 1;
            c->out_samples += (unsigned long long)out_n * (unsigned long long)n_blocks;
            return AFP_OK;
        }
    }
    const int eb = afp_out_elem_bytes(c);
    for (int b = 0; b < n_blocks; ++b) {
        rc = process_block_range(c, in + (int64_t)b * in_block_stride, in_pitch,
                                 static_cast<unsigned char*>(out) + (int64_t)b * out_block_stride * (eb == 3 ? 1 : eb), out_pitch, frames, 0, S, st);
        if (rc) return rc;
        finish_block(c, frames);
    }
    return AFP_OK;
}

static int ensure_host_staging(afp_ctx* c) {
    if (c->d_in[0]) return AFP_OK;
    const int S = c->cfg.n_streams;
    const int eb = afp_out_elem_bytes(c);
    const long long max_out = (long long)afp_out_frames(c, c->cfg.max_frames);
    for (int i = 0; i < afp_ctx::kSlots; ++i) {   // each slot can hold one whole block of all streams
        AFP_CUDA(c, cudaMalloc(&c->d_in[i], (size_t)S * c->cfg.max_frames * sizeof(float)));
        AFP_CUDA(c, cudaMalloc(&c->d_out[i], (size_t)S * max_out * eb));
        AFP_CUDA(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        AFP_CUDA(c, cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
        AFP_CUDA(c, cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    AFP_CUDA(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    AFP_CUDA(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    return AFP_OK;
}

// Pipeline unit = a chunk of streams of one block.  Few, large units keep the copy engines at
// full PCIe rate and the launch/event overhead low; a single-block call is still cut into chunks
// so that its own H2D, kernel and D2H overlap.
static int pick_slot_streams(const afp_ctx* c, int frames, int n_blocks) {
    const int S = c->cfg.n_streams;
    int chunks = 1;
    const long long in_bytes = (long long)S * frames * 4;
    while (chunks * n_blocks < 8 && chunks < 8 && in_bytes / (chunks * 2) >= (1 << 20) && S / (chunks * 2) >= 64)
        chunks *= 2;
    return ((S + chunks - 1) / chunks + 1) & ~1;   // even: row pairs stay inside a chunk
}

int afp_process_host_blocks(afp_ctx* c, const float* in, void* out, int32_t frames, int32_t n_blocks) {
    if (!c) return AFP_ERR_INVALID;
    if (!in || !out) return fail(c, AFP_ERR_INVALID, "null buffer");
    if (n_blocks < 0) return fail(c, AFP_ERR_INVALID, "negative block count");
    int rc = check_ready(c, frames);
    if (rc) return rc;
    if (frames == 0 || n_blocks == 0) return AFP_OK;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    rc = ensure_tmp(c);
    if (rc) return rc;
    rc = ensure_host_staging(c);
    if (rc) return rc;
    const int S = c->cfg.n_streams, eb = afp_out_elem_bytes(c);
    const int out_n = afp_out_frames(c, frames);
    c->slot_streams = pick_slot_streams(c, frames, n_blocks);
    const size_t in_row = (size_t)frames * sizeof(float), out_row = (size_t)out_n * eb;
    const unsigned char* hin = reinterpret_cast<const unsigned char*>(in);
    unsigned char* hout = static_cast<unsigned char*>(out);
    if (n_blocks == 1 && c->slot_streams >= S) {
        // one unit (the sounddevice callback: one small block): nothing to overlap, so copy in, run and copy out on ONE
        // stream - no cross-stream event hops on the audio thread's critical path
        AFP_CUDA(c, cudaMemcpyAsync(c->d_in[0], hin, (size_t)S * in_row, cudaMemcpyHostToDevice, c->own_stream));
        rc = process_block_range(c, c->d_in[0], frames, c->d_out[0], (int64_t)out_n * (eb == 3 ? 3 : 1), frames, 0, S, c->own_stream);
        if (rc) return rc;
        AFP_CUDA(c, cudaMemcpyAsync(hout, c->d_out[0], (size_t)S * out_row, cudaMemcpyDeviceToHost, c->own_stream));
        finish_block(c, frames);
        AFP_CUDA(c, cudaStreamSynchronize(c->own_stream));
        return check_range(c);
    }
    long long unit = 0;
    for (int b = 0; b < n_blocks; ++b) {
        for (int s0 = 0; s0 < S; s0 += c->slot_streams, ++unit) {
            const int ns = S - s0 < c->slot_streams ? S - s0 : c->slot_streams;
            const int slot = (int)(unit % afp_ctx::kSlots);
            const bool reused = unit >= afp_ctx::kSlots;
            // H2D: the slot's previous kernel must have consumed its input
            if (reused) AFP_CUDA(c, cudaStreamWaitEvent(c->s_in, c->ev_k[slot], 0));
            AFP_CUDA(c, cudaMemcpyAsync(c->d_in[slot], hin + ((size_t)b * S + s0) * in_row, ns * in_row,
                                        cudaMemcpyHostToDevice, c->s_in));
            AFP_CUDA(c, cudaEventRecord(c->ev_in[slot], c->s_in));
            // compute: input landed, and the slot's previous output has been copied out
            AFP_CUDA(c, cudaStreamWaitEvent(c->own_stream, c->ev_in[slot], 0));
            if (reused) AFP_CUDA(c, cudaStreamWaitEvent(c->own_stream, c->ev_out[slot], 0));
            rc = process_block_range(c, c->d_in[slot], frames, c->d_out[slot], (int64_t)out_n * (eb == 3 ? 3 : 1),
                                     frames, s0, ns, c->own_stream);
            if (rc) return rc;
            AFP_CUDA(c, cudaEventRecord(c->ev_k[slot], c->own_stream));
            // D2H
            AFP_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_k[slot], 0));
            AFP_CUDA(c, cudaMemcpyAsync(hout + ((size_t)b * S + s0) * out_row, c->d_out[slot], ns * out_row,
                                        cudaMemcpyDeviceToHost, c->s_out));
            AFP_CUDA(c, cudaEventRecord(c->ev_out[slot], c->s_out));
        }
        finish_block(c, frames);
    }
    AFP_CUDA(c, cudaStreamSynchronize(c->s_out));
    AFP_CUDA(c, cudaStreamSynchronize(c->own_stream));
    return check_range(c);
}

int afp_process_host(afp_ctx* c, const float* in, void* out, int32_t frames) {
    return afp_process_host_blocks(c, in, out, frames, 1);
}

int afp_process_host_interleaved(afp_ctx* c, const float* in, void* out, int32_t frames, int32_t channels) {
    if (!c) return AFP_ERR_INVALID;
    if (!in || !out) return fail(c, AFP_ERR_INVALID, "null buffer");
    const int S = c->cfg.n_streams;
    if (channels < 1 || S % channels != 0) return fail(c, AFP_ERR_INVALID, "n_streams is not a multiple of channels");
    if (channels == 1) return afp_process_host_blocks(c, in, out, frames, 1);   // mono: the two layouts coincide
    int rc = check_ready(c, frames);
    if (rc) return rc;
    if (frames == 0) return AFP_OK;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    rc = ensure_tmp(c);
    if (rc) return rc;
    rc = ensure_host_staging(c);
    if (rc) return rc;
    const int eb = afp_out_elem_bytes(c), out_n = afp_out_frames(c, frames), sessions = S / channels;
    if (!c->d_il_in) {
        AFP_CUDA(c, cudaMalloc(&c->d_il_in, (size_t)S * c->cfg.max_frames * sizeof(float)));
        AFP_CUDA(c, cudaMalloc(&c->d_il_out, (size_t)S * afp_out_frames(c, c->cfg.max_frames) * eb));
    }
    // Sessions travel in chunks through three streams (H2D | de-interleave, chain, interleave | D2H), like the planar
    // entry point: chunk i+1 is copied in while chunk i is computed and chunk i-1 is copied out.  Every chunk has its own
    // region of the (whole-batch) device buffers, so nothing is reused inside a call and the events only order the streams.
    int chunk_streams = pick_slot_streams(c, frames, 1);
    const int unit = channels % 2 == 0 ? channels : 2 * channels;          // whole sessions, and row pairs stay inside a chunk
    chunk_streams = (chunk_streams + unit - 1) / unit * unit;
    auto grid_for = [&](long long total) { long long g = (total + 255) / 256; return (int)(g < 1 ? 1 : (g > 8LL * c->n_sm ? 8LL * c->n_sm : g)); };
    const size_t in_row = (size_t)frames * sizeof(float), out_row = (size_t)out_n * eb;   // bytes per stream
    const unsigned char* hin = reinterpret_cast<const unsigned char*>(in);
    unsigned char* hout = static_cast<unsigned char*>(out);
    const bool one = chunk_streams >= S;     // a single chunk (the callback's one small block): everything on one stream
    cudaStream_t st_in = one ? c->own_stream : c->s_in, st_out = one ? c->own_stream : c->s_out;
    int i = 0;
    for (int s0 = 0; s0 < S; s0 += chunk_streams, ++i) {
        const int ns = S - s0 < chunk_streams ? S - s0 : chunk_streams, se = ns / channels, k = i % afp_ctx::kSlots;
        float* il_in = c->d_il_in + (size_t)s0 * frames;
        unsigned char* il_out = c->d_il_out + (size_t)s0 * out_row;
        float* pl_in = c->d_in[0] + (size_t)s0 * frames;
        unsigned char* pl_out = c->d_out[0] + (size_t)s0 * out_row;
        AFP_CUDA(c, cudaMemcpyAsync(il_in, hin + (size_t)s0 * in_row, (size_t)ns * in_row, cudaMemcpyHostToDevice, st_in));
        if (!one) {
            AFP_CUDA(c, cudaEventRecord(c->ev_in[k], c->s_in));
            AFP_CUDA(c, cudaStreamWaitEvent(c->own_stream, c->ev_in[k], 0));
        }
        deinterleave_kernel<<<grid_for((long long)se * frames), 256, 0, c->own_stream>>>(il_in, pl_in, se, frames, channels, frames);
        AFP_CUDA(c, cudaGetLastError());
        rc = process_block_range(c, pl_in, frames, pl_out, (int64_t)out_n * (eb == 3 ? 3 : 1), frames, s0, ns, c->own_stream);
        if (rc) return rc;
        const long long pitch_bytes = (long long)out_n * eb;
        const int g = grid_for((long long)se * out_n);
        if (eb == 4) interleave_kernel<4><<<g, 256, 0, c->own_stream>>>(pl_out, il_out, se, out_n, channels, pitch_bytes);
        else if (eb == 2) interleave_kernel<2><<<g, 256, 0, c->own_stream>>>(pl_out, il_out, se, out_n, channels, pitch_bytes);
        else interleave_kernel<3><<<g, 256, 0, c->own_stream>>>(pl_out, il_out, se, out_n, channels, pitch_bytes);
        AFP_CUDA(c, cudaGetLastError());
        c->launches += 2;
        if (!one) {
            AFP_CUDA(c, cudaEventRecord(c->ev_k[k], c->own_stream));
            AFP_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
        }
        AFP_CUDA(c, cudaMemcpyAsync(hout + (size_t)s0 * out_row, il_out, (size_t)ns * out_row, cudaMemcpyDeviceToHost, st_out));
    }
    if (!one) AFP_CUDA(c, cudaStreamSynchronize(c->s_out));
    AFP_CUDA(c, cudaStreamSynchronize(c->own_stream));
    finish_block(c, frames);
    return check_range(c);
}

// the stages that actually run, in launch order
static std::vector<const Stage*> live_stages(const afp_ctx* c) {
    if (c->fused.active) return {&c->fused};
    return {&c->up, &c->eq, &c->mainf};
}

// every device buffer that carries state between blocks: (current, other, bytes)
struct StateBuf { float* cur; float* other; size_t bytes; };
static std::vector<StateBuf> state_buffers(const afp_ctx* c) {
    std::vector<StateBuf> v;
    for (const Stage* s : live_stages(c))
        if (s->active && s->hist[0])
            v.push_back({s->hist[s->cur], s->hist[s->cur ^ 1], (size_t)c->cfg.n_streams * s->hist_pitch * sizeof(float)});
    if (c->xhist[0])
        v.push_back({c->xhist[c->xhist_cur], c->xhist[c->xhist_cur ^ 1],
                     (size_t)c->cfg.n_streams * c->xhist_pitch * sizeof(float)});
    return v;
}

int afp_reset(afp_ctx* c) {
    if (!c) return AFP_ERR_INVALID;
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    AFP_CUDA(c, cudaDeviceSynchronize());
    for (const StateBuf& b : state_buffers(c)) {
        AFP_CUDA(c, cudaMemset(b.cur, 0, b.bytes));
        AFP_CUDA(c, cudaMemset(b.other, 0, b.bytes));
    }
    if (c->range_flag) AFP_CUDA(c, cudaMemset(c->range_flag, 0, sizeof(unsigned int)));
    AFP_CUDA(c, cudaDeviceSynchronize());   // the memsets run on the legacy stream, the kernels on non-blocking ones
    c->out_samples = 0;
    return AFP_OK;
}

int afp_state_size(const afp_ctx* c, size_t* bytes) {
    if (!c || !bytes) return AFP_ERR_INVALID;
    size_t n = 4 * sizeof(uint64_t);
    for (const StateBuf& b : state_buffers(c)) n += b.bytes;
    *bytes = n;
    return AFP_OK;
}

int afp_get_state(afp_ctx* c, void* buf, size_t bytes) {
    if (!c || !buf) return fail(c, AFP_ERR_INVALID, "null argument");
    size_t need = 0; afp_state_size(c, &need);
    if (bytes < need) return fail(c, AFP_ERR_INVALID, "state buffer too small");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    AFP_CUDA(c, cudaDeviceSynchronize());
    uint64_t* hdr = static_cast<uint64_t*>(buf);
    hdr[0] = kStateMagic; hdr[1] = need; hdr[2] = c->out_samples; hdr[3] = 0;
    unsigned char* p = static_cast<unsigned char*>(buf) + 4 * sizeof(uint64_t);
    for (const StateBuf& b : state_buffers(c)) {
        AFP_CUDA(c, cudaMemcpy(p, b.cur, b.bytes, cudaMemcpyDeviceToHost));
        p += b.bytes;
    }
    return AFP_OK;
}

int afp_set_state(afp_ctx* c, const void* buf, size_t bytes) {
    if (!c || !buf) return fail(c, AFP_ERR_INVALID, "null argument");
    size_t need = 0; afp_state_size(c, &need);
    const uint64_t* hdr = static_cast<const uint64_t*>(buf);
    if (bytes < need || hdr[0] != kStateMagic || hdr[1] != need)
        return fail(c, AFP_ERR_INVALID, "state blob does not match this context's filters");
    AFP_CUDA(c, cudaSetDevice(c->cfg.device));
    AFP_CUDA(c, cudaDeviceSynchronize());
    c->out_samples = hdr[2];
    const unsigned char* p = static_cast<const unsigned char*>(buf) + 4 * sizeof(uint64_t);
    for (const StateBuf& b : state_buffers(c)) {
        AFP_CUDA(c, cudaMemcpy(b.cur, p, b.bytes, cudaMemcpyHostToDevice));
        p += b.bytes;
    }
    AFP_CUDA(c, cudaDeviceSynchronize());   // (pageable-memory copies may still be in flight on the legacy stream)
    return AFP_OK;
}

int afp_get_stage_info(const afp_ctx* c, int32_t stage, afp_stage_info* info) {
    if (!c || !info) return AFP_ERR_INVALID;
    int k = 0;
    for (const Stage* s : live_stages(c)) {
        if (!s->active) continue;
        if (k++ != stage) continue;
        info->kind = s->kind; info->upsample = s->U; info->n_taps = s->n_taps;
        info->taps_per_phase_padded = s->kq_pad; info->tile_taps = s->last_T ? s->last_T : s->T;
        info->threads = s->geo.nt; info->threads_per_segment = s->geo.nt ? (1 << s->geo.tg_log2) : 0;
        info->smem_bytes = s->geo.smem; info->grid = s->geo.grid; info->history = s->kq_pad;
        info->kernel = s->last_kernel;
        return AFP_OK;
    }
    return AFP_ERR_INVALID;
}

int64_t afp_kernel_launches(const afp_ctx* c) { return c ? c->launches : 0; }

int afp_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return AFP_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
    return AFP_OK;
}

int afp_host_alloc_wc(void** ptr, size_t bytes) {
    if (!ptr) return AFP_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocWriteCombined);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc(write-combined)");
    return AFP_OK;
}

int afp_host_free(void* ptr) {
    if (!ptr) return AFP_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFreeHost");
    return AFP_OK;
}

int afp_measure_fma_peak(int32_t device, double* tflops, double* sm_mhz) {
    if (!tflops || !sm_mhz) return AFP_ERR_INVALID;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    const int grid = prop.multiProcessorCount * 2, threads = 256, iters = 4000;
    float *out = nullptr, *in = nullptr; long long* cyc = nullptr;
    if (cudaMalloc(&out, sizeof(float) * grid * threads) != cudaSuccess ||
        cudaMalloc(&in, sizeof(float) * 4096) != cudaSuccess ||
        cudaMalloc(&cyc, sizeof(long long) * grid) != cudaSuccess) {
        cudaFree(out); cudaFree(in); cudaFree(cyc);      // (cudaFree(nullptr) is a no-op)
        cudaGetLastError();
        return fail(nullptr, AFP_ERR_NOMEM, "cudaMalloc failed in afp_measure_fma_peak");
    }
    cudaMemset(in, 0, sizeof(float) * 4096);
    fma_peak_kernel<8><<<grid, threads>>>(out, in, 100, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<8><<<grid, threads>>>(out, in, iters, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    e = cudaGetLastError();
    std::vector<long long> hc(grid);
    cudaMemcpy(hc.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long v : hc) cmax = v > cmax ? v : cmax;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "fma_peak_kernel");
    const double fma = 2.0 * 4 * 8 * 8 * (double)iters * threads * grid;   // FMAs (2 per FFMA2 lane)
    *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    *sm_mhz = (double)cmax / (best * 1e-3) / 1e6;
    return AFP_OK;
}

}  // extern "C"
