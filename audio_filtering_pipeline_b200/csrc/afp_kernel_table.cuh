// Kernel instantiation table shared by afp_abi.cu (lookup) and afp_kernels_u{1..4}.cu (definitions).
#pragma once
#include "afp_fir_kernel.cuh"

namespace afp {

typedef void (*kern_t)(const FirParams);

// Epilogues compiled with format / dither fixed (decim == 1); everything else runs EPI_GENERIC.
constexpr int kEpiF32 = epi_code(FMT_F32, DITHER_NONE);
constexpr int kEpiI24T = epi_code(FMT_I24_IN_I32, DITHER_TPDF);
constexpr int kEpiI16T = epi_code(FMT_I16, DITHER_TPDF);
constexpr int kEpiP24T = epi_code(FMT_I24_PACKED, DITHER_TPDF);
constexpr int kAllEpi[] = {EPI_GENERIC, kEpiF32, kEpiI24T, kEpiI16T, kEpiP24T};

// NT = 128 (only used when a 256-thread CTA does not fit in shared memory) and the multi-table
// (per-band EQ) kernels exist with the generic epilogue only.
template <int U, int T, bool ALLOW_MULTI>
kern_t pick_kernel(int nt, bool multi, int epi) {
    if (multi) {
        if (!ALLOW_MULTI) return nullptr;
        return nt == 256 ? (kern_t)fir_stream_kernel<U, T, 256, ALLOW_MULTI, EPI_GENERIC>
                         : (kern_t)fir_stream_kernel<U, T, 128, ALLOW_MULTI, EPI_GENERIC>;
    }
    if (nt == 128) return (kern_t)fir_stream_kernel<U, T, 128, false, EPI_GENERIC>;
    switch (epi) {
        case kEpiF32: return (kern_t)fir_stream_kernel<U, T, 256, false, kEpiF32>;
        case kEpiI24T: return (kern_t)fir_stream_kernel<U, T, 256, false, kEpiI24T>;
        case kEpiI16T: return (kern_t)fir_stream_kernel<U, T, 256, false, kEpiI16T>;
        case kEpiP24T: return (kern_t)fir_stream_kernel<U, T, 256, false, kEpiP24T>;
        default: return (kern_t)fir_stream_kernel<U, T, 256, false, EPI_GENERIC>;
    }
}

// per-stream-group filters: taps from a per-item table in global memory; generic epilogue, 256-thread CTAs
template <int U, int T>
kern_t pick_grouped_kernel() { return (kern_t)fir_stream_kernel<U, T, 256, false, EPI_GENERIC, true>; }

}  // namespace afp

namespace afp { struct TcParams; typedef void (*tc_kern_t)(const TcParams); }
afp::tc_kern_t afp_kernel_tc(int row, int nts, int epi);   // tensor-core path: row length 32 / 64, nts = frames / (8 row)
// single-row kernel (afp_fir_row.cuh): 256-thread CTAs, lookup by tile, layout (pad), store path, epilogue
// spec_tgl: log2(group size) when the launch is the streaming form (selects the instantiation with that
// size compiled in, if there is one), else 0
afp::kern_t afp_kernel_row_u1(int T, int pad, int store, int epi, int spec_tgl);
afp::kern_t afp_kernel_row_u2(int T, int pad, int store, int epi, int spec_tgl);
afp::kern_t afp_kernel_row_u3(int T, int pad, int store, int epi, int spec_tgl);
afp::kern_t afp_kernel_row_u4(int T, int pad, int store, int epi, int spec_tgl);
afp::kern_t afp_kernel_u1(int T, int nt, bool multi, int epi);
afp::kern_t afp_kernel_u2(int T, int nt, bool multi, int epi);
afp::kern_t afp_kernel_u3(int T, int nt, bool multi, int epi);
afp::kern_t afp_kernel_u4(int T, int nt, bool multi, int epi);
afp::kern_t afp_kernel_grouped(int U, int T);   // afp_kernels_grouped.cu
