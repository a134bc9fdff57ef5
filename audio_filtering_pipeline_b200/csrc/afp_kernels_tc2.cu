// fir_tc2_kernel (split-FP16 tcgen05 FIR with resident Hankel tap cores) instantiations; see afp_fir_tc2.cuh.
#include "afp_kernel_table.cuh"
#include "afp_fir_tc2.cuh"

afp::tc2_kern_t afp_kernel_tc2(int epi, int up) {
    using namespace afp;
    // polyphase interpolators (U = 2 / 3 / 4): plain float output specialised, everything else through the generic epilogue
    if (up == 2) return epi == kEpiF32 ? (tc2_kern_t)fir_tc2_kernel<kEpiF32, false, 2> : (tc2_kern_t)fir_tc2_kernel<EPI_GENERIC, false, 2>;
    if (up == 3) return epi == kEpiF32 ? (tc2_kern_t)fir_tc2_kernel<kEpiF32, false, 3> : (tc2_kern_t)fir_tc2_kernel<EPI_GENERIC, false, 3>;
    if (up == 4) return epi == kEpiF32 ? (tc2_kern_t)fir_tc2_kernel<kEpiF32, false, 4> : (tc2_kern_t)fir_tc2_kernel<EPI_GENERIC, false, 4>;
    if (up != 1) return nullptr;
    switch (epi) {
        case epi_code(FMT_F32, DITHER_RPDF): return (tc2_kern_t)fir_tc2_kernel<epi_code(FMT_F32, DITHER_RPDF)>;   // C1: the reference's float dither
        case kEpiF32: return (tc2_kern_t)fir_tc2_kernel<kEpiF32>;
        case kEpiI24T: return (tc2_kern_t)fir_tc2_kernel<kEpiI24T>;
        case kEpiI16T: return (tc2_kern_t)fir_tc2_kernel<kEpiI16T>;
        case kEpiP24T: return (tc2_kern_t)fir_tc2_kernel<kEpiP24T>;
        default: return (tc2_kern_t)fir_tc2_kernel<EPI_GENERIC>;
    }
}

// per-stream filters (afp_assign_groups): one stream per work item, generic epilogue
afp::tc2_kern_t afp_kernel_tc2_grouped() { return (afp::tc2_kern_t)afp::fir_tc2_kernel<afp::EPI_GENERIC, true>; }
