// fir_tc_kernel (tcgen05 / TMEM Toeplitz FIR) instantiations; see afp_fir_tc.cuh.
#include "afp_kernel_table.cuh"
#include "afp_fir_tc.cuh"

namespace {
template <int NTS, int R>
afp::tc_kern_t pick_tc(int epi) {
    using namespace afp;
    switch (epi) {
        case kEpiF32: return (tc_kern_t)fir_tc_kernel<kEpiF32, NTS, R>;
        case kEpiI24T: return (tc_kern_t)fir_tc_kernel<kEpiI24T, NTS, R>;
        case kEpiI16T: return (tc_kern_t)fir_tc_kernel<kEpiI16T, NTS, R>;
        case kEpiP24T: return (tc_kern_t)fir_tc_kernel<kEpiP24T, NTS, R>;
        default: return (tc_kern_t)fir_tc_kernel<EPI_GENERIC, NTS, R>;
    }
}
}  // namespace

// row = 64: frames / 512 in {1, 2};  row = 32: frames / 256 in {1, 2, 4}
afp::tc_kern_t afp_kernel_tc(int row, int nts, int epi) {
    if (row == 64) {
        if (nts == 1) return pick_tc<1, 64>(epi);
        if (nts == 2) return pick_tc<2, 64>(epi);
    } else if (row == 32) {
        if (nts == 1) return pick_tc<1, 32>(epi);
        if (nts == 2) return pick_tc<2, 32>(epi);
        if (nts == 4) return pick_tc<4, 32>(epi);
    }
    return nullptr;
}
