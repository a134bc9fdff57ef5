// fir_tc_kernel (tcgen05 / TMEM Toeplitz FIR) instantiations; see afp_fir_tc.cuh.
#include "afp_kernel_table.cuh"
#include "afp_fir_tc.cuh"

namespace {
template <int NTS>
afp::tc_kern_t pick_tc(int epi) {
    using namespace afp;
    switch (epi) {
        case kEpiF32: return (tc_kern_t)fir_tc_kernel<kEpiF32, NTS>;
        case kEpiI24T: return (tc_kern_t)fir_tc_kernel<kEpiI24T, NTS>;
        case kEpiI16T: return (tc_kern_t)fir_tc_kernel<kEpiI16T, NTS>;
        case kEpiP24T: return (tc_kern_t)fir_tc_kernel<kEpiP24T, NTS>;
        default: return (tc_kern_t)fir_tc_kernel<EPI_GENERIC, NTS>;
    }
}
}  // namespace

afp::tc_kern_t afp_kernel_tc(int nts, int epi) {
    switch (nts) {
        case 1: return pick_tc<1>(epi);
        case 2: return pick_tc<2>(epi);
        case 4: return pick_tc<4>(epi);
    }
    return nullptr;
}
