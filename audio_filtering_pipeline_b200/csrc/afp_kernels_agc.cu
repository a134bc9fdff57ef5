// agc_kernel instantiations (afp_agc.cuh): the fast max/min form and the checked form that mirrors
// the reference's divide-by-zero path, each with 1 / 2 / 4 / 8 scans of 128 samples in flight per warp.
#include "afp_agc.cuh"

namespace {
template <bool CHECKED>
afp::agc_kern_t pick(int scan_blocks) {
    switch (scan_blocks) {
        case 1: return (afp::agc_kern_t)afp::agc_kernel<CHECKED, 1>;
        case 2: return (afp::agc_kern_t)afp::agc_kernel<CHECKED, 2>;
        case 4: return (afp::agc_kern_t)afp::agc_kernel<CHECKED, 4>;
        case 8: return (afp::agc_kern_t)afp::agc_kernel<CHECKED, 8>;
        default: return nullptr;
    }
}
}  // namespace

afp::agc_kern_t afp_kernel_agc(bool checked, int scan_blocks) {
    return checked ? pick<true>(scan_blocks) : pick<false>(scan_blocks);
}
