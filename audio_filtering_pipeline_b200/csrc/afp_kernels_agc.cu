// agc_kernel instantiations (afp_agc.cuh): variant (fast / exact gain / checked recurrence) x one or two
// 512-sample scans in flight per warp.
#include "afp_agc.cuh"

namespace {
template <int VARIANT>
afp::agc_kern_t pick(int scan_blocks) {
    return scan_blocks == 1 ? (afp::agc_kern_t)afp::agc_kernel<VARIANT, 1> : (afp::agc_kern_t)afp::agc_kernel<VARIANT, 2>;
}
}  // namespace

afp::agc_kern_t afp_kernel_agc(int variant, int scan_blocks) {
    switch (variant) {
        case afp::AGC_FAST: return pick<afp::AGC_FAST>(scan_blocks);
        case afp::AGC_EXACT: return pick<afp::AGC_EXACT>(scan_blocks);
        default: return pick<afp::AGC_CHECKED>(scan_blocks);
    }
}
