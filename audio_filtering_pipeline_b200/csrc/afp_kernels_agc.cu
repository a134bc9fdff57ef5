// agc_kernel instantiations (afp_agc.cuh): variant (fast / exact gain / checked recurrence) x one or two
// 512-sample scans in flight per warp x one or two warps per stream in the gain phase.
#include "afp_agc.cuh"

namespace {
template <int VARIANT>
afp::agc_kern_t pick(int scan_blocks, int team) {
    if (team == 2)
        return scan_blocks == 1 ? (afp::agc_kern_t)afp::agc_kernel<VARIANT, 1, 2> : (afp::agc_kern_t)afp::agc_kernel<VARIANT, 2, 2>;
    return scan_blocks == 1 ? (afp::agc_kern_t)afp::agc_kernel<VARIANT, 1, 1> : (afp::agc_kern_t)afp::agc_kernel<VARIANT, 2, 1>;
}
}  // namespace

afp::agc_kern_t afp_kernel_agc(int variant, int scan_blocks, int team) {
    switch (variant) {
        case afp::AGC_FAST: return pick<afp::AGC_FAST>(scan_blocks, team);
        case afp::AGC_EXACT: return pick<afp::AGC_EXACT>(scan_blocks, team);
        default: return pick<afp::AGC_CHECKED>(scan_blocks, team);
    }
}
