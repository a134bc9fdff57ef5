// agc_kernel instantiations (afp_agc.cuh): fast gain / exact gain / checked recurrence.
#include "afp_agc.cuh"

afp::agc_kern_t afp_kernel_agc(int variant) {
    switch (variant) {
        case afp::AGC_FAST: return (afp::agc_kern_t)afp::agc_kernel<afp::AGC_FAST>;
        case afp::AGC_PROBE_F32: return (afp::agc_kern_t)afp::agc_kernel<afp::AGC_PROBE_F32>;
        case afp::AGC_EXACT: return (afp::agc_kern_t)afp::agc_kernel<afp::AGC_EXACT>;
        default: return (afp::agc_kern_t)afp::agc_kernel<afp::AGC_CHECKED>;
    }
}
