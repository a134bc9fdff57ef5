// Instantiation helper for the single-row kernel (afp_fir_row.cuh): one shape = (U, T, PAD, STORE), five
// epilogues each.  Shapes are chosen by afp_abi.cu::pick_row_shape from measurements.
#pragma once
#include "afp_kernel_table.cuh"
#include "afp_fir_row.cuh"

namespace afp {

// TGL = 0: general form (any group size, any launch geometry).  TGL > 0: the streaming form with that
// group size compiled in (-6.5 % executed instructions, +1-5 % throughput: profiles/r01_experiments/row_probe7.txt).
template <int U, int T, int PAD, int STORE, int TGL>
kern_t pick_row_kernel_1(int epi) {
    constexpr bool ST = TGL != 0;
    switch (epi) {
        case kEpiF32: return (kern_t)fir_row_kernel<U, T, 256, kEpiF32, STORE, PAD, TGL, ST>;
        case kEpiI24T: return (kern_t)fir_row_kernel<U, T, 256, kEpiI24T, STORE, PAD, TGL, ST>;
        case kEpiI16T: return (kern_t)fir_row_kernel<U, T, 256, kEpiI16T, STORE, PAD, TGL, ST>;
        case kEpiP24T: return (kern_t)fir_row_kernel<U, T, 256, kEpiP24T, STORE, PAD, TGL, ST>;
        default: return (kern_t)fir_row_kernel<U, T, 256, EPI_GENERIC, STORE, PAD, TGL, ST>;
    }
}
// spec_tgl: the caller's group size when the launch is the streaming form, else 0
template <int U, int T, int PAD, int STORE, int TGL>
kern_t pick_row_kernel(int epi, int spec_tgl) {
    return spec_tgl == TGL ? pick_row_kernel_1<U, T, PAD, STORE, TGL>(epi) : pick_row_kernel_1<U, T, PAD, STORE, 0>(epi);
}

}  // namespace afp
