// Instantiation helper for the single-row kernel (afp_fir_row.cuh): one shape = (U, T, PAD, STORE), five
// epilogues each.  Shapes are chosen by afp_abi.cu::pick_row_shape from measurements.
#pragma once
#include "afp_kernel_table.cuh"
#include "afp_fir_row.cuh"

namespace afp {

template <int U, int T, int PAD, int STORE>
kern_t pick_row_kernel(int epi) {
    switch (epi) {
        case kEpiF32: return (kern_t)fir_row_kernel<U, T, 256, kEpiF32, STORE, PAD>;
        case kEpiI24T: return (kern_t)fir_row_kernel<U, T, 256, kEpiI24T, STORE, PAD>;
        case kEpiI16T: return (kern_t)fir_row_kernel<U, T, 256, kEpiI16T, STORE, PAD>;
        case kEpiP24T: return (kern_t)fir_row_kernel<U, T, 256, kEpiP24T, STORE, PAD>;
        default: return (kern_t)fir_row_kernel<U, T, 256, EPI_GENERIC, STORE, PAD>;
    }
}

}  // namespace afp
