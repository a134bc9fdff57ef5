// fir_row_kernel instantiations for U = 1: 16-sample tiles, padded layout + cp.async, direct stores.
#include "afp_kernels_row.cuh"

afp::kern_t afp_kernel_row_u1(int T, int pad, int store, int epi, int spec_tgl) {
    if (T == 16 && pad == 1 && store == 0) return afp::pick_row_kernel<1, 16, 1, 0, 5>(epi, spec_tgl);
    return nullptr;
}
