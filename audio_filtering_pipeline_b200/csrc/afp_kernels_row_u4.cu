// fir_row_kernel instantiations for U = 4 (linear layout + bulk copies, transposed stores): 4-sample tiles
// for short phases, 8-sample tiles above.
#include "afp_kernels_row.cuh"

afp::kern_t afp_kernel_row_u4(int T, int pad, int store, int epi, int spec_tgl) {
    if (T == 4 && pad == 0 && store == 2) return afp::pick_row_kernel<4, 4, 0, 2, 7>(epi, spec_tgl);
    if (T == 8 && pad == 0 && store == 2) return afp::pick_row_kernel<4, 8, 0, 2, 5>(epi, spec_tgl);
    return nullptr;
}
