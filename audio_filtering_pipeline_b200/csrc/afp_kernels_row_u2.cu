// fir_row_kernel instantiations for U = 2 (linear layout + bulk copies): 8-sample tiles with direct
// stores for short phases, 16-sample tiles with transposed stores above.
#include "afp_kernels_row.cuh"

afp::kern_t afp_kernel_row_u2(int T, int pad, int store, int epi, int spec_tgl) {
    if (T == 8 && pad == 0 && store == 0) return afp::pick_row_kernel<2, 8, 0, 0, 6>(epi, spec_tgl);
    if (T == 16 && pad == 0 && store == 2) return afp::pick_row_kernel<2, 16, 0, 2, 5>(epi, spec_tgl);
    return nullptr;
}
