// fir_row_kernel instantiations for U = 3 (phase-major tap-pair form, linear layout + bulk copies,
// transposed stores): 8-sample tiles.
#include "afp_kernels_row.cuh"

afp::kern_t afp_kernel_row_u3(int T, int pad, int store, int epi, int spec_tgl) {
    if (T == 8 && pad == 0 && store == 2) return afp::pick_row_kernel<3, 8, 0, 2, 5>(epi, spec_tgl);
    return nullptr;
}
