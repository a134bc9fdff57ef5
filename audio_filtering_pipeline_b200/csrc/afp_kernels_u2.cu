// fir_stream_kernel instantiations for up-sampling factor U = 2 (tiles [8]); see afp_kernel_table.cuh.
#include "afp_kernel_table.cuh"

afp::kern_t afp_kernel_u2(int T, int nt, bool multi, int epi) {
    switch (T) {
        case 8: return afp::pick_kernel<2, 8, false>(nt, multi, epi);
    }
    return nullptr;
}
