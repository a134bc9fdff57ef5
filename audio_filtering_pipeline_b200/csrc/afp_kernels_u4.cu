// fir_stream_kernel instantiations for up-sampling factor U = 4 (tiles [8, 4]); see afp_kernel_table.cuh.
#include "afp_kernel_table.cuh"

afp::kern_t afp_kernel_u4(int T, int nt, bool multi, int epi) {
    switch (T) {
        case 8: return afp::pick_kernel<4, 8, false>(nt, multi, epi);
        case 4: return afp::pick_kernel<4, 4, false>(nt, multi, epi);
    }
    return nullptr;
}
