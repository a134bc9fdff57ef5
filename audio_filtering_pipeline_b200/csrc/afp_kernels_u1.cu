// fir_stream_kernel instantiations for up-sampling factor U = 1 (tiles [16]); see afp_kernel_table.cuh.
#include "afp_kernel_table.cuh"

afp::kern_t afp_kernel_u1(int T, int nt, bool multi, int epi) {
    switch (T) {
        case 16: return afp::pick_kernel<1, 16, true>(nt, multi, epi);
    }
    return nullptr;
}
