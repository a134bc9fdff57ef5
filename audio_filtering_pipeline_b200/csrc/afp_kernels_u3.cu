// fir_stream_kernel instantiations for up-sampling factor U = 3 (tiles [4]); see afp_kernel_table.cuh.
#include "afp_kernel_table.cuh"

afp::kern_t afp_kernel_u3(int T, int nt, bool multi, int epi) {
    switch (T) {
        case 4: return afp::pick_kernel<3, 4, false>(nt, multi, epi);
    }
    return nullptr;
}
