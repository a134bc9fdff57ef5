// fir_stream_kernel instantiations with per-stream-group tap tables (afp_assign_groups); see afp_kernel_table.cuh.
#include "afp_kernel_table.cuh"

afp::kern_t afp_kernel_grouped(int U, int T) {
    switch (U * 100 + T) {
        case 116: return afp::pick_grouped_kernel<1, 16>();
        case 208: return afp::pick_grouped_kernel<2, 8>();
        case 304: return afp::pick_grouped_kernel<3, 4>();
        case 408: return afp::pick_grouped_kernel<4, 8>();
        case 404: return afp::pick_grouped_kernel<4, 4>();
    }
    return nullptr;
}
