// cluster2d_inst.cu -- cluster-resident 2-D kernels: nvcc -c -DBLG_INST_BWD={0,1} cluster2d_inst.cu
#include "kernels.h"
#include "cluster2d.cuh"

namespace blg {

#if BLG_INST_BWD
PassKernel bwd_cluster2d_entry(bool prof, int m0) {
    if (m0 == 13) return prof ? bwd_cluster2d_kernel<kC2Threads, true, 13> : bwd_cluster2d_kernel<kC2Threads, false, 13>;
    return prof ? bwd_cluster2d_kernel<kC2Threads, true, 16> : bwd_cluster2d_kernel<kC2Threads, false, 16>;
}
#else
PassKernel fwd_cluster2d_entry(bool prof, int m0) {
    if (m0 == 13) return prof ? fwd_cluster2d_kernel<kC2Threads, true, 13> : fwd_cluster2d_kernel<kC2Threads, false, 13>;
    return prof ? fwd_cluster2d_kernel<kC2Threads, true, 16> : fwd_cluster2d_kernel<kC2Threads, false, 16>;
}
bool cluster2d_supports(int n_ops, const int *kind, const int *axis) { return classify2d(n_ops, kind, axis).ok; }
void cluster2d_params(int *threads, int *m0, int *m1, int *cells, int *wpad) {
    *threads = kC2Threads;
    *m0 = kC2M0;
    *m1 = kC2M1;
    *cells = kC2Cells;
    *wpad = kC2WPad;
}
#endif

}  // namespace blg
