// stream2d_inst.cu -- fused 2-D stream kernels: nvcc -c -DBLG_INST_BWD={0,1} stream2d_inst.cu
#include "kernels.h"
#include "stream2d.cuh"

namespace blg {

#if BLG_INST_BWD
PassKernel bwd_stream2d_entry() { return bwd_stream2d_kernel<512>; }
#else
PassKernel fwd_stream2d_entry() { return fwd_stream2d_kernel<512>; }
int stream2d_chunk() { return kM2d; }
bool stream2d_supports(int n_ops, const int *kind, const int *axis) { return classify2d(n_ops, kind, axis).ok; }
#endif

}  // namespace blg
