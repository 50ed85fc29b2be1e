// fast1d_il_inst.cu -- one translation unit per (direction, tiles per compute warp) of the interleaved 1-D kernels:
//   nvcc -c -DBLG_INST_TPW=1 -DBLG_INST_BWD=0 fast1d_il_inst.cu
#include "fast1d_il.cuh"
#include "kernels.h"

#if !defined(BLG_INST_TPW)
#error "compile with -DBLG_INST_TPW=.. -DBLG_INST_BWD={0,1}"
#endif

namespace blg {

#define BLG_CAT2_(a, b) a##b
#define BLG_CAT2(a, b) BLG_CAT2_(a, b)

#if BLG_INST_BWD
PassKernel BLG_CAT2(bwd_fast1d_il_t, BLG_INST_TPW)() { return bwd_fast1d_il_kernel<BLG_INST_TPW>; }
#else
PassKernel BLG_CAT2(fwd_fast1d_il_t, BLG_INST_TPW)() { return fwd_fast1d_il_kernel<BLG_INST_TPW>; }
#endif

}  // namespace blg
