// fast1d_mma_inst.cu -- one translation unit per (direction, tiles per compute warp, threads) of the DMMA 1-D kernels:
//   nvcc -c -DBLG_INST_TPW=4 -DBLG_INST_NT=160 -DBLG_INST_BWD=0 fast1d_mma_inst.cu
#include "fast1d_mma.cuh"
#include "kernels.h"

#if !defined(BLG_INST_TPW) || !defined(BLG_INST_NT)
#error "compile with -DBLG_INST_TPW=.. -DBLG_INST_NT=.. -DBLG_INST_BWD={0,1}"
#endif

namespace blg {

#define BLG_CAT4_(a, b, c, d) a##b##c##d
#define BLG_CAT4(a, b, c, d) BLG_CAT4_(a, b, c, d)

#if BLG_INST_BWD
PassKernel BLG_CAT4(bwd_fast1d_mma_t, BLG_INST_TPW, _nt, BLG_INST_NT)() { return bwd_fast1d_mma_kernel<BLG_INST_TPW, BLG_INST_NT>; }
#else
// prof: per-warp phase counters and the per-step event trace into PassArgs::trace (plan option trace)
PassKernel BLG_CAT4(fwd_fast1d_mma_t, BLG_INST_TPW, _nt, BLG_INST_NT)(bool prof) {
    return prof ? fwd_fast1d_mma_kernel<BLG_INST_TPW, BLG_INST_NT, true> : fwd_fast1d_mma_kernel<BLG_INST_TPW, BLG_INST_NT, false>;
}
#endif

}  // namespace blg
