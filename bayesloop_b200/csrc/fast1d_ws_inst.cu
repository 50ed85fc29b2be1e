// fast1d_ws_inst.cu -- one translation unit per (direction, M, threads) of the warp-specialised 1-D kernels:
//   nvcc -c -DBLG_INST_M=11 -DBLG_INST_NT=128 -DBLG_INST_BWD=0 fast1d_ws_inst.cu
#include "fast1d_ws.cuh"
#include "kernels.h"

#if !defined(BLG_INST_M) || !defined(BLG_INST_NT)
#error "compile with -DBLG_INST_M=.. -DBLG_INST_NT=.. -DBLG_INST_BWD={0,1}"
#endif

namespace blg {

#define BLG_CAT4_(a, b, c, d) a##b##c##d
#define BLG_CAT4(a, b, c, d) BLG_CAT4_(a, b, c, d)

#if BLG_INST_BWD
PassKernel BLG_CAT4(bwd_fast1d_ws_m, BLG_INST_M, _nt, BLG_INST_NT)() { return bwd_fast1d_ws_kernel<BLG_INST_M, BLG_INST_NT>; }
#else
PassKernel BLG_CAT4(fwd_fast1d_ws_m, BLG_INST_M, _nt, BLG_INST_NT)() { return fwd_fast1d_ws_kernel<BLG_INST_M, BLG_INST_NT>; }
#endif

}  // namespace blg
