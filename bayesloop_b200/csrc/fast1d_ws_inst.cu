// fast1d_ws_inst.cu -- one translation unit per (direction, M, threads) of the warp-specialised 1-D kernels:
//   nvcc -c -DBLG_INST_M=9 -DBLG_INST_ML=5 -DBLG_INST_NT=160 -DBLG_INST_BWD=0 fast1d_ws_inst.cu
// (M cells per thread in the compute warps, ML in the last one)
#include "fast1d_ws.cuh"
#include "kernels.h"

#if !defined(BLG_INST_M) || !defined(BLG_INST_ML) || !defined(BLG_INST_NT)
#error "compile with -DBLG_INST_M=.. -DBLG_INST_ML=.. -DBLG_INST_NT=.. -DBLG_INST_BWD={0,1}"
#endif

namespace blg {

#define BLG_CAT6_(a, b, c, d, e, f) a##b##c##d##e##f
#define BLG_CAT6(a, b, c, d, e, f) BLG_CAT6_(a, b, c, d, e, f)

#if BLG_INST_BWD
PassKernel BLG_CAT6(bwd_fast1d_ws_m, BLG_INST_M, _l, BLG_INST_ML, _nt, BLG_INST_NT)() {
    return bwd_fast1d_ws_kernel<BLG_INST_M, BLG_INST_ML, BLG_INST_NT>;
}
#else
PassKernel BLG_CAT6(fwd_fast1d_ws_m, BLG_INST_M, _l, BLG_INST_ML, _nt, BLG_INST_NT)() {
    return fwd_fast1d_ws_kernel<BLG_INST_M, BLG_INST_ML, BLG_INST_NT>;
}
#endif

}  // namespace blg
