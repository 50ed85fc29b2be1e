// fast1d_inst.cu -- one translation unit per (direction, M) of the fast 1-D kernels:
//   nvcc -c -DBLG_INST_M=9 -DBLG_INST_BWD=0 fast1d_inst.cu -o fast1d_fwd9.o
#include "fast1d.cuh"
#include "kernels.h"

#ifndef BLG_INST_M
#error "compile with -DBLG_INST_M={5,7,9} -DBLG_INST_BWD={0,1}"
#endif

namespace blg {

#define BLG_CAT2(a, b) a##b
#define BLG_CAT(a, b) BLG_CAT2(a, b)

#if BLG_INST_BWD
#define BLG_KERNEL bwd_fast1d_kernel
#define BLG_ENTRY BLG_CAT(bwd_fast1d_entry_m, BLG_INST_M)
#else
#define BLG_KERNEL fwd_fast1d_kernel
#define BLG_ENTRY BLG_CAT(fwd_fast1d_entry_m, BLG_INST_M)
#endif

PassKernel BLG_ENTRY(int nt) {
    if (nt <= 128) return BLG_KERNEL<BLG_INST_M, 128, 4>;
    if (nt <= 160) return BLG_KERNEL<BLG_INST_M, 160, 4>;
    if (nt <= 256) return BLG_KERNEL<BLG_INST_M, 256, 4>;
    return BLG_KERNEL<BLG_INST_M, 1024, 1>;
}

}  // namespace blg
