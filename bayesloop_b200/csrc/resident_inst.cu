// resident_inst.cu -- generic resident / stream kernels: nvcc -c -DBLG_INST_BWD={0,1} resident_inst.cu
#include "kernels.h"
#include "resident.cuh"

namespace blg {

#if BLG_INST_BWD
PassKernel bwd_resident_entry(int nt, bool stream) {
    if (stream) return bwd_resident_kernel<1024, 1, true>;
    if (nt <= 256) return bwd_resident_kernel<256, 4, false>;
    if (nt <= 512) return bwd_resident_kernel<512, 2, false>;
    return bwd_resident_kernel<1024, 1, false>;
}
#else
PassKernel fwd_resident_entry(int nt, bool stream) {
    if (stream) return fwd_resident_kernel<1024, 1, true>;
    if (nt <= 256) return fwd_resident_kernel<256, 4, false>;
    if (nt <= 512) return fwd_resident_kernel<512, 2, false>;
    return fwd_resident_kernel<1024, 1, false>;
}
#endif

}  // namespace blg
