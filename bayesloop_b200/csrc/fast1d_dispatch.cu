// fast1d_dispatch.cu -- selects the (M, threads) instantiation of the fast 1-D kernels (one object file per M and
// direction, see fast1d_inst.cu).
#include "kernels.h"

namespace blg {

#define BLG_DECL(M)                      \
    PassKernel fwd_fast1d_entry_m##M(int); \
    PassKernel bwd_fast1d_entry_m##M(int);
BLG_DECL(5)
BLG_DECL(7)
BLG_DECL(9)
#undef BLG_DECL

PassKernel fwd_fast1d_entry(int M, int nt) {
    switch (M) {
        case 5: return fwd_fast1d_entry_m5(nt);
        case 7: return fwd_fast1d_entry_m7(nt);
        case 9: return fwd_fast1d_entry_m9(nt);
        default: return nullptr;
    }
}

PassKernel bwd_fast1d_entry(int M, int nt) {
    switch (M) {
        case 5: return bwd_fast1d_entry_m5(nt);
        case 7: return bwd_fast1d_entry_m7(nt);
        case 9: return bwd_fast1d_entry_m9(nt);
        default: return nullptr;
    }
}

}  // namespace blg
