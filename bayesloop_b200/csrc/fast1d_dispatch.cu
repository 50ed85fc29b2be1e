// fast1d_dispatch.cu -- selects the (M, threads) instantiation of the fast 1-D kernels (one object file per M and
// direction, see fast1d_inst.cu).
#include "kernels.h"

namespace blg {

#define BLG_DECL(M)                      \
    PassKernel fwd_fast1d_entry_m##M(int); \
    PassKernel bwd_fast1d_entry_m##M(int);
BLG_DECL(9)
#undef BLG_DECL

PassKernel fwd_fast1d_entry(int M, int nt) {
    switch (M) {
        case 9: return fwd_fast1d_entry_m9(nt);
        default: return nullptr;
    }
}

PassKernel bwd_fast1d_entry(int M, int nt) {
    switch (M) {
        case 9: return bwd_fast1d_entry_m9(nt);
        default: return nullptr;
    }
}

// warp-specialised kernels (fast1d_ws_inst.cu): 4 compute warps + 1 service warp (160 threads) or 8 + 1 (288), so
// that every SM sub-partition carries the same number of compute warps of every chain
#define BLG_WS_ALL(X) X(3, 160) X(5, 160) X(7, 160) X(9, 160) X(11, 160) X(7, 288) X(9, 288) X(11, 288)
#define BLG_WS(M, NT)                          \
    PassKernel fwd_fast1d_ws_m##M##_nt##NT(); \
    PassKernel bwd_fast1d_ws_m##M##_nt##NT();
BLG_WS_ALL(BLG_WS)
#undef BLG_WS

PassKernel fwd_fast1d_ws_entry(int M, int nt) {
#define BLG_WS(MM, NT) \
    if (nt == NT && M == MM) return fwd_fast1d_ws_m##MM##_nt##NT();
    BLG_WS_ALL(BLG_WS)
#undef BLG_WS
    return nullptr;
}

PassKernel bwd_fast1d_ws_entry(int M, int nt) {
#define BLG_WS(MM, NT) \
    if (nt == NT && M == MM) return bwd_fast1d_ws_m##MM##_nt##NT();
    BLG_WS_ALL(BLG_WS)
#undef BLG_WS
    return nullptr;
}

}  // namespace blg
