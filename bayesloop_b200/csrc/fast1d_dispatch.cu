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

// warp-specialised kernels (fast1d_ws_inst.cu)
#define BLG_WS(M, NT)                          \
    PassKernel fwd_fast1d_ws_m##M##_nt##NT(); \
    PassKernel bwd_fast1d_ws_m##M##_nt##NT();
BLG_WS(3, 128)
BLG_WS(7, 128)
BLG_WS(11, 128)
BLG_WS(11, 256)
BLG_WS(9, 160)
BLG_WS(7, 192)
#undef BLG_WS

PassKernel fwd_fast1d_ws_entry(int M, int nt) {
    if (nt == 128 && M == 3) return fwd_fast1d_ws_m3_nt128();
    if (nt == 128 && M == 7) return fwd_fast1d_ws_m7_nt128();
    if (nt == 128 && M == 11) return fwd_fast1d_ws_m11_nt128();
    if (nt == 256 && M == 11) return fwd_fast1d_ws_m11_nt256();
    if (nt == 160 && M == 9) return fwd_fast1d_ws_m9_nt160();
    if (nt == 192 && M == 7) return fwd_fast1d_ws_m7_nt192();
    return nullptr;
}

PassKernel bwd_fast1d_ws_entry(int M, int nt) {
    if (nt == 128 && M == 3) return bwd_fast1d_ws_m3_nt128();
    if (nt == 128 && M == 7) return bwd_fast1d_ws_m7_nt128();
    if (nt == 128 && M == 11) return bwd_fast1d_ws_m11_nt128();
    if (nt == 256 && M == 11) return bwd_fast1d_ws_m11_nt256();
    if (nt == 160 && M == 9) return bwd_fast1d_ws_m9_nt160();
    if (nt == 192 && M == 7) return bwd_fast1d_ws_m7_nt192();
    return nullptr;
}

}  // namespace blg
