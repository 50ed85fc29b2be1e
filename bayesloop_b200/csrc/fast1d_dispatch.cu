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
// that every SM sub-partition carries the same number of compute warps of every chain; M cells per thread, ML in the
// last compute warp (uneven split: fewer FP64 lanes spent on cells beyond the grid).  Keep in step with
// __graft_entry__.WS_UNITS.
#define BLG_WS_ALL(X)                                                                                     \
    X(3, 3, 160) X(5, 3, 160) X(5, 5, 160) X(7, 3, 160) X(7, 7, 160) X(9, 5, 160) X(9, 9, 160) X(11, 7, 160) \
    X(11, 11, 160) X(7, 7, 288) X(9, 9, 288) X(11, 11, 288)
#define BLG_WS(M, ML, NT)                              \
    PassKernel fwd_fast1d_ws_m##M##_l##ML##_nt##NT(); \
    PassKernel bwd_fast1d_ws_m##M##_l##ML##_nt##NT();
BLG_WS_ALL(BLG_WS)
#undef BLG_WS

PassKernel fwd_fast1d_ws_entry(int M, int ML, int nt) {
#define BLG_WS(MM, LL, NT) \
    if (nt == NT && M == MM && ML == LL) return fwd_fast1d_ws_m##MM##_l##LL##_nt##NT();
    BLG_WS_ALL(BLG_WS)
#undef BLG_WS
    return nullptr;
}

PassKernel bwd_fast1d_ws_entry(int M, int ML, int nt) {
#define BLG_WS(MM, LL, NT) \
    if (nt == NT && M == MM && ML == LL) return bwd_fast1d_ws_m##MM##_l##LL##_nt##NT();
    BLG_WS_ALL(BLG_WS)
#undef BLG_WS
    return nullptr;
}

// the geometries above, for the layout search in api.cu: {M, ML, threads}, terminated by M = 0
const int *fast1d_ws_geometries() {
    static const int table[] = {
#define BLG_WS(MM, LL, NT) MM, LL, NT,
        BLG_WS_ALL(BLG_WS)
#undef BLG_WS
        0, 0, 0};
    return table;
}

// DMMA kernels (fast1d_mma_inst.cu): tiles of 64 cells per compute warp x threads.  Keep in step with
// __graft_entry__.MMA_UNITS.
#define BLG_MMA_ALL(X) \
    X(1, 160) X(2, 160) X(3, 160) X(4, 160) X(5, 160) X(6, 160) X(1, 288) X(2, 288) X(4, 288) X(5, 288) X(6, 288)
#define BLG_MMA(TPW, NT)                           \
    PassKernel fwd_fast1d_mma_t##TPW##_nt##NT(bool); \
    PassKernel bwd_fast1d_mma_t##TPW##_nt##NT();
BLG_MMA_ALL(BLG_MMA)
#undef BLG_MMA

PassKernel fwd_fast1d_mma_entry(int tpw, int nt, bool prof) {
#define BLG_MMA(TPW, NT) \
    if (nt == NT && tpw == TPW) return fwd_fast1d_mma_t##TPW##_nt##NT(prof);
    BLG_MMA_ALL(BLG_MMA)
#undef BLG_MMA
    return nullptr;
}

PassKernel bwd_fast1d_mma_entry(int tpw, int nt) {
#define BLG_MMA(TPW, NT) \
    if (nt == NT && tpw == TPW) return bwd_fast1d_mma_t##TPW##_nt##NT();
    BLG_MMA_ALL(BLG_MMA)
#undef BLG_MMA
    return nullptr;
}

// interleaved kernels (fast1d_il_inst.cu): tiles of 64 cells per compute warp and chain (16 compute warps)
PassKernel fwd_fast1d_il_t1();
PassKernel fwd_fast1d_il_t2();
PassKernel bwd_fast1d_il_t1();
PassKernel bwd_fast1d_il_t2();

PassKernel fwd_fast1d_il_entry(int tpw) { return tpw == 1 ? fwd_fast1d_il_t1() : tpw == 2 ? fwd_fast1d_il_t2() : nullptr; }

PassKernel bwd_fast1d_il_entry(int tpw) {
    return tpw == 1 ? bwd_fast1d_il_t1() : tpw == 2 ? bwd_fast1d_il_t2() : nullptr;
}

}  // namespace blg
