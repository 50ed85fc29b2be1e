// kernels.h -- host-visible entry points of the pass kernels.  Every kernel family is compiled in its own
// translation unit (fast1d_inst.cu, resident_inst.cu, cluster2d_inst.cu ...) so that the library builds in parallel;
// api.cu only sees these function pointers (all pass kernels share the signature void(const PassArgs)).
#pragma once

#include "common.cuh"

namespace blg {

using PassKernel = void (*)(const PassArgs);

// single-role fused 1-D kernels (fast1d.cuh): M = 9 outputs per thread (the fallback for grids the warp-specialised
// kernels do not cover); nt = threads of the launch (<= 128, <= 160, <= 256, else 1024).  nullptr: not compiled.
PassKernel fwd_fast1d_entry(int M, int nt);
PassKernel bwd_fast1d_entry(int M, int nt);

// warp-specialised fast 1-D kernels (fast1d_ws.cuh): nt = 160 (4 compute warps) or 288 (8 compute warps); the compute
// warps own M cells per thread (ML <= M in the last one), one service warp normalises / stores / prefetches.
// fast1d_ws_geometries: the compiled {M, ML, nt} triples, terminated by M = 0.
PassKernel fwd_fast1d_ws_entry(int M, int ML, int nt);
PassKernel bwd_fast1d_ws_entry(int M, int ML, int nt);
const int *fast1d_ws_geometries();

// the same kernels with the convolution on FP64 matrix instructions (fast1d_mma.cuh): tpw tiles of 64 cells per compute
// warp (1..6), nt = 160 (4 compute warps) or 288 (8, tpw >= 4)
PassKernel fwd_fast1d_mma_entry(int tpw, int nt, bool prof = false);
PassKernel bwd_fast1d_mma_entry(int tpw, int nt);

// ... and with the chains of an SM interleaved inside one CTA of 16 compute warps + 1 service warp (fast1d_il.cuh):
// tpw tiles per compute warp and chain (1: grids up to 1024 cells, 2: up to 2048); grid = one CTA per sm_assign list
PassKernel fwd_fast1d_il_entry(int tpw);
PassKernel bwd_fast1d_il_entry(int tpw);

// generic resident kernels (resident.cuh): nt in {256, 512, 1024}; stream = state in global scratch (1024 threads)
PassKernel fwd_resident_entry(int nt, bool stream);
PassKernel bwd_resident_entry(int nt, bool stream);

// which transition programs the cluster-resident 2-D kernels understand (classify2d in cluster2d.cuh)
bool cluster2d_supports(int n_ops, const int *kind, const int *axis);

// cluster-resident 2-D kernels (cluster2d.cuh): 512 threads, one cluster of 2/4/8 CTAs per combo
// prof: per-phase cycle counters into PassArgs::trace; m0 in {16, 13}: rows per axis-0 work item
PassKernel fwd_cluster2d_entry(bool prof, int m0);
PassKernel bwd_cluster2d_entry(bool prof, int m0);
// layout parameters of cluster2d.cuh: threads, rows per axis-0 item, cells per axis-1 item, cells per thread
void cluster2d_params(int *threads, int *m0, int *m1, int *cells, int *wpad);

// tiled OnlineStudy step (online2d.cuh): 64 x 64 tiles x hypotheses, T = 1, rows promised separable by the caller
// (BLG_F_SEPARABLE_ROWS).  online2d_plan: false when tile + halo do not fit in shared memory; scratch holds
// online2d_scratch_doubles(B, G, L) doubles; online2d_run returns a cudaError_t (0 = launched the three kernels).
struct O2Launch {
    int async;  // tile loads through cp.async (the default; measured 1.92 vs 2.58 ms per C5 step, profiles/r2a_online_ab.txt)
    int pipelined;  // the output tile has its own shared-memory buffer (set by online2d_plan when it fits)
    int tilesY, tilesX, P, inRowsMax, w0len, w1len;
    size_t smemBytes;
};
bool online2d_plan(int n0, int n1, int r0max, int r1max, bool async, O2Launch *L);
size_t online2d_scratch_doubles(long long B, long long G, const O2Launch &L);
int online2d_run(const PassArgs &a, const O2Launch &L, double *scratch, cudaStream_t st);

}  // namespace blg
