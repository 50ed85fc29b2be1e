// online2d_inst.cu -- tiled OnlineStudy step (online2d.cuh): layout + launch of K7 / K8
#include "kernels.h"
#include "online2d.cuh"

namespace blg {

bool online2d_plan(int n0, int n1, int r0max, int r1max, bool async, O2Launch *L) {
    L->async = async ? 1 : 0;
    L->tilesY = (n0 + o2::kTH - 1) / o2::kTH;
    L->tilesX = (n1 + o2::kTW - 1) / o2::kTW;
    L->P = (o2::kTW + 2 * r1max) | 1;
    L->inRowsMax = o2::kTH + 2 * r0max;
    L->w0len = o2::padded_taps(r0max, o2::kM0);
    L->w1len = o2::padded_taps(r1max, o2::kM1);
    const size_t doubles = (size_t)L->inRowsMax * L->P + (size_t)o2::kTH * L->P + L->w0len + L->w1len + 4 * kMaxWarps;
    L->smemBytes = doubles * sizeof(double);
    return L->smemBytes <= 232448;  // 227 KB opt-in maximum per CTA on sm_100
}

int online2d_run(const PassArgs &a, const O2Launch &L, double *scratch, cudaStream_t st) {
    O2Geom geo;
    geo.tilesY = L.tilesY;
    geo.tilesX = L.tilesX;
    geo.P = L.P;
    geo.inRowsMax = L.inRowsMax;
    geo.w0len = L.w0len;
    geo.w1len = L.w1len;
    geo.scratch = scratch;
    geo.partial = scratch + (size_t)a.B * a.pb.G;
    // 64-row tiles, 512 threads, one CTA per SM.  (A 32-row / 256-thread variant with two CTAs per SM was measured in
    // round 2 and dropped: 2.75 ms vs 2.58 ms per C5 step without cp.async, 2.09 vs 1.92 ms with it.)
    void (*tile)(const PassArgs, const O2Geom) = L.async ? online2d_tile_kernel<true, o2::kTH, o2::kThreads>
                                                         : online2d_tile_kernel<false, o2::kTH, o2::kThreads>;
    const int threads = o2::kThreads;
    cudaError_t e = cudaFuncSetAttribute(tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes);
    if (e != cudaSuccess) return (int)e;
    const unsigned tiles = (unsigned)(L.tilesY * L.tilesX);
    tile<<<(unsigned)a.B * tiles, threads, L.smemBytes, st>>>(a, geo);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    long long chunks = (a.pb.G + 256LL * 8 - 1) / (256LL * 8);  // 8 cells per thread
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    online2d_finish_kernel<<<dim3((unsigned)chunks, (unsigned)a.B), 256, 0, st>>>(a, geo);
    return (int)cudaGetLastError();
}

}  // namespace blg
