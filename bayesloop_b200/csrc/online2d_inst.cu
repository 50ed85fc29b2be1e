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
    const size_t base = (size_t)L->inRowsMax * L->P + (size_t)o2::kTH * L->P + L->w0len + L->w1len + 4 * kMaxWarps;
    const size_t limit = 232448;  // 227 KB opt-in maximum per CTA on sm_100
    // pipelined: the output tile gets a buffer of its own (the next tile loads while this one is finished)
    L->pipelined = (base + (size_t)o2::kTH * o2::kOutP) * sizeof(double) <= limit ? 1 : 0;
    L->smemBytes = (base + (L->pipelined ? (size_t)o2::kTH * o2::kOutP : 0)) * sizeof(double);
    return L->smemBytes <= limit;
}

size_t online2d_scratch_doubles(long long B, long long G, const O2Launch &L) {
    // unnormalised cells [B][G] | per-tile partial sums [B][tiles][2] | descriptors [B] | queue counter | 1 / sum [B]
    return (size_t)B * G + (size_t)B * L.tilesY * L.tilesX * 2 + (size_t)B * (sizeof(O2Hyp) / sizeof(double)) + 2 + (size_t)B;
}

int online2d_run(const PassArgs &a, const O2Launch &L, double *scratch, cudaStream_t st) {
    O2Geom geo;
    geo.tilesY = L.tilesY;
    geo.tilesX = L.tilesX;
    geo.P = L.P;
    geo.inRowsMax = L.inRowsMax;
    geo.w0len = L.w0len;
    geo.w1len = L.w1len;
    geo.pipelined = L.pipelined;
    geo.scratch = scratch;
    geo.partial = scratch + (size_t)a.B * a.pb.G;
    geo.hyp = reinterpret_cast<O2Hyp *>(geo.partial + (size_t)a.B * L.tilesY * L.tilesX * 2);
    geo.counter = reinterpret_cast<int *>(geo.hyp + a.B);
    cudaError_t e = cudaMemsetAsync(geo.counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    online2d_desc_kernel<<<(unsigned)((a.B + 127) / 128), 128, 0, st>>>(a, geo.hyp);
    // persistent CTAs, one per SM (64-row tiles, 512 threads), units of tiles from an atomic queue.  (A 32-row /
    // 256-thread variant with two CTAs per SM was measured in round 2 and dropped: 2.09 vs 1.92 ms per C5 step.)
    void (*tile)(const PassArgs, const O2Geom) = L.async ? online2d_tile_kernel<true> : online2d_tile_kernel<false>;
    e = cudaFuncSetAttribute(tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = (long long)L.tilesY * L.tilesX;
    int chunk = kO2Chunk;
    while (chunk > 1 && a.B * ((tiles + chunk - 1) / chunk) < 6LL * a.num_sms) chunk /= 2;
    geo.chunk = chunk;
    const long long units = a.B * ((tiles + chunk - 1) / chunk);
    const long long grid = units < a.num_sms ? units : a.num_sms;
    tile<<<(unsigned)grid, o2::kThreads, L.smemBytes, st>>>(a, geo);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    long long chunks = (a.pb.G + 256LL * 8 - 1) / (256LL * 8);  // 8 cells per thread
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    double *inv = reinterpret_cast<double *>(geo.counter) + 1;  // [B] behind the queue counter
    online2d_sums_kernel<<<(unsigned)((a.B * 32 + 127) / 128), 128, 0, st>>>(a, geo, inv);
    online2d_finish_kernel<<<dim3((unsigned)chunks, (unsigned)a.B), 256, 0, st>>>(a, geo, inv);
    return (int)cudaGetLastError();
}

}  // namespace blg
