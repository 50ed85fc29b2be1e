// Microbenchmark: do FP64 FMAs and 64-bit shared-memory loads overlap on one SM?  Each thread runs `iters` rounds of
// NF independent DFMAs and NL independent LDS.64 (conflict-free, stride 9 doubles).  If the pipes overlap, time(NF,NL)
// ~ max(time(NF,0), time(0,NL)); if they serialise, ~ the sum.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NL>
__global__ void mix_kernel(double *out, int iters, double a, double b) {
    __shared__ double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i * 1e-9;
    __syncthreads();
    double acc[NF > 0 ? NF : 1];
    double ld[NL > 0 ? NL : 1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < (NL > 0 ? NL : 1); ++i) ld[i] = 0.0;
    int base = (threadIdx.x * 9) & 2047;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NL; ++i) ld[i] += buf[(base + i * 131 + it) & 4095];
#pragma unroll
        for (int i = 0; i < NF; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < (NL > 0 ? NL : 1); ++i) s += ld[i];
    if (s == 12345.678) out[0] = s;
}

template <int NF, int NL>
void run(int warps, int sms, double *d, double ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix_kernel<NF, NL><<<sms, warps * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    mix_kernel<NF, NL><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9 / iters;
    printf("warps/SM=%2d  DFMA=%2d LDS.64=%2d per round: %.1f cycles/round  (DFMA alone would need %.1f, LDS alone %.1f at 2 wavefronts each)\n",
           warps, NF, NL, cyc, warps * NF / 1.83, warps * NL * 2.0);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    double *d;
    cudaMalloc(&d, 1024);
    const int sms = p.multiProcessorCount;
    for (int warps : {4, 16}) {
        run<9, 0>(warps, sms, d, ghz);
        run<0, 2>(warps, sms, d, ghz);
        run<9, 2>(warps, sms, d, ghz);
        run<0, 4>(warps, sms, d, ghz);
        run<9, 4>(warps, sms, d, ghz);
        run<18, 4>(warps, sms, d, ghz);
    }
    return 0;
}
