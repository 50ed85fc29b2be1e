// Microbenchmark: does the FP64 tensor instruction (mma.sync.m8n8k4.f64 = "DMMA", 256 MAC per warp instruction) run on
// a pipe of its own next to DFMA on B200, or do the two share the FP64 units?  Per SM: NW warps, each thread runs
// `iters` rounds of NF independent DFMAs and NM independent DMMAs.  If the pipes are separate,
// time(NF, NM) ~ max(time(NF, 0), time(0, NM)); if shared, ~ the sum.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix_bench dmma_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void mix_kernel(double *out, int iters, double a, double b) {
    double acc[NF > 0 ? NF : 1];
    double c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) c0[i] = c1[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NF) acc[i] = fma(acc[i], a, b);
            if (i < NM) dmma(c0[i], c1[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

// warp-specialised mix: even warps DFMA only, odd warps DMMA only
template <int N>
__global__ void split_kernel(double *out, int iters, double a, double b) {
    double acc[N], c1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = c1[i] = threadIdx.x * 1e-9 + i;
    if ((threadIdx.x >> 5) & 4) {  // warps 4-7 (one per sub-partition): DMMA
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < N; ++i) dmma(acc[i], c1[i], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int NF, int NM>
void run(int warps, int sms, double *d, double ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix_kernel<NF, NM><<<sms, warps * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    mix_kernel<NF, NM><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9 / iters;
    const double mac = (double)warps * (32.0 * NF + 256.0 * NM);
    printf("warps/SM=%2d  DFMA=%2d DMMA=%2d per round: %8.1f cycles/round  %.1f MAC/clk/SM (DFMA part %.1f, DMMA part %.1f)\n",
           warps, NF, NM, cyc, mac / cyc, warps * 32.0 * NF / cyc, warps * 256.0 * NM / cyc);
}

template <int N>
void run_split(int sms, double *d, double ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    split_kernel<N><<<sms, 256>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    split_kernel<N><<<sms, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9 / iters;
    printf("split: 4 DFMA warps + 4 DMMA warps, %d per round each: %8.1f cycles/round (DFMA warps alone need %.1f, DMMA warps alone: see DMMA-only line x%d/8)\n",
           N, cyc, 4.0 * N * 32 / 58.5, N);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal)\n", p.name, p.multiProcessorCount, ghz);
    double *d;
    cudaMalloc(&d, 1024);
    const int sms = p.multiProcessorCount;
    run<8, 0>(8, sms, d, ghz);
    run<8, 0>(16, sms, d, ghz);
    run<0, 1>(8, sms, d, ghz);
    run<0, 4>(8, sms, d, ghz);
    run<0, 8>(8, sms, d, ghz);
    run<0, 8>(16, sms, d, ghz);
    run<8, 1>(8, sms, d, ghz);
    run<8, 2>(8, sms, d, ghz);
    run<8, 4>(8, sms, d, ghz);
    run<8, 8>(8, sms, d, ghz);
    run<8, 8>(16, sms, d, ghz);
    run_split<8>(sms, d, ghz);
    return 0;
}
