// Microbenchmark: FP64 FMA latency / throughput per SM on this GPU as a function of warps per SM and independent
// chains per thread (ILP).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void ffma_kernel(float *out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
void run(int warps, int blocksPerSM, int sms, double *d, double clockGHz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<ILP><<<sms * blocksPerSM, warps * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<sms * blocksPerSM, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clockGHz * 1e9;
    const double warpInstrPerSM = (double)iters * ILP * warps * blocksPerSM;
    printf("DFMA ILP=%d warps/SM=%3d: %.3f ms  %.2f cycles per dependent step, %.3f warp-DFMA/cycle/SM (%.1f lanes/clk/SM)\n",
           ILP, warps * blocksPerSM, ms, cycles / iters, warpInstrPerSM / cycles, 32 * warpInstrPerSM / cycles);
}

template <int ILP>
void runf(int warps, int blocksPerSM, int sms, float *d, double clockGHz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ffma_kernel<ILP><<<sms * blocksPerSM, warps * 32>>>(d, 100, 1.0000001f, 1e-9f);
    cudaEventRecord(e0);
    ffma_kernel<ILP><<<sms * blocksPerSM, warps * 32>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clockGHz * 1e9;
    const double warpInstrPerSM = (double)iters * ILP * warps * blocksPerSM;
    printf("FFMA ILP=%d warps/SM=%3d: %.3f ms  %.2f cycles per dependent step, %.3f warp-FFMA/cycle/SM (%.1f lanes/clk/SM)\n",
           ILP, warps * blocksPerSM, ms, cycles / iters, warpInstrPerSM / cycles, 32 * warpInstrPerSM / cycles);
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
    run<1>(1, 1, sms, d, ghz);
    run<2>(1, 1, sms, d, ghz);
    run<4>(1, 1, sms, d, ghz);
    run<8>(1, 1, sms, d, ghz);
    run<8>(4, 1, sms, d, ghz);
    run<8>(8, 1, sms, d, ghz);
    run<8>(16, 1, sms, d, ghz);
    run<8>(32, 1, sms, d, ghz);
    run<1>(32, 1, sms, d, ghz);
    run<1>(32, 2, sms, d, ghz);
    run<4>(5, 4, sms, d, ghz);
    runf<8>(16, 1, sms, (float *)d, ghz);
    runf<8>(32, 1, sms, (float *)d, ghz);
    return 0;
}
