// Microbenchmark of the register-blocked sliding-window convolution (conv_item, fast1d.cuh) in isolation: cycles per
// tap per warp as a function of resident warps per SM and outputs per thread.  Not a bench number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../bayesloop_b200/csrc -o conv_bench conv_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "fast1d.cuh"

using namespace blg;

// variant with the next chunk's weights prefetched one chunk ahead
template <int M>
__device__ __forceinline__ void conv_item_pf(const double *__restrict__ line, int i0, int R, const double *__restrict__ W,
                                             double (&acc)[M]) {
    constexpr int MP = M + 1;
    const double *p = line + (i0 - R);
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = p[m];
        acc[m] = 0.0;
    }
    p += M;
    const int chunks = (2 * R + M) / M;
    const double *wc = W;
    double w[M], wn[M];
#pragma unroll
    for (int k = 0; k < M / 2; ++k) {
        const double2 t = reinterpret_cast<const double2 *>(wc)[k];
        w[2 * k] = t.x;
        w[2 * k + 1] = t.y;
    }
    w[M - 1] = wc[M - 1];
    for (int c = 0; c < chunks; ++c) {
        wc += MP;
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
            const double2 t = reinterpret_cast<const double2 *>(wc)[k];
            wn[2 * k] = t.x;
            wn[2 * k + 1] = t.y;
        }
        wn[M - 1] = wc[M - 1];
#pragma unroll
        for (int u = 0; u < M; ++u) {
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w[u], win[(u + m) % M], acc[m]);
            win[u] = p[u];
        }
        p += M;
#pragma unroll
        for (int u = 0; u < M; ++u) w[u] = wn[u];
    }
}

__device__ __forceinline__ double vfma(double a, double b, double c) {
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}

// variant 2: program order pinned with volatile asm (tap-major: same accumulator every M instructions)
template <int M>
__device__ __forceinline__ void conv_item_v(const double *__restrict__ line, int i0, int R, const double *__restrict__ W,
                                            double (&acc)[M]) {
    constexpr int MP = M + 1;
    const double *p = line + (i0 - R);
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = p[m];
        acc[m] = 0.0;
    }
    p += M;
    const int chunks = (2 * R + M) / M;
    const double *wc = W;
    for (int c = 0; c < chunks; ++c) {
        double w[M];
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
            const double2 t = reinterpret_cast<const double2 *>(wc)[k];
            w[2 * k] = t.x;
            w[2 * k + 1] = t.y;
        }
        w[M - 1] = wc[M - 1];
#pragma unroll
        for (int u = 0; u < M; ++u) {
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = vfma(w[u], win[(u + m) % M], acc[m]);
            win[u] = p[u];
        }
        p += M;
        wc += MP;
    }
}

template <int M, int V>
__global__ void __launch_bounds__(512, 1) conv_kernel(double *out, int iters, int R, int n, int halo) {
    extern __shared__ __align__(16) double sm[];
    double *line = sm + halo;
    double *W = sm + n + 2 * halo + 32;
    const int wlen = ((2 * R + M) / M + 2) * (M + 1);
    for (int i = threadIdx.x; i < n + 2 * halo; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    for (int i = threadIdx.x; i < wlen; i += blockDim.x) W[i] = 1.0 / (2 * R + 1);
    __syncthreads();
    const int i0 = threadIdx.x * M;
    double acc[M], tot = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0)
            conv_item<M>(line, i0, R, W, acc);
        else if (V == 1)
            conv_item_pf<M>(line, i0, R, W, acc);
        else
            conv_item_v<M>(line, i0, R, W, acc);
#pragma unroll
        for (int m = 0; m < M; ++m) tot += acc[m];
        line[i0] = tot * 1e-30 + 1.0;  // keeps the loop honest (dependent store), own cell only
    }
    const long long t1 = clock64();
    if (tot == 12345.0) out[0] = tot;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}

template <int M, int V>
void run(int warps, int ctasPerSM, int sms, int R, double *d) {
    const int iters = 2000;
    const int n = warps * 32 * M, halo = R + 2 * M + 2;
    const size_t bytes = (size_t)(n + 2 * halo + 32 + ((2 * R + M) / M + 2) * (M + 1)) * 8;
    cudaFuncSetAttribute(conv_kernel<M, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    conv_kernel<M, V><<<sms * ctasPerSM, warps * 32, bytes>>>(d, 10, R, n, halo);
    conv_kernel<M, V><<<sms * ctasPerSM, warps * 32, bytes>>>(d, iters, R, n, halo);
    cudaDeviceSynchronize();
    double h[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const int taps = (2 * R + M) / M * M;
    const double cyc = h[1] / iters;
    printf("M=%2d V=%d R=%3d warps/CTA=%d CTAs/SM=%d (%2d warps/SM): %.0f cycles per conv, %.2f cycles per tap per warp, "
           "FP64 pipe %.1f%% (16 lanes/clk/SMSP)\n",
           M, V, R, warps, ctasPerSM, warps * ctasPerSM, cyc, cyc / taps,
           100.0 * (double)taps * M * 2.0 * warps * ctasPerSM / 4.0 / cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}

template <int M, int V>
void runw(int warps, int sms, int R, double *d) { run<M, V>(warps, 1, sms, R, d); }

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *d;
    cudaMalloc(&d, 1024);
    for (int R : {33, 67}) {
        run<11, 0>(1, 1, sms, R, d);
        run<11, 2>(1, 1, sms, R, d);
        run<11, 0>(4, 1, sms, R, d);
        run<11, 2>(4, 1, sms, R, d);
        run<9, 0>(4, 1, sms, R, d);
        run<9, 2>(4, 1, sms, R, d);
        run<7, 2>(4, 1, sms, R, d);
        run<5, 0>(4, 1, sms, R, d);
        run<5, 2>(4, 1, sms, R, d);
    }
    // warps per SMSP sweep inside ONE CTA per SM (co-residency certain)
    for (int w : {4, 8, 12, 16}) {
        run<11, 0>(w, 1, sms, 67, d);
        run<11, 2>(w, 1, sms, 67, d);
        run<9, 0>(w, 1, sms, 67, d);
    }
    return 0;
}
