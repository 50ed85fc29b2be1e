// dsmem_bench.cu -- how fast can one CTA push a block of shared memory into a neighbour CTA of its cluster?
//   (a) st.async.shared::cluster ... v2.f64 by all threads (what cluster2d.cuh's publish phase does)
//   (b) cp.async.bulk.shared::cluster.shared::cta (one bulk copy issued by one thread, mbarrier completion at the receiver)
// Every CTA of an 8-CTA cluster sends `bytes` to rank+1 and rank-1 simultaneously (the halo-exchange pattern).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu && ./dsmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, unsigned r) {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(512, 1) push_kernel(int bytes, int iters, long long *out) {
    extern __shared__ __align__(16) double sm[];
    // layout: [src: bytes][dst_up: bytes][dst_dn: bytes][mbar x2]
    double *src = sm, *dstUp = sm + bytes / 8, *dstDn = sm + 2 * (bytes / 8);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 3 * (bytes / 8));
    unsigned rank, size;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(size));
    for (int i = threadIdx.x; i < 3 * (bytes / 8); i += blockDim.x) sm[i] = 1.0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const unsigned up = (rank + size - 1) % size, dn = (rank + 1) % size;
    const uint32_t remUp = mapa(smem_u32(dstDn), up), remDn = mapa(smem_u32(dstUp), dn);  // my rows land in their halo
    const uint32_t barUp = mapa(smem_u32(&bar[0]), up), barDn = mapa(smem_u32(&bar[0]), dn);
    long long t0 = clock64();
    uint32_t parity = 0;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(2 * bytes) : "memory");
        if (MODE == 0) {
            for (int e = threadIdx.x; e < bytes / 16; e += blockDim.x) {
                double2 v = reinterpret_cast<double2 *>(src)[e];
                asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remUp + 16 * e),
                             "d"(v.x), "d"(v.y), "r"(barUp)
                             : "memory");
                asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remDn + 16 * e),
                             "d"(v.x), "d"(v.y), "r"(barDn)
                             : "memory");
            }
        } else {
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remUp),
                             "r"(smem_u32(src)), "r"(bytes), "r"(barUp)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remDn),
                             "r"(smem_u32(src)), "r"(bytes), "r"(barDn)
                             : "memory");
            }
        }
        mbar_wait(smem_u32(&bar[0]), parity);  // both neighbours' blocks have landed here
        parity ^= 1;
        // everybody has received before anybody sends again (keeps the phases of the single mbarrier apart)
        asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    long long *d;
    cudaMalloc(&d, 1024 * sizeof(long long));
    const int iters = 200;
    for (int kb : {4, 16, 34, 64}) {
        const int bytes = kb * 1024;
        const size_t smem = 3 * (size_t)bytes + 64;
        for (int mode = 0; mode < 2; ++mode) {
            auto k = mode ? push_kernel<1> : push_kernel<0>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<8 * 15, 512, smem>>>(bytes, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[120];
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < 120; ++i) mean += (double)h[i] / 120;
            printf("%s  %2d KB up + %2d KB down per CTA: %8.0f cycles per exchange  -> %.1f B/clk sent per SM  (%s)\n",
                   mode ? "cp.async.bulk smem->dsmem" : "st.async v2.f64 (512 thr) ", kb, kb, mean / iters, 2.0 * bytes / (mean / iters),
                   cudaGetErrorString(e));
        }
    }
    return 0;
}
