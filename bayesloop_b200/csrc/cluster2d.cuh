// cluster2d.cuh -- K5/K6: forward / backward passes for 2-D grids that do not fit in ONE SM's shared memory
// (200^2, 256^2: BASELINE.json configs[2], [3]) with the state resident ON CHIP across a thread-block cluster.
//
// One cluster of C CTAs (C = 2, 4 or 8, one CTA per SM) owns one hyper-parameter combination for all T steps.  The
// grid is cut into C bands of rows (axis 0); every CTA keeps its band of the (unnormalised) state in shared memory
// for the whole recursion, so the only HBM traffic is the compulsory one: the alpha[t] row store (forward), the
// alpha[t] row load + posterior store (backward) and the likelihood row shared by all combos.  Per time step:
//
//   0   GRW on axis 0 (across bands): every thread convolves M0 rows of one column out of shared memory into
//       registers; the R0 rows it needs from the neighbouring bands were PUSHED into this CTA's halo rows through
//       distributed shared memory by the neighbours at the end of the previous step
//   1   GRW on axis 1 (inside a row): M1 cells of one row per thread, reflect boundary by index
//   E   elementwise: prior x likelihood (forward) / alpha x beta, beta x likelihood (backward), partial sums
//   P   publish: halo rows and the per-CTA partial sums (evidence increment; sum(alpha beta), sum(beta),
//       sum(post/lik)) travel as st.async stores that complete transaction bytes on an mbarrier IN THE RECEIVING CTA;
//       the receiver waits on its own mbarrier -- no cluster-wide barrier and no memory fence on the critical path.
//       The only cluster barrier per step is a split one (relaxed arrive after the axis-0 reads, wait before the
//       pushes) that protects the halo rows; it is hidden behind the axis-1 convolution.
//
// HBM rows: the forward pass stages the likelihood band of the NEXT step with one bulk-async (TMA) copy and, when
// the caller allows unnormalised rows (BLG_F_RAW_ALPHA: a backward pass follows, which is scale-free per row), stores
// alpha[t] with one bulk-async copy straight out of the state buffer; the backward pass receives alpha[t-1] by TMA into
// the band whose smoothed row u = alpha * beta has just left by TMA (BLG_F_RAW_POSTERIOR: unnormalised, its factor
// 1/sum(u) goes to row_scale[b][t]; otherwise normalised in place first).  With raw rows no phase of a step needs the
// cluster-wide sums of the previous one before its elementwise sweep, so their all-to-all hides behind the
// convolutions.  The likelihood row of the backward pass comes through registers (coalesced loads issued after the
// last convolution, L2-prefetched two steps ahead by a bulk prefetch): a second staging band does not fit at 256^2.
//
// Both convolutions commute (separable, linear), so axis 0 always runs first whatever the program order; the results
// agree with the reference order to rounding (1e-16 relative).  Semantics: core.py:372-417, :434-470,
// transitionModels.py:96-115 (GaussianRandomWalk), :300-312 (ChangePoint reset before / after the random walks).
#pragma once

#include "common.cuh"

namespace blg {

struct C2Ops {
    int k0, k1;        // program index of the GRW acting on axis 0 / axis 1 (-1: none)
    int pre, post;     // program index of a RESET before all / after all GRWs (-1: none)
    bool ok;
};

// Host and device share this classification: which programs the cluster-resident 2-D kernels understand.
__host__ __device__ inline C2Ops classify2d(int n_ops, const int *kind, const int *axis) {
    C2Ops o;
    o.k0 = o.k1 = o.pre = o.post = -1;
    o.ok = true;
    int firstGrw = -1, lastGrw = -1;
    for (int k = 0; k < n_ops; ++k)
        if (kind[k] == BLG_OP_GRW) {
            if (firstGrw < 0) firstGrw = k;
            lastGrw = k;
            if (axis[k] == 0 && o.k0 < 0)
                o.k0 = k;
            else if (axis[k] == 1 && o.k1 < 0)
                o.k1 = k;
            else
                o.ok = false;
        }
    for (int k = 0; k < n_ops; ++k) {
        if (kind[k] == BLG_OP_GRW) continue;
        if (kind[k] != BLG_OP_RESET) {
            o.ok = false;
            continue;
        }
        if (firstGrw < 0 || k > lastGrw) {
            if (o.post < 0)
                o.post = k;
            else
                o.ok = false;
        } else if (k < firstGrw) {
            if (o.pre < 0)
                o.pre = k;
            else
                o.ok = false;
        } else {
            o.ok = false;
        }
    }
    return o;
}

constexpr int kC2Threads = 512;
constexpr int kC2M0 = 16;      // rows per work item of the axis-0 convolution (template parameter M0: 16 or 13, whichever
                               // wastes fewer rows of the band: 32-row bands -> 16, 25-row bands -> 13)
constexpr int kC2M1 = 17;      // cells per work item of the axis-1 convolution (odd: conflict-free 64-bit LDS)
constexpr int kC2Cells = 16;   // cells per thread of the elementwise phases (band <= 16 * 512 cells)
constexpr int kC2MaxCluster = 8;
constexpr int kC2WPad = 4;     // zero taps after a weight table (the tap loops fetch one weight ahead)
// misc region (doubles): [0,128) RedScratch of build_weights, [128,176) per-warp partials [3][16],
// [176,224) cluster slots [2][3][8], [224,226) "halo rows have arrived" mbarriers by step parity, [226,228) "partial
// sums have arrived" mbarriers by step parity, [228] TMA mbarrier
constexpr int kC2WarpPart = 128, kC2Slots = 176, kC2HaloBar = 224, kC2SumBar = 226, kC2Mbar = 228;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned c2_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned c2_cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void c2_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without memory ordering: used where only "my shared-memory READS are done" has to be signalled (the values
// were consumed by arithmetic before the arrive); saves the MEMBAR.ALL.GPU of the releasing form
__device__ __forceinline__ void c2_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void c2_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t c2_map(uint32_t saddr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// shared-memory accesses by 32-bit address (keeps generic-pointer arithmetic out of the hot loops).  Volatile keeps
// them ordered among themselves and with the barriers; no "memory" clobber on purpose: it would force every
// address-taken register array (convolution windows, sweep batches) into local memory.  Plain C++ accesses to the
// same buffers are always separated from these by a CTA barrier.
__device__ __forceinline__ double c2_lds(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 c2_lds2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void c2_sts(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}
// store into (possibly another CTA's) shared memory; the bytes complete on the mbarrier `rbar` of the SAME target CTA
__device__ __forceinline__ void c2_st_async(uint32_t raddr, double v, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void c2_st_async2(uint32_t raddr, double2 v, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "d"(v.x), "d"(v.y), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void c2_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void c2_prefetch_l2(const void *g, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------------ convolutions
// M outputs of one COLUMN.  `col`: address of row (i0 - R) of the column; rows are `pitchB` bytes apart; loads are
// clamped to `last` (the column's cell in the last row of the buffer).  Weights W[0 .. chunks*M] with zeros after tap
// 2R, so the tap loop needs no guards; one weight is fetched a tap ahead.
template <int M>
__device__ __forceinline__ void c2_conv_col(uint32_t col, uint32_t pitchB, uint32_t last, int taps, uint32_t W,
                                            double (&acc)[M]) {
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = c2_lds(min(col + (uint32_t)m * pitchB, last));
        acc[m] = 0.0;
    }
    uint32_t p = col + (uint32_t)M * pitchB;
    double w = c2_lds(W);
    const int full = taps / M, rem = taps - full * M;
    for (int c = 0; c < full; ++c) {
#pragma unroll
        for (int u = 0; u < M; ++u) {
            const double wn = c2_lds(W + 8u * (uint32_t)(u + 1));
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
            win[u] = c2_lds(min(p + (uint32_t)u * pitchB, last));
            w = wn;
        }
        p += (uint32_t)M * pitchB;
        W += 8u * (uint32_t)M;
    }
#pragma unroll
    for (int u = 0; u < M - 1; ++u) {  // remainder taps (uniform guards)
        if (u < rem) {
            const double wn = c2_lds(W + 8u * (uint32_t)(u + 1));
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
            win[u] = c2_lds(min(p + (uint32_t)u * pitchB, last));
            w = wn;
        }
    }
}

// Address of cell i of a row under NI_EXTEND_REFLECT, valid for -n <= i < 2n (guaranteed by R + M <= n):
// i < 0 -> -1-i = ~i,  i >= n -> 2n-1-i,  i.e. min(max(i, ~i), 2n-1-i).
__device__ __forceinline__ uint32_t c2_reflect_addr(uint32_t row, int i, int n2m1) {
    return row + 8u * (uint32_t)min(max(i, ~i), n2m1 - i);
}

// M outputs of one ROW starting at cell i0 (`row`: address of cell 0), R + M <= n.
template <int M>
__device__ __forceinline__ void c2_conv_row(uint32_t row, int i0, int n, int R, int taps, uint32_t W, double (&acc)[M]) {
    double win[M];
    int e = i0 - R;
    n = 2 * n - 1;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = c2_lds(c2_reflect_addr(row, e + m, n));
        acc[m] = 0.0;
    }
    e += M;
    double w = c2_lds(W);
    const int full = taps / M, rem = taps - full * M;
    for (int c = 0; c < full; ++c) {
#pragma unroll
        for (int u = 0; u < M; ++u) {
            const double wn = c2_lds(W + 8u * (uint32_t)(u + 1));
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
            win[u] = c2_lds(c2_reflect_addr(row, e + u, n));
            w = wn;
        }
        e += M;
        W += 8u * (uint32_t)M;
    }
#pragma unroll
    for (int u = 0; u < M - 1; ++u) {  // remainder taps (uniform guards)
        if (u < rem) {
            const double wn = c2_lds(W + 8u * (uint32_t)(u + 1));
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
            win[u] = c2_lds(c2_reflect_addr(row, e + u, n));
            w = wn;
        }
    }
}

// ------------------------------------------------------------------------------------------------ per-combo state
// Kept deliberately small: the convolutions need ~85 registers per thread and everything that lives across them
// competes with the 128-register budget of a 512-thread CTA (spilled values are reloaded through an L1 that every
// cluster-barrier wait invalidates).  Geometry is re-read from the kernel parameters (constant bank) where needed.
extern __shared__ __align__(16) double c2_smem[];
__device__ __forceinline__ double *c2_X(const PassArgs &a) { return c2_smem + a.c2_off_x; }
__device__ __forceinline__ double *c2_Xb(const PassArgs &a) { return c2_smem + a.c2_off_x + a.c2_h0 * a.pb.n1; }
__device__ __forceinline__ double *c2_S(const PassArgs &a) { return c2_smem + a.c2_off_s; }
__device__ __forceinline__ double *c2_misc(const PassArgs &a) { return c2_smem + a.off_misc; }

struct C2 {
    uint32_t xAddr, xbAddr, sAddr, w0Addr, w1Addr;
    double parPre, parPost;
    int R0, R1;
    int lo0, hi0, lo1, hi1, loPre, hiPre, loPost, hiPost;  // active windows of this pass direction (empty: absent)
    int rank, C, r0, nb, cnt;
    uint32_t bits;  // bit 0: parity of the publish phase in flight, bit 1: parity of the sums not yet collected,
                    // bits 2,3: phase of the halo mbarriers, bits 4,5: phase of the sum mbarriers
};

__device__ __forceinline__ bool c2_in(int lo, int hi, long long idx) { return idx >= (long long)lo && idx < (long long)hi; }

// Geometry, per-combo parameters, weights.  Returns false (uniformly over the cluster) if a radius exceeds what the
// host sized the layout for.
template <bool BWD>
__device__ __forceinline__ bool c2_setup(const PassArgs &a, long long b, C2 &s) {
    const int n0 = a.pb.n0, n1 = a.pb.n1, nbmax = a.c2_nb;
    s.rank = (int)c2_cluster_rank();
    s.C = (int)c2_cluster_size();
    s.r0 = s.rank * nbmax;
    s.nb = min(nbmax, n0 - s.r0);
    s.cnt = s.nb * n1;
    s.xAddr = smem_u32(c2_X(a));
    s.xbAddr = smem_u32(c2_Xb(a));
    s.sAddr = smem_u32(c2_S(a));
    s.bits = 0u;
    const C2Ops ops = classify2d(a.pg.n_ops, a.pg.kind, a.pg.axis);
    const int NK = a.pg.n_ops;
    auto load = [&](int k, double &par, int &rad, int &lo, int &hi) {
        par = 0.0;
        rad = 0;
        lo = hi = 0;
        if (k < 0) return;
        par = a.pg.param[b * NK + k];
        rad = a.pg.radius[b * NK + k];
        lo = a.pg.window[(b * NK + k) * 4 + (BWD ? 2 : 0)];
        hi = a.pg.window[(b * NK + k) * 4 + (BWD ? 3 : 1)];
    };
    int dummy;
    double sig0, sig1;
    load(ops.k0, sig0, s.R0, s.lo0, s.hi0);
    load(ops.k1, sig1, s.R1, s.lo1, s.hi1);
    load(ops.pre, s.parPre, dummy, s.loPre, s.hiPre);
    load(ops.post, s.parPost, dummy, s.loPost, s.hiPost);
    if (!(sig0 > 0.0) || s.R0 <= 0) s.R0 = 0;  // transitionModels.py:110-113: identity
    if (!(sig1 > 0.0) || s.R1 <= 0) s.R1 = 0;
    double *W0 = c2_smem + a.off_w + (ops.k0 >= 0 ? a.pg.w_off[ops.k0] : 0);
    double *W1 = c2_smem + a.off_w + (ops.k1 >= 0 ? a.pg.w_off[ops.k1] : 0);
    s.w0Addr = smem_u32(W0);
    s.w1Addr = smem_u32(W1);
    // keep the five shared-memory addresses in registers: rematerialising them costs a special-register read
    asm volatile("" : "+r"(s.xAddr), "+r"(s.xbAddr), "+r"(s.sAddr), "+r"(s.w0Addr), "+r"(s.w1Addr));
    bool ok = true;
    if (s.R0 > a.c2_h0 || s.R0 > n0 - (s.C - 1) * nbmax) ok = false;
    if (s.R1 + kC2M1 > n1) ok = false;
    RedScratch rs;
    rs.buf = c2_misc(a);
    rs.phase = 0;
    if (ops.k0 >= 0 && s.R0 > 0) {
        if (2 * s.R0 + 3 > a.pg.w_len[ops.k0])
            ok = false;
        else
            build_weights(W0, a.pg.w_len[ops.k0], sig0, s.R0, rs);
    }
    if (ops.k1 >= 0 && s.R1 > 0) {
        if (2 * s.R1 + 3 > a.pg.w_len[ops.k1])
            ok = false;
        else
            build_weights(W1, a.pg.w_len[ops.k1], sig1, s.R1, rs);
    }
    return ok;
}

__device__ __forceinline__ uint64_t *c2_halo_bar(const PassArgs &a, int parity) {
    return reinterpret_cast<uint64_t *>(c2_misc(a) + kC2HaloBar) + parity;
}
__device__ __forceinline__ uint64_t *c2_sum_bar(const PassArgs &a, int parity) {
    return reinterpret_cast<uint64_t *>(c2_misc(a) + kC2SumBar) + parity;
}

// Publish phase (after the elementwise sweep): per-warp partials -> CTA barrier -> (1) halo rows of the new state to
// the neighbours (or, at the grid boundary, reflected into this CTA's own halo rows), 16 bytes per st.async;
// (2) this CTA's K sums into slot[rank] of every CTA.  Everything completes transaction bytes on the receiver's
// mbarrier of this step's parity.  `afterSync` runs on all threads right after the CTA barrier (state band final).
template <int K, typename AfterSync>
__device__ __forceinline__ void c2_publish(const PassArgs &a, C2 &s, double (&v)[K], AfterSync afterSync) {
    double *wpart = c2_misc(a) + kC2WarpPart;
    const int par = (int)(s.bits & 1u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double t = warp_sum(v[k]);
        if (lane == 0) wpart[k * 16 + warp] = t;
    }
    __syncthreads();
    const uint32_t barLocal = smem_u32(c2_halo_bar(a, par)), sumLocal = smem_u32(c2_sum_bar(a, par));
    if (threadIdx.x == 0) {
        c2_expect_tx(c2_sum_bar(a, par), (uint32_t)(s.C * K * sizeof(double)));
        if (s.R0 > 0) c2_expect_tx(c2_halo_bar(a, par), (uint32_t)(2 * s.R0 * a.pb.n1 * sizeof(double)));
    }
    afterSync();
    // the partial sums first (their all-to-all is the longer chain), from the last warp: warp 0 issues the TMA copies
    if (warp == kC2Threads / 32 - 1 && lane < s.C) {
        double *slots = c2_misc(a) + kC2Slots + par * 3 * kC2MaxCluster;
        const uint32_t rbar = c2_map(sumLocal, (unsigned)lane);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int w = 0; w < kC2Threads / 32; w += 2) {
                t0 += wpart[k * 16 + w];
                t1 += wpart[k * 16 + w + 1];
            }
            c2_st_async(c2_map(smem_u32(&slots[k * kC2MaxCluster + s.rank]), (unsigned)lane), t0 + t1, rbar);
        }
    }
    if (s.R0 > 0) {
        const int n1 = a.pb.n1, pairs = s.R0 * n1 / 2;
        const bool top = s.rank == 0, bottom = s.rank == s.C - 1;
        const unsigned upRank = top ? (unsigned)s.rank : (unsigned)(s.rank - 1);
        const unsigned dnRank = bottom ? (unsigned)s.rank : (unsigned)(s.rank + 1);
        const uint32_t upX = c2_map(s.xAddr, upRank), upBar = c2_map(barLocal, upRank);
        const uint32_t dnX = c2_map(s.xAddr, dnRank), dnBar = c2_map(barLocal, dnRank);
        const uint32_t rowB = (uint32_t)n1 * 8u;
        const uint32_t srcDn = s.xbAddr + (uint32_t)(s.nb - s.R0) * rowB;
        for (int e = threadIdx.x; e < pairs; e += kC2Threads) {
            const int k = (2 * e) / n1;
            const uint32_t cB = (uint32_t)(2 * e - k * n1) * 8u;
            // rows [0, R0) of the band -> rows nb .. nb+R0-1 of the upper neighbour (its band has nbmax rows);
            // top CTA: reflect, its own row -1-k = row k
            const double2 a0 = c2_lds2(s.xbAddr + 16u * (uint32_t)e);
            const uint32_t dUp = top ? (uint32_t)(a.c2_h0 - 1 - k) * rowB + cB
                                     : (uint32_t)(a.c2_h0 + a.c2_nb) * rowB + 16u * (uint32_t)e;
            c2_st_async2(upX + dUp, a0, upBar);
            // rows [nb-R0, nb) -> rows -R0 .. -1 of the lower neighbour; bottom CTA: reflect, row nb+j = row nb-1-j
            const double2 a1 = c2_lds2(srcDn + 16u * (uint32_t)e);
            const uint32_t dDn = bottom ? (uint32_t)(a.c2_h0 + s.nb + s.R0 - 1 - k) * rowB + cB
                                        : (uint32_t)(a.c2_h0 - s.R0) * rowB + 16u * (uint32_t)e;
            c2_st_async2(dnX + dDn, a1, dnBar);
        }
    }
    s.bits ^= 1u;
}

// Wait until the halo rows of the last publish phase have landed in this CTA.
__device__ __forceinline__ void c2_collect_halo(const PassArgs &a, C2 &s) {
    const int par = (int)((s.bits & 1u) ^ 1u);
    if (s.R0 > 0) {
        mbar_wait(c2_halo_bar(a, par), (s.bits >> (2 + par)) & 1u);
        s.bits ^= 4u << par;
    }
}

// Wait until the C x K partial sums of the oldest uncollected publish phase have landed, then add the slots in rank
// order (bit-identical in every CTA).
template <int K>
__device__ __forceinline__ void c2_collect(const PassArgs &a, C2 &s, double (&v)[K]) {
    const int par = (int)((s.bits >> 1) & 1u);
    mbar_wait(c2_sum_bar(a, par), (s.bits >> (4 + par)) & 1u);
    s.bits ^= (16u << par) | 2u;
    const double *slots = c2_misc(a) + kC2Slots + par * 3 * kC2MaxCluster;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double tot = 0.0;
        for (int j = 0; j < s.C; ++j) tot += slots[k * kC2MaxCluster + j];
        v[k] = tot;
    }
}

// Overwrite the band and its halo rows with reset_base * par (ChangePoint listed before the random walks): the halo
// rows come straight from global memory, no exchange needed.
__device__ __forceinline__ void c2_reset_band(const PassArgs &a, const C2 &s, double par) {
    const double *rb = a.reset_base;
    for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) c2_Xb(a)[g] = __ldg(rb + (size_t)s.r0 * a.pb.n1 + g) * par;
    const int per = s.R0 * a.pb.n1;
    for (int e = threadIdx.x; e < 2 * per; e += kC2Threads) {
        const int side = e >= per ? 1 : 0, q = e - side * per, k = q / a.pb.n1, c = q - k * a.pb.n1;
        const int grow = reflect_any(side ? s.r0 + s.nb + k : s.r0 - 1 - k, a.pb.n0);
        const int lrow = side ? a.c2_h0 + s.nb + k : a.c2_h0 - 1 - k;
        c2_X(a)[(size_t)lrow * a.pb.n1 + c] = __ldg(rb + (size_t)grow * a.pb.n1 + c) * par;
    }
    __syncthreads();
}

// The two convolutions of one step, in place.  `beforeWrite` runs on all threads before the first CTA barrier that
// precedes an overwrite of the band (a bulk store may still be reading it); `afterConv` runs after the arithmetic of
// the last convolution (early issue of loads consumed by the elementwise phase).  Returns with all threads
// synchronised on the new band; the split cluster barrier "halo rows consumed" has been ARRIVED at (the caller
// waits before it pushes).  Returns true if the band was overwritten (beforeWrite has run).
struct C2Identity {
    __device__ __forceinline__ double operator()(uint32_t, double v) const { return v; }
};

// `epi(offset, value)` maps every output of the axis-1 convolution on its way back into the band (offset = byte offset
// of the cell inside the band): the forward pass fuses prior x likelihood and the partial sums of the step there.
template <bool TIMED, int M0, typename BeforeWrite, typename AfterConv, typename AfterWrite, typename Epi = C2Identity>
__device__ __forceinline__ bool c2_transition(const PassArgs &a, const C2 &s, bool act0, bool act1, long long &tmid,
                                              BeforeWrite beforeWrite, AfterConv afterConv, AfterWrite afterWrite,
                                              Epi epi = Epi()) {
    const int n1 = a.pb.n1;
    const uint32_t rowB = (uint32_t)n1 * 8u;
    if (act0) {
        const int S0 = (a.c2_nb + M0 - 1) / M0;
        const int w = threadIdx.x;
        const bool has = w < n1 * S0;
        const int seg = w / n1, c = w - seg * n1, i0 = seg * M0;
        double acc[M0];
        if (has)
            c2_conv_col<M0>(s.xAddr + (uint32_t)(a.c2_h0 + i0 - s.R0) * rowB + 8u * (uint32_t)c, rowB,
                               s.xAddr + (uint32_t)(a.c2_rows - 1) * rowB + 8u * (uint32_t)c, 2 * s.R0 + 1, s.w0Addr, acc);
        beforeWrite();
        __syncthreads();
        c2_arrive_relaxed();  // halo rows consumed
        if (has) {
            const uint32_t out = s.xbAddr + (uint32_t)i0 * rowB + 8u * (uint32_t)c;
#pragma unroll
            for (int m = 0; m < M0; ++m)
                if (i0 + m < s.nb) c2_sts(out + (uint32_t)m * rowB, acc[m]);
        }
        __syncthreads();
    } else {
        c2_arrive_relaxed();
    }
    tmid = TIMED ? clock64() : 0;
    if (act1) {
        const int S1 = (n1 + kC2M1 - 1) / kC2M1;
        const int w = threadIdx.x;
        const int l = w / S1, sg = w - l * S1, i0 = sg * kC2M1;
        const bool has = l < s.nb;
        double acc[kC2M1];
        if (has) c2_conv_row<kC2M1>(s.xbAddr + (uint32_t)l * rowB, i0, n1, s.R1, 2 * s.R1 + 1, s.w1Addr, acc);
        afterConv();
        if (!act0) beforeWrite();
        __syncthreads();
        if (has) {
            const uint32_t off = (uint32_t)l * rowB + 8u * (uint32_t)i0;
#pragma unroll
            for (int m = 0; m < kC2M1; ++m)
                if (i0 + m < n1) c2_sts(s.xbAddr + off + 8u * (uint32_t)m, epi(off + 8u * (uint32_t)m, acc[m]));
        }
        afterWrite();  // the accumulators are dead: room for the second batch of early loads
        __syncthreads();
    } else {
        afterConv();
        afterWrite();
    }
    return act0 || act1;
}

// likelihood-table values of the thread's 16 elementwise cells for time step t, issued early in two halves of 8
// (register budget): cells 0-7 after the arithmetic of the last convolution, cells 8-15 after its write-back, when
// the accumulators are dead; both are consumed after the following CTA barrier(s), which hide the L2 latency.
__device__ __forceinline__ const double *c2_lik_row(const PassArgs &a, const C2 &s, long long t) {
    return a.lik_table + t * (long long)a.pb.G + (size_t)s.r0 * a.pb.n1;
}
template <int K0>
__device__ __forceinline__ void c2_lik(const PassArgs &a, const C2 &s, long long t, double (&lk)[kC2Cells]) {
    if (a.pb.om_kind != BLG_OM_TABLE) return;
    const double *lt = c2_lik_row(a, s, t);
#pragma unroll
    for (int k = K0; k < K0 + 8; ++k) lk[k] = __ldg(lt + min((int)threadIdx.x + k * kC2Threads, s.cnt - 1));
}

// Elementwise sweep over the thread's cells g = tid + k * threads of the band.  Fast path (likelihood table, no
// reset in this step): branch-free, two phases per batch of 8 cells over register arrays -- `load(k, gi)` issues
// every read of cell gi (index clamped into the band), `apply(k, g, valid)` computes and stores -- so all loads of
// a batch are in flight before the first dependent instruction.  Otherwise cell by cell: `slow(g, lik)`.
template <typename Load, typename Apply, typename Slow>
__device__ __forceinline__ void c2_sweep(const PassArgs &a, const C2 &s, const LikTables &tb, long long t, bool fast,
                                         Load load, Apply apply, Slow slow) {
    const bool table = a.pb.om_kind == BLG_OM_TABLE;
    if (table && fast) {
        const int lastCell = s.cnt - 1;
        const int own = min((int)threadIdx.x, lastCell);  // cells beyond the band re-read the thread's OWN first cell
#pragma unroll
        for (int k0 = 0; k0 < kC2Cells; k0 += 8) {
#pragma unroll
            for (int k = k0; k < k0 + 8; ++k) {
                const int g = (int)threadIdx.x + k * kC2Threads;
                load(k, g <= lastCell ? g : own);
            }
#pragma unroll
            for (int k = k0; k < k0 + 8; ++k) {
                const int g = threadIdx.x + k * kC2Threads;
                apply(k, g, g <= lastCell);
            }
        }
    } else {
        const int n1 = a.pb.n1;
        const int dr = kC2Threads / n1, dc = kC2Threads - dr * n1;
        int r = threadIdx.x / n1, c = threadIdx.x - r * n1;
        const StepC *sc = a.steps + t * a.pb.ncols_eff;
        const double *lt = table ? a.lik_table + t * (long long)a.pb.G + (size_t)s.r0 * n1 : nullptr;
#pragma unroll 1
        for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) {
            slow(g, table ? __ldg(lt + g) : lik_cell(a.pb, tb, sc, s.r0 + r, c));
            r += dr;
            c += dc;
            if (c >= n1) {
                c -= n1;
                ++r;
            }
        }
    }
}

// row[g] = factor * band[g] for the thread's cells, streaming stores; batches of 8 shared-memory loads in flight
__device__ __forceinline__ void c2_flush(const C2 &s, uint32_t srcAddr, double *row, double factor) {
    const int lastCell = s.cnt - 1;
#pragma unroll
    for (int k0 = 0; k0 < kC2Cells; k0 += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = c2_lds(srcAddr + 8u * (uint32_t)min((int)threadIdx.x + (k0 + k) * kC2Threads, lastCell));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int g = threadIdx.x + (k0 + k) * kC2Threads;
            if (g <= lastCell) __stcs(row + g, v[k] * factor);
        }
    }
}

// band[g] *= factor for the thread's cells (same cell ownership as the sweeps: no barrier needed before)
__device__ __forceinline__ void c2_scale_inplace(const C2 &s, uint32_t addr, double factor) {
    const int lastCell = s.cnt - 1;
#pragma unroll
    for (int k0 = 0; k0 < kC2Cells; k0 += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = c2_lds(addr + 8u * (uint32_t)min((int)threadIdx.x + (k0 + k) * kC2Threads, lastCell));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int g = threadIdx.x + (k0 + k) * kC2Threads;
            if (g <= lastCell) c2_sts(addr + 8u * (uint32_t)g, v[k] * factor);
        }
    }
}

__device__ __forceinline__ void c2_init_barriers(const PassArgs &a, uint64_t *tma) {
    if (threadIdx.x == 0) {
        mbar_init(c2_halo_bar(a, 0), 1);
        mbar_init(c2_halo_bar(a, 1), 1);
        mbar_init(c2_sum_bar(a, 0), 1);
        mbar_init(c2_sum_bar(a, 1), 1);
        mbar_init(tma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
}

// ------------------------------------------------------------------------------------------------ K5 forward
// PROF: per-CTA cycle counters of the step phases (thread 0), written to a.trace[blockIdx.x * 8 + k]:
// 0 wait for the halo rows (+ normaliser and flush when rows are stored normalised), 1 axis-0 stage, 2 axis-1 stage,
// 3 staged likelihood + split barrier + normaliser, 4 elementwise sweep, 5 publish, 6 whole loop, 7 steps
template <int NT, bool PROF, int M0>
__global__ void __launch_bounds__(NT, 1) fwd_cluster2d_kernel(const PassArgs a) {
    static_assert(NT == kC2Threads, "layout constants assume kC2Threads");
    const DevProblem &pb = a.pb;
    const long long T = a.T;
    const int G = pb.G;
    LikTables tb;
    tb.A0 = pb.tabA[0];
    tb.A1 = pb.tabA[1];
    tb.A2 = pb.tabA[2];
    tb.B0 = pb.tabB[0];
    tb.B1 = pb.tabB[1];
    const long long slot = blockIdx.x / c2_cluster_size();
    const long long b = a.order ? a.order[slot] : slot;
    C2 s;
    const bool ok = c2_setup<false>(a, b, s);
    const bool lead = s.rank == 0 && threadIdx.x == 0;
    if (!ok) {  // uniform over the cluster
        if (lead) {
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
        }
        return;
    }
    const int n1 = a.pb.n1;
    const bool table = pb.om_kind == BLG_OM_TABLE;
    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    const bool raw = store && (a.flags & BLG_F_RAW_ALPHA);  // rows may stay unnormalised: bulk store out of the state
    uint64_t *barLik = reinterpret_cast<uint64_t *>(c2_misc(a) + kC2Mbar);
    uint32_t likPhase = 0;
    bool likInFlight = false;
    const uint32_t bandBytes = (uint32_t)(s.cnt * sizeof(double));
    const double *likBand = table ? a.lik_table + (size_t)s.r0 * n1 : nullptr;
    {   // every cell of the state buffer is finite from the start (slack rows are read, never used)
        const int total = a.c2_x_doubles;
        for (int e = threadIdx.x; e < total; e += kC2Threads) c2_X(a)[e] = 0.0;
        c2_init_barriers(a, barLik);
        __syncthreads();
        const double *init = a.prior + (size_t)s.r0 * n1;
        for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) c2_Xb(a)[g] = init[g];
        if (table) {
            if (threadIdx.x == 0) bulk_load(c2_S(a), likBand, bandBytes, barLik);
            likInFlight = true;
        }
    }
    c2_arrive();  // all CTAs of the cluster are resident and initialised before anybody pushes
    c2_wait();
    double *seq = store ? a.alpha_seq + b * a.seq_stride + (size_t)s.r0 * n1 : nullptr;
    const double *rb = a.reset_base;
    LogProduct lp;
    lp.init();
    bool dead = false, pendingStore = false;
    double kappa = 1.0;
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tStart = PROF ? clock64() : 0;
    const bool lateNorm = raw || !store;  // nothing needs the normaliser before the elementwise phase
    // the previous step's evidence increment: norm = sum over the cluster (core.py:385)
    auto takeNorm = [&](long long tPrev) -> bool {
        double part[1];
        c2_collect<1>(a, s, part);
        const double norm = part[0];
        if (!(norm > 0.0)) return false;  // core.py:388-400
        kappa = fast_rcp(norm);
        if (lead) {
            lp.mul(norm);                                              // core.py:403
            if (a.local) a.local[b * a.row_stride + tPrev] = norm * pb.lc_prod;   // core.py:404
        }
        return true;
    };

    for (long long t = 0; t < T; ++t) {
        const long long c0 = PROF ? clock64() : 0;
        if (t > 0) {
            c2_collect_halo(a, s);  // the neighbours' rows of the previous step's state
            if (!lateNorm) {
                if (!takeNorm(t - 1)) {
                    dead = true;
                    break;
                }
                // alpha[t-1] = kappa * X (core.py:389, :408)
                c2_flush(s, s.xbAddr, seq + (t - 1) * (long long)G, kappa);
            }
        }
        const bool trans = t > 0;
        const long long idx = t - 1;
        const bool post = trans && c2_in(s.loPost, s.hiPost, idx);
        const bool pre = trans && !post && c2_in(s.loPre, s.hiPre, idx);
        const bool act0 = trans && !post && s.R0 > 0 && c2_in(s.lo0, s.hi0, idx);
        const bool act1 = trans && !post && s.R1 > 0 && c2_in(s.lo1, s.hi1, idx);
        auto drainStore = [&]() {  // the bulk store of alpha[t-1] must have read the band before it is overwritten
            if (pendingStore && threadIdx.x == 0) bulk_wait_read<0>();
            pendingStore = false;
        };
        if (pre) {
            drainStore();
            __syncthreads();
            c2_reset_band(a, s, s.parPre);
        }
        const long long c1 = PROF ? clock64() : 0;
        long long cm = c1;
        // With a staged likelihood band and an active axis-1 stage the elementwise phase is FUSED into the write-back
        // of that convolution (alpha <- prior * likelihood, core.py:375-382): two passes over the band less.  The
        // normaliser of the previous step is collected right after the convolution arithmetic (hidden behind it).
        const bool fused = table && act1;
        bool normOk = true, normTaken = false;
        double kmul = (pre || post) ? 1.0 : kappa;
        double part[1] = {0.0};
        auto beforeEpilogue = [&]() {
            if (!fused) return;
            mbar_wait(barLik, likPhase);
            likPhase ^= 1u;
            likInFlight = false;
            if (t > 0 && lateNorm) {
                normOk = takeNorm(t - 1);
                normTaken = true;
                if (!pre) kmul = kappa;
            }
        };
        const bool wrote = c2_transition<PROF, M0>(a, s, act0, act1, cm, drainStore, beforeEpilogue, []() {},
                                               [&](uint32_t off, double v) {
                                                   if (!fused) return v;
                                                   const double y = v * kmul * c2_lds(s.sAddr + off);
                                                   part[0] += y;
                                                   return y;
                                               });
        if (!wrote && pendingStore) {
            drainStore();
            __syncthreads();
        }
        const long long c2 = PROF ? clock64() : 0;
        if (table && !fused) {
            mbar_wait(barLik, likPhase);
            likPhase ^= 1u;
            likInFlight = false;
        }
        c2_wait();  // every CTA has consumed its halo rows
        if (!normTaken && t > 0 && lateNorm) normOk = takeNorm(t - 1);
        if (!normOk) {  // core.py:388-400 (uniform over the cluster)
            dead = true;
            break;
        }
        const long long c3 = PROF ? clock64() : 0;
        if (!fused) {
            // E: alpha <- prior * likelihood (core.py:375-382), unnormalised
            if (!(pre || post)) kmul = kappa;
            double x[kC2Cells], l[kC2Cells];
            c2_sweep(
                a, s, tb, t, !post,
                [&](int k, int gi) {
                    x[k] = c2_lds(s.xbAddr + 8u * (uint32_t)gi);
                    l[k] = c2_lds(s.sAddr + 8u * (uint32_t)gi);
                },
                [&](int k, int g, bool valid) {
                    const double y = x[k] * kmul * l[k];
                    if (valid) c2_sts(s.xbAddr + 8u * (uint32_t)g, y);
                    part[0] += valid ? y : 0.0;
                },
                [&](int g, double lik) {
                    const double v = post ? __ldg(rb + (size_t)s.r0 * n1 + g) * s.parPost : c2_Xb(a)[g] * kmul;
                    const double y = v * lik;
                    c2_Xb(a)[g] = y;
                    part[0] += y;
                });
        }
        if (raw) fence_proxy_async();  // the band is read by the bulk-async store below
        const long long c4 = PROF ? clock64() : 0;
        const bool more = table && t + 1 < T;
        c2_publish<1>(a, s, part, [&]() {
            if (threadIdx.x == 0) {
                if (more) bulk_load(c2_S(a), likBand + (t + 1) * (long long)G, bandBytes, barLik);
                if (raw) bulk_store(seq + t * (long long)G, c2_Xb(a), bandBytes);
            }
        });
        if (more) likInFlight = true;
        pendingStore = raw;
        if (PROF) {
            const long long c5 = clock64();
            tk[0] += c1 - c0;
            tk[1] += cm - c1;
            tk[2] += c2 - cm;
            tk[3] += c3 - c2;
            tk[4] += c4 - c3;
            tk[5] += c5 - c4;
            tk[7] += 1;
        }
    }
    if (!dead) {  // the last step's evidence increment and row
        c2_collect_halo(a, s);
        if (!takeNorm(T - 1)) {
            dead = true;
        } else if (store && !raw) {
            c2_flush(s, s.xbAddr, seq + (T - 1) * (long long)G, kappa);
        }
    }
    if (likInFlight) mbar_wait(barLik, likPhase);  // a combo that died leaves with its last staging copy landed
    if (threadIdx.x == 0) bulk_wait_all();
    if (PROF && a.trace && threadIdx.x == 0) {
        tk[6] = clock64() - tStart;
        for (int k = 0; k < 8; ++k) a.trace[(long long)blockIdx.x * 8 + k] = tk[k];
    }
    if (lead) {
        double logE = lp.log_value();
        if (dead)
            logE = -INFINITY;
        else
            logE += log(pb.lc_prod);  // core.py:417
        a.logE[b] = logE;
        if (a.alive) a.alive[b] = dead ? 0 : 1;
    }
    c2_arrive_relaxed();  // nobody leaves while a peer may still have stores in flight to it
    c2_wait();
}

// ------------------------------------------------------------------------------------------------ K6 backward
// State X = beta_i * lik_i (unnormalised by one step); S = u_i = alpha_i * beta_i, the unnormalised smoothed
// posterior of step i, which leaves for HBM (divided by its cluster-wide sum) at the beginning of the next
// iteration and is then refilled with alpha[i-1] by one bulk-async (TMA) copy that lands during the convolutions.
// PROF counters (a.trace[blockIdx.x * 8 + k]): 0 collect (halo rows + sums), 1 posterior flush + alpha TMA issue,
// 2 axis-0 stage, 3 axis-1 stage + likelihood loads, 4 wait alpha + split barrier, 5 sweep, 6 publish, 7 steps
template <int NT, bool PROF, int M0>
__global__ void __launch_bounds__(NT, 1) bwd_cluster2d_kernel(const PassArgs a) {
    static_assert(NT == kC2Threads, "layout constants assume kC2Threads");
    const DevProblem &pb = a.pb;
    const long long T = a.T;
    const int G = pb.G;
    LikTables tb;
    tb.A0 = pb.tabA[0];
    tb.A1 = pb.tabA[1];
    tb.A2 = pb.tabA[2];
    tb.B0 = pb.tabB[0];
    tb.B1 = pb.tabB[1];
    const long long slot = blockIdx.x / c2_cluster_size();
    const long long b = a.order ? a.order[slot] : slot;
    if (a.alive && a.alive[b] != 1) return;  // the forward pass aborted (core.py:400); uniform over the cluster
    C2 s;
    if (!c2_setup<true>(a, b, s)) return;
    const bool lead = s.rank == 0 && threadIdx.x == 0;
    const int n1 = a.pb.n1;
    uint64_t *bar = reinterpret_cast<uint64_t *>(c2_misc(a) + kC2Mbar);
    uint32_t phase = 0;
    double *seq = a.alpha_seq + b * a.seq_stride + (size_t)s.r0 * n1;
    // filtering rows (out-of-place smoothing when alpha_src is given)
    const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride + (size_t)s.r0 * n1 : seq;
    const uint32_t bandBytes = (uint32_t)(s.cnt * sizeof(double));
    {
        const int total = a.c2_x_doubles;
        for (int e = threadIdx.x; e < total; e += kC2Threads) c2_X(a)[e] = 0.0;
        c2_init_barriers(a, bar);
        __syncthreads();
        if (threadIdx.x == 0) bulk_load(c2_S(a), src + (T - 1) * (long long)G, bandBytes, bar);
    }
    c2_arrive();
    c2_wait();
    const double *rb = a.reset_base;
    const double beta0 = 1.0 / (double)G;  // core.py:424-425
    bool dead = false;
    double kb = 1.0, inv = 1.0;
    double sums[3];
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    // elementwise phase for row j: betaNew = beta0 | reset | X * scale;  u = alpha_j * betaNew -> S;  X = betaNew * lik_j
    // raw mode (BLG_F_RAW_POSTERIOR): the unnormalised row u_i leaves by a bulk store right after the sweep that made
    // it and its factor 1/sum(u_i) goes to row_scale; the sums are then only needed by the NEXT sweep, so their
    // all-to-all hides behind the convolutions (like the evidence increment of the forward pass)
    const bool raw = a.row_scale != nullptr;
    // sums of row `row`: normalisers, local evidence, row scale; false = zero norm (core.py:440-452)
    auto takeSums = [&](long long row) -> bool {
        c2_collect<3>(a, s, sums);
        const double sab = sums[0], sbb = sums[1], q = sums[2];
        if (!(sab > 0.0) || !(sbb > 0.0)) return false;
        inv = fast_rcp(sab);  // posterior = alpha*beta / sum(alpha*beta)  core.py:436-441
        kb = fast_rcp(sbb);   // core.py:470 (beta only enters scale-free expressions)
        if (lead) {
            if (a.local) a.local[b * a.row_stride + row] = fast_div(1.0, q * inv * pb.lc_prod);  // core.py:463
            if (raw) a.row_scale[b * a.row_stride + row] = inv;
        }
        return true;
    };
    // elementwise phase for row j; lateRow >= 0: first collect the sums of that (previous) row
    auto elementwise = [&](int mode /*0: beta0, 1: X*scale, 2: post reset*/, bool unitScale, long long j,
                           double (&lk)[kC2Cells], long long lateRow) -> bool {
        const long long e0 = PROF ? clock64() : 0;
        mbar_wait(bar, phase);  // alpha[j] band has landed in S
        phase ^= 1u;
        c2_wait();  // halo rows consumed everywhere
        if (lateRow >= 0 && !takeSums(lateRow)) return false;
        const double scale = unitScale ? 1.0 : kb;
        const long long e1 = PROF ? clock64() : 0;
        sums[0] = sums[1] = sums[2] = 0.0;
        double x[kC2Cells], al[kC2Cells];
        c2_sweep(
            a, s, tb, j, mode == 1,
            [&](int k, int gi) {
                x[k] = c2_lds(s.xbAddr + 8u * (uint32_t)gi);
                al[k] = c2_lds(s.sAddr + 8u * (uint32_t)gi);
            },
            [&](int k, int g, bool valid) {
                const double bn = x[k] * scale;
                const double u = al[k] * bn;  // core.py:436 (unnormalised)
                const double lik = lk[k];
                if (valid) {
                    c2_sts(s.sAddr + 8u * (uint32_t)g, u);
                    c2_sts(s.xbAddr + 8u * (uint32_t)g, bn * lik);  // core.py:467
                }
                sums[0] += valid ? u : 0.0;
                sums[1] += valid ? bn : 0.0;
                // core.py:463 post/lik without the IEEE division subroutine: the fast reciprocal flushes subnormal
                // operands, so tiny likelihoods are rescaled by 2^600 first (exact); lik == 0 gives NaN like 0/0 does
                // (u = 0 there: alpha carries the same likelihood factor)
                const bool small = lik < 1e-290;
                const double qv = u * fast_rcp(small ? lik * 0x1p600 : lik);
                sums[2] += valid ? (small ? qv * 0x1p600 : qv) : 0.0;
            },
            [&](int g, double lik) {
                double bn;
                if (mode == 0)
                    bn = beta0;
                else if (mode == 2)
                    bn = __ldg(rb + (size_t)s.r0 * n1 + g) * s.parPost;
                else
                    bn = c2_Xb(a)[g] * scale;
                const double u = c2_S(a)[g] * bn;
                c2_S(a)[g] = u;
                sums[0] += u;
                sums[1] += bn;
                sums[2] += fast_div(u, lik);
                c2_Xb(a)[g] = bn * lik;
            });
        if (raw) fence_proxy_async();  // S is read by the bulk-async store below
        const long long e2 = PROF ? clock64() : 0;
        c2_publish<3>(a, s, sums, [&]() {
            if (raw && threadIdx.x == 0) bulk_store(seq + j * (long long)G, c2_S(a), bandBytes);
        });
        if (PROF) {
            const long long e3 = clock64();
            tk[4] += e1 - e0;
            tk[5] += e2 - e1;
            tk[6] += e3 - e2;
        }
        return true;
    };

    {
        double lk[kC2Cells];
        c2_lik<0>(a, s, T - 1, lk);
        c2_lik<8>(a, s, T - 1, lk);
        c2_arrive_relaxed();  // pairs with the wait inside elementwise()
        elementwise(0, true, T - 1, lk, -1);
    }
    for (long long i = T - 1; i >= 0; --i) {
        const long long c0 = PROF ? clock64() : 0;
        c2_collect_halo(a, s);
        if (!raw) {
            if (!takeSums(i)) {
                dead = true;
                break;
            }
            // smoothed posterior of step i (core.py:441): normalised in place, then one bulk-async store of the band
            c2_scale_inplace(s, s.sAddr, inv);
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) bulk_store(seq + i * (long long)G, c2_S(a), bandBytes);
        }
        const long long c1 = PROF ? clock64() : 0;
        if (i == 0) {
            if (raw && !takeSums(0)) dead = true;
            break;
        }
        // alpha[i-1] is fetched into the same buffer as soon as the store has read it: issued by thread 0 after its
        // share of the axis-0 convolution (no waiting), or right away when that stage is idle
        bool alphaPending = true;
        auto issueAlpha = [&]() {
            if (alphaPending && threadIdx.x == 0) {
                bulk_wait_read<0>();
                bulk_load(c2_S(a), src + (i - 1) * (long long)G, bandBytes, bar);
                if (pb.om_kind == BLG_OM_TABLE && i >= 2)
                    c2_prefetch_l2(a.lik_table + (i - 2) * (long long)G + (size_t)s.r0 * n1, bandBytes);
            }
            alphaPending = false;
        };
        const bool post = c2_in(s.loPost, s.hiPost, i);
        const bool pre = !post && c2_in(s.loPre, s.hiPre, i);
        const bool act0 = !post && s.R0 > 0 && c2_in(s.lo0, s.hi0, i);
        const bool act1 = !post && s.R1 > 0 && c2_in(s.lo1, s.hi1, i);
        if (pre) {
            __syncthreads();  // the publish phase of the previous sweep may still be reading the band rows it pushes
            c2_reset_band(a, s, s.parPre);
        }
        if (!act0) issueAlpha();
        double lk[kC2Cells];
        const long long c2 = PROF ? clock64() : 0;
        long long cm = c2;
        c2_transition<PROF, M0>(a, s, act0, act1, cm, issueAlpha, [&]() { c2_lik<0>(a, s, i - 1, lk); },
                            [&]() { c2_lik<8>(a, s, i - 1, lk); });
        if (PROF) {
            const long long c3 = clock64();
            tk[0] += c1 - c0;
            tk[1] += c2 - c1;
            tk[2] += cm - c2;
            tk[3] += c3 - cm;
            tk[7] += 1;
        }
        if (!elementwise(post ? 2 : 1, pre, i - 1, lk, raw ? i : -1)) {
            dead = true;
            break;
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();
    if (PROF && a.trace && threadIdx.x == 0)
        for (int k = 0; k < 8; ++k) a.trace[(long long)blockIdx.x * 8 + k] = tk[k];
    if (dead && lead) {
        a.logE[b] = -INFINITY;
        if (a.alive) a.alive[b] = -1;
    }
    c2_arrive_relaxed();  // nobody leaves while a peer may still have stores in flight to it
    c2_wait();
}

}  // namespace blg
