// cluster2d.cuh -- K5/K6: forward / backward passes for 2-D grids that do not fit in ONE SM's shared memory
// (200^2, 256^2: BASELINE.json configs[2], [3]) with the state resident ON CHIP across a thread-block cluster.
//
// One cluster of C CTAs (C = 2, 4 or 8, one CTA per SM) owns one hyper-parameter combination for all T steps.  The
// grid is cut into C bands of rows (axis 0); every CTA keeps its band of the (unnormalised) state in shared memory
// for the whole recursion, so the only HBM traffic is the compulsory one: the alpha[t] row store (forward), the
// alpha[t] row load + posterior store (backward) and the likelihood row shared by all combos.  Per time step:
//
//   F   flush   the row of the previous step leaves for HBM, coalesced, with the lazy normaliser applied
//   0   GRW on axis 0 (across bands): every thread convolves M0 rows of one column out of shared memory into
//       registers; the R0 rows it needs from the neighbouring bands were PUSHED into this CTA's halo rows through
//       distributed shared memory (st.shared::cluster) by the neighbours at the end of the previous step
//   1   GRW on axis 1 (inside a row): M1 cells of one row per thread, reflect boundary by index
//   E   elementwise: prior x likelihood (forward) / alpha x beta, beta x likelihood (backward), partial sums, halo
//       pushes for the next step
//   A   ONE full cluster barrier per step (barrier.cluster arrive.release / wait.acquire): it publishes the halo rows
//       and the per-CTA partial sums (evidence increment; sum(alpha beta), sum(beta), sum(post/lik)); a second,
//       split barrier (arrive after the axis-0 reads, wait before the pushes) protects the halo rows and is hidden
//       behind the axis-1 convolution.
//
// Both convolutions commute (separable, linear), so axis 0 always runs first whatever the program order; the results
// agree with the reference order to rounding (1e-16 relative).  Semantics: core.py:372-417, :434-470,
// transitionModels.py:96-115 (GaussianRandomWalk), :300-312 (ChangePoint reset before / after the random walks).
#pragma once

#include "common.cuh"
#include "stream2d.cuh"  // classify2d, in_window

namespace blg {

constexpr int kC2Threads = 512;
constexpr int kC2M0 = 16;      // rows per work item of the axis-0 convolution
constexpr int kC2M1 = 17;      // cells per work item of the axis-1 convolution (odd: conflict-free 64-bit LDS)
constexpr int kC2Cells = 16;   // cells per thread of the elementwise phases (band <= 16 * 512 cells)
constexpr int kC2MaxCluster = 8;
// misc region (doubles): [0,128) RedScratch of build_weights, [128,176) per-warp partials [3][16],
// [176,224) cluster slots [2][3][8], [224] mbarrier
constexpr int kC2WarpPart = 128, kC2Slots = 176, kC2Mbar = 224;

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
__device__ __forceinline__ uint32_t c2_map(const void *p, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void c2_st_remote(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void c2_prefetch_l2(const void *g, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

// M outputs of one COLUMN: col points at row (i0 - R) of the column, rows are `pitch` doubles apart.  Exact tap
// count (no zero-padded taps), so nothing beyond row i0 + M - 1 + R is read.
template <int M>
__device__ __forceinline__ void c2_conv_col(const double *__restrict__ col, int pitch, int R, const double *__restrict__ W,
                                            double (&acc)[M]) {
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = col[(size_t)m * pitch];
        acc[m] = 0.0;
    }
    const int taps = 2 * R + 1;
    const double *p = col + (size_t)M * pitch;
    for (int j0 = 0; j0 < taps; j0 += M) {
#pragma unroll
        for (int u = 0; u < M; ++u) {
            if (j0 + u < taps) {
                const double w = W[j0 + u];
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
                if (j0 + u + 1 < taps) win[u] = p[(size_t)u * pitch];
            }
        }
        p += (size_t)M * pitch;
    }
}

// M outputs of one ROW starting at cell i0; NI_EXTEND_REFLECT by index (valid while R + M <= n).
template <int M>
__device__ __forceinline__ void c2_conv_row(const double *__restrict__ row, int i0, int n, int R, const double *__restrict__ W,
                                            double (&acc)[M]) {
    double win[M];
    int e = i0 - R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = row[reflect_once(e + m, n)];
        acc[m] = 0.0;
    }
    e += M;
    const int taps = 2 * R + 1;
    for (int j0 = 0; j0 < taps; j0 += M) {
#pragma unroll
        for (int u = 0; u < M; ++u) {
            if (j0 + u < taps) {
                const double w = W[j0 + u];
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = fma(w, win[(u + m) % M], acc[m]);
                if (j0 + u + 1 < taps) win[u] = row[reflect_once(e + u, n)];
            }
        }
        e += M;
    }
}

struct C2 {
    double *X, *Xb, *S, *W0, *W1, *misc;
    double sig0, sig1, parPre, parPost;
    int R0, R1;
    int w0[4], w1[4], wPre[4], wPost[4];
    Stream2dOps ops;
    int rank, C, r0, nb, nbmax, H0, n0, n1, cnt;
    int slotParity;
};

// Geometry, per-combo parameters, weights.  Returns false (uniformly over the cluster) if a radius exceeds what the
// host sized the layout for.
__device__ __forceinline__ bool c2_setup(const PassArgs &a, double *sm, long long b, C2 &s) {
    s.n0 = a.pb.n0;
    s.n1 = a.pb.n1;
    s.rank = (int)c2_cluster_rank();
    s.C = (int)c2_cluster_size();
    s.nbmax = a.c2_nb;
    s.H0 = a.c2_h0;
    s.r0 = s.rank * s.nbmax;
    s.nb = min(s.nbmax, s.n0 - s.r0);
    s.cnt = s.nb * s.n1;
    s.X = sm + a.c2_off_x;
    s.Xb = s.X + (size_t)s.H0 * s.n1;
    s.S = sm + a.c2_off_s;
    s.misc = sm + a.off_misc;
    s.slotParity = 0;
    s.ops = classify2d(a.pg.n_ops, a.pg.kind, a.pg.axis);
    const int K = a.pg.n_ops;
    auto load = [&](int k, double &par, int &rad, int *w) {
        par = 0.0;
        rad = 0;
        for (int q = 0; q < 4; ++q) w[q] = 0;
        if (k < 0) return;
        par = a.pg.param[b * K + k];
        rad = a.pg.radius[b * K + k];
        for (int q = 0; q < 4; ++q) w[q] = a.pg.window[(b * K + k) * 4 + q];
    };
    int dummy;
    load(s.ops.k0, s.sig0, s.R0, s.w0);
    load(s.ops.k1, s.sig1, s.R1, s.w1);
    load(s.ops.pre, s.parPre, dummy, s.wPre);
    load(s.ops.post, s.parPost, dummy, s.wPost);
    if (!(s.sig0 > 0.0) || s.R0 <= 0) s.R0 = 0;  // transitionModels.py:110-113: identity
    if (!(s.sig1 > 0.0) || s.R1 <= 0) s.R1 = 0;
    s.W0 = sm + a.off_w + (s.ops.k0 >= 0 ? a.pg.w_off[s.ops.k0] : 0);
    s.W1 = sm + a.off_w + (s.ops.k1 >= 0 ? a.pg.w_off[s.ops.k1] : 0);
    bool ok = true;
    if (s.R0 > s.H0 || s.R0 > s.n0 - (s.C - 1) * s.nbmax) ok = false;
    if (s.R1 + kC2M1 > s.n1) ok = false;
    RedScratch rs;
    rs.buf = s.misc;
    rs.phase = 0;
    if (s.ops.k0 >= 0 && s.R0 > 0) {
        if (2 * s.R0 + 1 > a.pg.w_len[s.ops.k0])
            ok = false;
        else
            build_weights(s.W0, 2 * s.R0 + 1, s.sig0, s.R0, rs);
    }
    if (s.ops.k1 >= 0 && s.R1 > 0) {
        if (2 * s.R1 + 1 > a.pg.w_len[s.ops.k1])
            ok = false;
        else
            build_weights(s.W1, 2 * s.R1 + 1, s.sig1, s.R1, rs);
    }
    return ok;
}

// Cluster-wide sums of K values: warp shuffles, per-warp partials, every CTA pushes its total into slot[rank] of
// every CTA (st.shared::cluster), ONE cluster barrier (which also publishes the halo rows pushed before), and all
// CTAs add the C slots in the same order -> bit-identical results in every CTA.
template <int K>
__device__ __forceinline__ void c2_reduce(C2 &s, double (&v)[K]) {
    double *wpart = s.misc + kC2WarpPart, *slots = s.misc + kC2Slots + s.slotParity * 3 * kC2MaxCluster;
    s.slotParity ^= 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double t = warp_sum(v[k]);
        if (lane == 0) wpart[k * 16 + warp] = t;
    }
    __syncthreads();
    if ((int)threadIdx.x < s.C) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double tot = 0.0;
            for (int w = 0; w < kC2Threads / 32; ++w) tot += wpart[k * 16 + w];
            c2_st_remote(c2_map(&slots[k * kC2MaxCluster + s.rank], (unsigned)threadIdx.x), tot);
        }
    }
    c2_arrive();
    c2_wait();
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
    for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) s.Xb[g] = __ldg(rb + (size_t)s.r0 * s.n1 + g) * par;
    const int per = s.R0 * s.n1;
    for (int e = threadIdx.x; e < 2 * per; e += kC2Threads) {
        const int side = e >= per ? 1 : 0, q = e - side * per, k = q / s.n1, c = q - k * s.n1;
        const int grow = reflect_any(side ? s.r0 + s.nb + k : s.r0 - 1 - k, s.n0);
        const int lrow = side ? s.H0 + s.nb + k : s.H0 - 1 - k;
        s.X[(size_t)lrow * s.n1 + c] = __ldg(rb + (size_t)grow * s.n1 + c) * par;
    }
    __syncthreads();
}

// The two convolutions of one step, in place.  Returns with all threads synchronised on the new band; the split
// cluster barrier "halo rows consumed" has been ARRIVED at (the caller waits before it pushes).
template <typename AfterConv>
__device__ __forceinline__ void c2_transition(const C2 &s, bool act0, bool act1, long long &tmid, AfterConv afterConv) {
    const int n1 = s.n1;
    if (act0) {
        const int S0 = (s.nbmax + kC2M0 - 1) / kC2M0;
        const int w = threadIdx.x;
        const bool has = w < n1 * S0;
        const int seg = w / n1, c = w - seg * n1, i0 = seg * kC2M0;
        double acc[kC2M0];
        if (has) c2_conv_col<kC2M0>(s.X + (size_t)(s.H0 + i0 - s.R0) * n1 + c, n1, s.R0, s.W0, acc);
        __syncthreads();
        c2_arrive_relaxed();  // halo rows consumed
        if (has) {
#pragma unroll
            for (int m = 0; m < kC2M0; ++m)
                if (i0 + m < s.nb) s.Xb[(size_t)(i0 + m) * n1 + c] = acc[m];
        }
        __syncthreads();
    } else {
        c2_arrive_relaxed();
    }
    tmid = clock64();
    if (act1) {
        const int S1 = (n1 + kC2M1 - 1) / kC2M1;
        const int w = threadIdx.x;
        const int l = w / S1, sg = w - l * S1, i0 = sg * kC2M1;
        const bool has = l < s.nb;
        double acc[kC2M1];
        if (has) c2_conv_row<kC2M1>(s.Xb + (size_t)l * n1, i0, n1, s.R1, s.W1, acc);
        afterConv();
        __syncthreads();
        if (has) {
#pragma unroll
            for (int m = 0; m < kC2M1; ++m)
                if (i0 + m < n1) s.Xb[(size_t)l * n1 + i0 + m] = acc[m];
        }
        __syncthreads();
    } else {
        afterConv();
    }
}

// Push cell (r, c) of the band into the halo rows that mirror it: the neighbouring CTAs' (remote) or, at the grid
// boundary, this CTA's own reflected rows.
__device__ __forceinline__ void c2_push(const C2 &s, int r, int c, double y, uint32_t upBase, uint32_t dnBase) {
    if (r < s.R0) {
        if (s.rank > 0)
            c2_st_remote(upBase + (uint32_t)(((s.H0 + s.nbmax + r) * s.n1 + c) * 8), y);  // their row nb + r
        else
            s.X[(size_t)(s.H0 - 1 - r) * s.n1 + c] = y;  // reflect: row -1-r = row r
    }
    if (r >= s.nb - s.R0) {
        if (s.rank < s.C - 1)
            c2_st_remote(dnBase + (uint32_t)(((s.H0 - (s.nb - r)) * s.n1 + c) * 8), y);  // their row -(nb - r)
        else
            s.X[(size_t)(s.H0 + s.nb + (s.nb - 1 - r)) * s.n1 + c] = y;  // reflect: row nb + k = row nb-1-k
    }
}

// likelihood-table values of the thread's elementwise cells for time step t (issued early: consumed after two CTA
// barriers, which hide most of the L2 latency).  Without a table the likelihood is evaluated inside the elementwise loop.
__device__ __forceinline__ void c2_lik(const PassArgs &a, const C2 &s, long long t, double (&lk)[kC2Cells]) {
    if (a.pb.om_kind != BLG_OM_TABLE) return;
    const double *lt = a.lik_table + t * (long long)a.pb.G + (size_t)s.r0 * s.n1;
#pragma unroll
    for (int k = 0; k < kC2Cells; ++k) {
        const int g = threadIdx.x + k * kC2Threads;
        lk[k] = g < s.cnt ? __ldg(lt + g) : 0.0;
    }
}

// Elementwise sweep over the thread's cells g = tid + k * threads of the band: cell(g, r, c, lik).
template <typename Cell>
__device__ __forceinline__ void c2_sweep(const PassArgs &a, const C2 &s, const LikTables &tb, long long t,
                                         const double (&lk)[kC2Cells], Cell cell) {
    const int n1 = s.n1;
    const int dr = kC2Threads / n1, dc = kC2Threads - dr * n1;
    int r = threadIdx.x / n1, c = threadIdx.x - r * n1;
    if (a.pb.om_kind == BLG_OM_TABLE) {
#pragma unroll
        for (int k = 0; k < kC2Cells; ++k) {
            const int g = threadIdx.x + k * kC2Threads;
            if (g < s.cnt) cell(g, r, c, lk[k]);
            r += dr;
            c += dc;
            if (c >= n1) {
                c -= n1;
                ++r;
            }
        }
    } else {
        const StepC *sc = a.steps + t * a.pb.ncols_eff;
#pragma unroll 1
        for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) {
            cell(g, r, c, lik_cell(a.pb, tb, sc, s.r0 + r, c));
            r += dr;
            c += dc;
            if (c >= n1) {
                c -= n1;
                ++r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K5 forward
// PROF: per-CTA cycle counters of the step phases (thread 0), written to a.trace[blockIdx.x * 8 + k]:
// 0 flush, 1 axis-0 stage, 2 axis-1 stage, 3 wait for the split barrier, 4 elementwise sweep + pushes, 5 reduction +
// cluster barrier, 6 whole loop, 7 steps
template <int NT, bool PROF>
__global__ void __launch_bounds__(NT, 1) fwd_cluster2d_kernel(const PassArgs a) {
    static_assert(NT == kC2Threads, "layout constants assume kC2Threads");
    extern __shared__ __align__(16) double sm[];
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
    const bool ok = c2_setup(a, sm, b, s);
    const bool lead = s.rank == 0 && threadIdx.x == 0;
    if (!ok) {  // uniform over the cluster
        if (lead) {
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
        }
        return;
    }
    const int n1 = s.n1;
    {   // every cell of the state buffer is finite from the start (slack rows are read, never used)
        const int total = a.c2_x_doubles;
        for (int e = threadIdx.x; e < total; e += kC2Threads) s.X[e] = 0.0;
        __syncthreads();
        const double *init = a.prior + (size_t)s.r0 * n1;
        for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) s.Xb[g] = init[g];
    }
    c2_arrive();  // all CTAs of the cluster are resident and initialised before anybody pushes
    c2_wait();
    const uint32_t upBase = s.rank > 0 ? c2_map(s.X, (unsigned)(s.rank - 1)) : 0u;
    const uint32_t dnBase = s.rank < s.C - 1 ? c2_map(s.X, (unsigned)(s.rank + 1)) : 0u;
    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    double *seq = store ? a.alpha_seq + b * T * (long long)G + (size_t)s.r0 * n1 : nullptr;
    const double *rb = a.reset_base;
    const uint32_t bandBytes = (uint32_t)(s.cnt * sizeof(double));
    LogProduct lp;
    lp.init();
    bool dead = false;
    double kappa = 1.0;
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tStart = clock64();

    for (long long t = 0; t < T; ++t) {
        const long long c0 = clock64();
        const bool trans = t > 0;
        const long long idx = t - 1;
        const bool post = trans && s.ops.post >= 0 && in_window(s.wPost, idx, false);
        const bool pre = trans && !post && s.ops.pre >= 0 && in_window(s.wPre, idx, false);
        const bool act0 = trans && !post && s.R0 > 0 && in_window(s.w0, idx, false);
        const bool act1 = trans && !post && s.R1 > 0 && in_window(s.w1, idx, false);
        // F: alpha[t-1] = kappa * X (core.py:389, :408)
        if (store && t > 0) {
            double *row = seq + (t - 1) * (long long)G;
            for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) __stcs(row + g, s.Xb[g] * kappa);
        }
        if (threadIdx.x == 0 && pb.om_kind == BLG_OM_TABLE && t + 1 < T)  // next likelihood band HBM -> L2
            c2_prefetch_l2(a.lik_table + (t + 1) * (long long)G + (size_t)s.r0 * n1, bandBytes);
        double keff = kappa;
        if (pre) {
            c2_reset_band(a, s, s.parPre);
            keff = 1.0;
        }
        double lk[kC2Cells];
        const long long c1 = clock64();
        long long cm = c1;
        c2_transition(s, act0, act1, cm, [&]() { c2_lik(a, s, t, lk); });
        const long long c2 = clock64();
        // E: alpha <- prior * likelihood (core.py:375-382), unnormalised; halo pushes for the next step
        c2_wait();  // every CTA has consumed its halo rows
        const long long c3 = clock64();
        double part[1] = {0.0};
        c2_sweep(a, s, tb, t, lk, [&](int g, int r, int c, double lik) {
            const double v = post ? __ldg(rb + (size_t)s.r0 * n1 + g) * s.parPost : s.Xb[g] * keff;
            const double y = v * lik;
            s.Xb[g] = y;
            part[0] += y;
            if (s.R0 > 0) c2_push(s, r, c, y, upBase, dnBase);
        });
        const long long c4 = clock64();
        c2_reduce<1>(s, part);  // core.py:385
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
        const double norm = part[0];
        if (!(norm > 0.0)) {  // core.py:388-400
            dead = true;
            break;
        }
        kappa = fast_rcp(norm);
        if (lead) {
            lp.mul(norm);                                         // core.py:403
            if (a.local) a.local[b * T + t] = norm * pb.lc_prod;  // core.py:404
        }
    }
    if (!dead && store) {
        double *row = seq + (T - 1) * (long long)G;
        for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) __stcs(row + g, s.Xb[g] * kappa);
    }
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
}

// ------------------------------------------------------------------------------------------------ K6 backward
// State X = beta_i * lik_i (unnormalised by one step); S = u_i = alpha_i * beta_i, the unnormalised smoothed
// posterior of step i, which leaves for HBM (divided by its cluster-wide sum) at the beginning of the next
// iteration and is then refilled with alpha[i-1] by one bulk-async (TMA) copy that lands during the convolutions.
template <int NT>
__global__ void __launch_bounds__(NT, 1) bwd_cluster2d_kernel(const PassArgs a) {
    static_assert(NT == kC2Threads, "layout constants assume kC2Threads");
    extern __shared__ __align__(16) double sm[];
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
    if (!c2_setup(a, sm, b, s)) return;
    const bool lead = s.rank == 0 && threadIdx.x == 0;
    const int n1 = s.n1;
    uint64_t *bar = reinterpret_cast<uint64_t *>(s.misc + kC2Mbar);
    uint32_t phase = 0;
    double *seq = a.alpha_seq + b * T * (long long)G + (size_t)s.r0 * n1;
    const uint32_t bandBytes = (uint32_t)(s.cnt * sizeof(double));
    {
        const int total = a.c2_x_doubles;
        for (int e = threadIdx.x; e < total; e += kC2Threads) s.X[e] = 0.0;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (threadIdx.x == 0) bulk_load(s.S, seq + (T - 1) * (long long)G, bandBytes, bar);
    }
    c2_arrive();
    c2_wait();
    const uint32_t upBase = s.rank > 0 ? c2_map(s.X, (unsigned)(s.rank - 1)) : 0u;
    const uint32_t dnBase = s.rank < s.C - 1 ? c2_map(s.X, (unsigned)(s.rank + 1)) : 0u;
    const double *rb = a.reset_base;
    const double beta0 = 1.0 / (double)G;  // core.py:424-425
    bool dead = false;
    double kb = 1.0;
    double sums[3];

    // elementwise phase for row j: betaNew = beta0 | reset | X * scale;  u = alpha_j * betaNew -> S;  X = betaNew * lik_j
    auto elementwise = [&](int mode /*0: beta0, 1: X*scale, 2: post reset*/, double scale, long long j,
                           const double (&lk)[kC2Cells]) {
        mbar_wait(bar, phase);  // alpha[j] band has landed in S
        phase ^= 1u;
        c2_wait();  // halo rows consumed everywhere
        sums[0] = sums[1] = sums[2] = 0.0;
        c2_sweep(a, s, tb, j, lk, [&](int g, int r, int c, double lik) {
            double bn;
            if (mode == 0)
                bn = beta0;
            else if (mode == 2)
                bn = __ldg(rb + (size_t)s.r0 * n1 + g) * s.parPost;
            else
                bn = s.Xb[g] * scale;
            const double u = s.S[g] * bn;  // core.py:436 (unnormalised)
            s.S[g] = u;
            sums[0] += u;
            sums[1] += bn;
            sums[2] += fast_div(u, lik);  // core.py:463
            const double y = bn * lik;    // core.py:467
            s.Xb[g] = y;
            if (s.R0 > 0) c2_push(s, r, c, y, upBase, dnBase);
        });
        c2_reduce<3>(s, sums);
    };

    {
        double lk[kC2Cells];
        c2_lik(a, s, T - 1, lk);
        c2_arrive_relaxed();  // pairs with the wait inside elementwise()
        elementwise(0, 1.0, T - 1, lk);
    }
    for (long long i = T - 1; i >= 0; --i) {
        const double sab = sums[0], sbb = sums[1], q = sums[2];
        if (!(sab > 0.0) || !(sbb > 0.0)) {  // core.py:440-452
            dead = true;
            break;
        }
        const double inv = fast_rcp(sab);
        kb = fast_rcp(sbb);  // core.py:470 (beta only enters scale-free expressions)
        if (lead && a.local) a.local[b * T + i] = fast_div(1.0, q * inv * pb.lc_prod);  // core.py:463
        // F: smoothed posterior of step i (core.py:441)
        {
            double *row = seq + i * (long long)G;
            for (int g = threadIdx.x; g < s.cnt; g += kC2Threads) __stcs(row + g, s.S[g] * inv);
        }
        if (i == 0) break;
        __syncthreads();  // S is free
        if (threadIdx.x == 0) {
            fence_proxy_async();
            bulk_load(s.S, seq + (i - 1) * (long long)G, bandBytes, bar);
            if (pb.om_kind == BLG_OM_TABLE && i >= 2)
                c2_prefetch_l2(a.lik_table + (i - 2) * (long long)G + (size_t)s.r0 * n1, bandBytes);
        }
        const bool post = s.ops.post >= 0 && in_window(s.wPost, i, true);
        const bool pre = !post && s.ops.pre >= 0 && in_window(s.wPre, i, true);
        const bool act0 = !post && s.R0 > 0 && in_window(s.w0, i, true);
        const bool act1 = !post && s.R1 > 0 && in_window(s.w1, i, true);
        double scale = kb;
        if (pre) {
            c2_reset_band(a, s, s.parPre);
            scale = 1.0;
        }
        double lk[kC2Cells];
        long long cm;
        c2_transition(s, act0, act1, cm, [&]() { c2_lik(a, s, i - 1, lk); });
        elementwise(post ? 2 : 1, scale, i - 1, lk);
    }
    if (dead && lead) {
        a.logE[b] = -INFINITY;
        if (a.alive) a.alive[b] = -1;
    }
}

}  // namespace blg
