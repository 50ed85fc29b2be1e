// fast1d.cuh -- K1f/K2f: specialisation of the resident kernels for the headline shape of the sweep: a 1-D grid and
// a transition program that is ONE GaussianRandomWalk (BASELINE.json configs[0], configs[1]).
//
// The time recursion is a chain of T dependent steps per combo, and a sweep only has a few hundred combos, so the
// kernels are built to minimise the LATENCY of one step (measured on B200, see profiles/):
//   * the state lives in shared memory WITH a reflected halo on both sides (every writer mirrors the cells near
//     the edges): the convolution inner loop is one LDS + M FMAs per tap, no boundary logic;
//   * each thread owns M consecutive cells for the whole kernel; convolution outputs stay in REGISTERS and are
//     consumed in place by the likelihood multiply (forward) / the posterior product (backward);
//   * ONE CTA barrier per step: the state is kept UNNORMALISED in shared memory and double buffered; its sum (the
//     evidence increment) is reduced with warp shuffles + one shared-memory exchange that rides on that barrier,
//     and the normaliser is applied lazily in the next step's multiply -- the reduction, the reciprocal and the
//     log-evidence bookkeeping leave the critical path.  Rows written to HBM are normalised exactly as before
//     (from registers, after the barrier);
//   * the likelihood row lik[t][.] (shared by all combos of the call, L2 resident) is fetched before the
//     convolution and consumed after it; backward: alpha[t] arrives through a double-buffered bulk-async (TMA)
//     prefetch with mbarrier completion, two steps ahead.
// Semantics are identical (core.py:372-417, :424-470; transitionModels.py:96-115): every quantity the reference
// normalises is either scale-free downstream (beta, alpha inside the backward product) or normalised on output.
#pragma once

#include "common.cuh"

namespace blg {

// write v to cell i of a haloed line and to its mirror images inside the halo (valid for halo <= n)
__device__ __forceinline__ void store_mirrored(double *line, int i, int n, int halo, double v) {
    line[i] = v;
    if (i < halo) line[-1 - i] = v;
    if (i >= n - halo) line[2 * n - 1 - i] = v;
}

// Weights are stored in chunks of M taps padded to MP = M+1 doubles (16-byte aligned chunks), so one chunk is
// fetched with (M-1)/2 LDS.128 + 1 LDS.64 broadcast loads instead of M LDS.64: the shared-memory pipe (one per SM,
// 128 B/clk) stays below the FP64 pipe (64 FMA/clk/SM): per chunk 2*M wavefronts for the inputs + (M+1)/2 for the
// weights against M*M/2 cycles of DFMA.
template <int M>
__device__ __forceinline__ void conv_item(const double *__restrict__ line, int i0, int R, const double *__restrict__ W,
                                          double (&acc)[M]) {
    static_assert(M % 2 == 1, "M must be odd: lanes stride M doubles => conflict-free 64-bit shared loads");
    constexpr int MP = M + 1;
    const double *p = line + (i0 - R);
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = p[m];
        acc[m] = 0.0;
    }
    p += M;
    // exact tap count: full chunks of M taps, then a guarded remainder (the zero-padded last chunk cost ~6 % of the
    // DFMAs of a reference-like sweep: (M-1)/2 of 2R+1 ~ 70 taps on average); adding w = 0 taps changes nothing, so
    // the results are bit-identical either way
    const int taps = 2 * R + 1;
    const int chunks = taps / M, rem = taps - chunks * M;
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
            for (int m = 0; m < M; ++m) acc[m] = fma(w[u], win[(u + m) % M], acc[m]);
            win[u] = p[u];
        }
        p += M;
        wc += MP;
    }
    if (rem > 0) {
        double w[M];
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
            const double2 t = reinterpret_cast<const double2 *>(wc)[k];
            w[2 * k] = t.x;
            w[2 * k + 1] = t.y;
        }
        w[M - 1] = wc[M - 1];
#pragma unroll
        for (int u = 0; u < M - 1; ++u) {
            if (u >= rem) break;  // uniform over the CTA
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma(w[u], win[(u + m) % M], acc[m]);
            if (u + 1 < rem) win[u] = p[u];
        }
    }
}

// Gaussian weights in the chunk-padded layout of conv_item (see build_weights in common.cuh for the formula).
// Returns the normalising factor 1/sum (1 for the identity kernel).
template <int M>
__device__ __forceinline__ double build_weights_chunked(double *W, int len, double sigma, int R, RedScratch &rs) {
    constexpr int MP = M + 1;
    double part = 0.0;
    if (!(sigma > 0.0) || R <= 0) {  // transitionModels.py:110-113: identity
        for (int q = threadIdx.x; q < len; q += blockDim.x) W[q] = q == 0 ? 1.0 : 0.0;
        __syncthreads();
        return 1.0;
    }
    const double h = -0.5 / (sigma * sigma);
    for (int q = threadIdx.x; q < len; q += blockDim.x) {
        const int c = q / MP, u = q - c * MP, j = c * M + u;
        double v = 0.0;
        if (u < M && j <= 2 * R) {
            const double x = (double)(j - R);
            v = exp(h * x * x);
        }
        W[q] = v;
        part += v;
    }
    const double inv = 1.0 / block_sum(part, rs);
    for (int q = threadIdx.x; q < len; q += blockDim.x) W[q] *= inv;
    __syncthreads();
    return inv;
}

// The same weights in the chunk layout of another M, normalised with the factor build_weights_chunked returned: the two
// tables hold bit-identical values (uneven split of the warp-specialised kernels, fast1d_ws.cuh).
template <int M>
__device__ __forceinline__ void copy_weights_chunked(double *W, int len, double sigma, int R, double inv) {
    constexpr int MP = M + 1;
    const bool identity = !(sigma > 0.0) || R <= 0;
    const double h = identity ? 0.0 : -0.5 / (sigma * sigma);
    for (int q = threadIdx.x; q < len; q += blockDim.x) {
        const int c = q / MP, u = q - c * MP, j = c * M + u;
        double v = 0.0;
        if (identity) {
            v = q == 0 ? 1.0 : 0.0;
        } else if (u < M && j <= 2 * R) {
            const double x = (double)(j - R);
            v = exp(h * x * x) * inv;
        }
        W[q] = v;
    }
    __syncthreads();
}

// Likelihood of the thread's M consecutive cells for one time step: the observation-model switch is outside the
// cell loop and the M exponentials are independent instruction streams (ILP hides the ~25-DFMA chain of exp()).
template <int M>
__device__ __forceinline__ void lik_cells(const PassArgs &a, const LikTables &tb, const StepC &s0, const StepC *sc,
                                          long long t, int i0, int n, double (&lk)[M]) {
    const DevProblem &pb = a.pb;
    if (pb.om_kind == BLG_OM_TABLE) {
#pragma unroll
        for (int m = 0; m < M; ++m) lk[m] = i0 + m < n ? __ldg(a.lik_table + t * (long long)n + i0 + m) : 0.0;
        return;
    }
    if (pb.ncols_eff == 1) {
        if (s0.skip != 0.0) {  // missing data: likelihood of ones (observationModels.py:53-54)
#pragma unroll
            for (int m = 0; m < M; ++m) lk[m] = 1.0;
            return;
        }
        double arg[M];
        if (pb.om_kind == BLG_OM_POISSON) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = min(i0 + m, n - 1);
                arg[m] = fma(s0.d0, tb.A1[li], -tb.A0[li]) - s0.c;
            }
        } else if (pb.om_kind == BLG_OM_WHITE_NOISE) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = min(i0 + m, n - 1);
                arg[m] = fma(-s0.d0 * s0.d0, tb.A0[li], tb.A1[li]);
            }
        } else if (pb.om_kind == BLG_OM_GAUSSIAN_MEAN) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = min(i0 + m, n - 1);
                const double r = s0.d0 - tb.A0[li];
                arg[m] = fma(-r * r, s0.d1, s0.c);
            }
        } else {  // BERNOULLI
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = min(i0 + m, n - 1);
                lk[m] = s0.d0 != 0.0 ? tb.A0[li] : 1.0 - tb.A0[li];
            }
            return;
        }
#pragma unroll
        for (int m = 0; m < M; ++m) lk[m] = exp(arg[m]);
        return;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) lk[m] = lik_cell(pb, tb, sc, min(i0 + m, n - 1), 0);
}

struct Fast1dSetup {
    double *buf0, *buf1;  // interior pointers of the two haloed state buffers
    double *W;
    double *P;            // per-warp partial sums: [2 parities][3 values][kMaxWarps]
    LikTables tb;
    RedScratch rs;
    double sigma;
    int R, f_lo, f_hi, b_lo, b_hi;
};

template <int M>
__device__ __forceinline__ void fast1d_setup(const PassArgs &a, double *sm, long long b, Fast1dSetup &s) {
    const DevProblem &pb = a.pb;
    const int halo = a.halo, pitch = a.Gp + 2 * halo;
    s.buf0 = sm + halo;
    s.buf1 = sm + pitch + halo;
    if (a.off_tab >= 0) {  // per-cell tables only when the likelihood is evaluated in the pass (no shared table)
        double *tab = sm + a.off_tab;
        double *A0 = tab, *A1 = A0 + a.n0p, *A2 = A1 + a.n0p;
        for (int i = threadIdx.x; i < pb.n0; i += blockDim.x) {
            A0[i] = pb.tabA[0] ? pb.tabA[0][i] : 0.0;
            A1[i] = pb.tabA[1] ? pb.tabA[1][i] : 0.0;
            A2[i] = pb.tabA[2] ? pb.tabA[2][i] : 0.0;
        }
        s.tb.A0 = A0;
        s.tb.A1 = A1;
        s.tb.A2 = A2;
        s.tb.B0 = A0;
        s.tb.B1 = A0;
    }
    s.rs.buf = sm + a.off_misc;
    s.rs.phase = 0;
    s.P = sm + a.off_misc + kMiscPartialOffset;
    s.W = sm + a.off_w;
    s.sigma = a.pg.param[b];
    s.R = a.pg.radius[b];
    const int *win = a.pg.window + b * 4;
    s.f_lo = win[0];
    s.f_hi = win[1];
    s.b_lo = win[2];
    s.b_hi = win[3];
    if (!(s.sigma > 0.0) || s.R <= 0) s.R = 0;
    if ((2 * s.R + M) / M * (M + 1) <= a.pg.w_len[0]) build_weights_chunked<M>(s.W, a.pg.w_len[0], s.sigma, s.R, s.rs);
}

template <int M>
__device__ __forceinline__ double tree_sum(const double (&x)[M]) {
    double t[M];
#pragma unroll
    for (int m = 0; m < M; ++m) t[m] = x[m];
#pragma unroll
    for (int w = 1; w < M; w *= 2)
#pragma unroll
        for (int m = 0; m + w < M; m += 2 * w) t[m] += t[m + w];
    return t[0];
}

// ------------------------------------------------------------------------------------------------ K1f forward
template <int M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) fwd_fast1d_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const long long b = a.sm_assign ? combo_of_sm(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6))
                                    : (a.order ? a.order[combo_of_block(a)] : combo_of_block(a));
    if (b < 0) return;
    trace_begin(a, b);
    const int n = pb.G, halo = a.halo;
    const long long T = a.T;
    Fast1dSetup s;
    fast1d_setup<M>(a, sm, b, s);
    if ((2 * s.R + M) / M * (M + 1) > a.pg.w_len[0]) {  // radius beyond blg_program.max_radius
        if (threadIdx.x == 0) {
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
        }
        return;
    }
    const int i0 = threadIdx.x * M;
    const bool owner = i0 < n;
    const bool service = threadIdx.x == blockDim.x - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double *cur = s.buf0, *nxt = s.buf1;
    {
        const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)n : a.prior;
        for (int g = threadIdx.x; g < n; g += blockDim.x) store_mirrored(cur, g, n, halo, init[g]);
    }
    __syncthreads();

    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    double *seq = store ? a.alpha_seq + b * a.seq_stride : nullptr;
    const int nce = pb.ncols_eff;
    const bool table = pb.om_kind == BLG_OM_TABLE;
    LogProduct lp;
    lp.init();
    bool dead = false;
    double kappa = 1.0;  // 1 / sum of the state in `cur` (the normaliser of the previous step, applied lazily)

    for (long long t = 0; t < T; ++t) {
        double v[M], lk[M];
        const StepC *sc = a.steps + t * nce;
        StepC s0;
        if (!table) {
            s0 = sc[0];  // issued before the convolution, consumed after it
        } else if (owner) {  // likelihood row of this step (shared by all combos, L2 resident)
#pragma unroll
            for (int m = 0; m < M; ++m) lk[m] = i0 + m < n ? __ldg(a.lik_table + t * (long long)n + i0 + m) : 0.0;
        }
        const bool trans = (t > 0 || (a.flags & BLG_F_TRANSITION_FIRST)) && (t - 1 >= s.f_lo) && (t - 1 < s.f_hi);
        double part = 0.0;
        if (owner) {
            if (trans && s.R > 0) {
                conv_item<M>(cur, i0, s.R, s.W, v);  // transitionModels.py:111
            } else {
#pragma unroll
                for (int m = 0; m < M; ++m) v[m] = cur[i0 + m];
            }
            if (!table) lik_cells<M>(a, s.tb, s0, sc, t, i0, n, lk);
            // alpha <- prior * likelihood (core.py:375-382); prior = T(alpha[t-1]) carries the lazy normaliser kappa
#pragma unroll
            for (int m = 0; m < M; ++m) {
                v[m] = i0 + m < n ? v[m] * kappa * lk[m] : 0.0;
                if (i0 + m < n) store_mirrored(nxt, i0 + m, n, halo, v[m]);
            }
            part = tree_sum<M>(v);
        }
        part = warp_sum(part);
        double *P = s.P + (t & 1) * 3 * kMaxWarps;
        if (lane == 0) P[warp] = part;
        __syncthreads();  // the only barrier of the step: new state and its partial sums are visible
        double norm = 0.0;  // core.py:385
        for (int w = 0; w < nw; ++w) norm += P[w];
        if (!(norm > 0.0)) {  // core.py:388-400
            dead = true;
            break;
        }
        kappa = fast_rcp(norm);
        if (store && owner) {  // core.py:389, :408 -- normalised filtering distribution, straight from registers
            double *row = seq + t * (long long)n;
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (i0 + m < n) __stcs(row + i0 + m, v[m] * kappa);
        }
        if (service) {
            lp.mul(norm);                                         // core.py:403
            if (a.local) a.local[b * a.row_stride + t] = norm * pb.lc_prod;  // core.py:404
        }
        double *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    if (!dead && (a.flags & BLG_F_SAVE_STATE) && a.final_state) {
        double *fs = a.final_state + b * (long long)n;
        for (int g = threadIdx.x; g < n; g += blockDim.x) fs[g] = cur[g] * kappa;
    }
    if (service) {
        double logE = lp.log_value();
        if (dead)
            logE = -INFINITY;
        else if (!(a.flags & BLG_F_INIT_STATE))
            logE += log(pb.lc_prod);  // core.py:417
        a.logE[b] = logE;
        if (a.alive) a.alive[b] = dead ? 0 : 1;
    }
    trace_end(a);
}

// ------------------------------------------------------------------------------------------------ K2f backward
template <int M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) bwd_fast1d_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const long long b = a.sm_assign ? combo_of_sm(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6))
                                    : (a.order ? a.order[combo_of_block(a)] : combo_of_block(a));
    if (b < 0) return;
    trace_begin(a, b);
    if (a.alive && a.alive[b] != 1) return;  // the forward pass aborted (core.py:400)
    const int n = pb.G, halo = a.halo;
    const long long T = a.T;
    Fast1dSetup s;
    fast1d_setup<M>(a, sm, b, s);
    if ((2 * s.R + M) / M * (M + 1) > a.pg.w_len[0]) return;
    const int i0 = threadIdx.x * M;
    const bool owner = i0 < n;
    const bool service = threadIdx.x == blockDim.x - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const bool acc = (a.flags & BLG_F_ACCUMULATE) != 0;
    const double wgt = acc ? exp(a.log_weight[b]) : 0.0;
    double *cur = s.buf0, *nxt = s.buf1;
    double *seq = a.alpha_seq + b * a.seq_stride;
    const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : seq;  // filtering rows (out-of-place smoothing)
    const bool staged = a.use_bulk != 0;
    double *S[2] = {sm + a.off_stage, sm + a.off_stage + a.Gp};  // alpha[t] staging ring
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + a.off_misc + kMiscBarrierOffset);
    uint32_t ph[2] = {0u, 0u};
    const uint32_t rowBytes = (uint32_t)(n * sizeof(double));
    if (staged) {
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (service) {
            bulk_load(S[(T - 1) & 1], src + (T - 1) * (long long)n, rowBytes, &bars[(T - 1) & 1]);
            if (T >= 2) bulk_load(S[(T - 2) & 1], src + (T - 2) * (long long)n, rowBytes, &bars[(T - 2) & 1]);
        }
    }
    const int nce = pb.ncols_eff;
    const bool table = pb.om_kind == BLG_OM_TABLE;
    double beta[M];
#pragma unroll
    for (int m = 0; m < M; ++m) beta[m] = (owner && i0 + m < n) ? 1.0 / (double)n : 0.0;  // core.py:424-425
    double kb = 1.0;  // keeps the (scale-free) beta recursion in range: 1 / sum(beta) of the previous step
    bool dead = false;
    long long i = T - 1;

    for (; i >= 0; --i) {
        const int sb = (int)(i & 1);
        const StepC *sc = a.steps + i * nce;
        StepC s0;
        double lk[M];
        if (!table) {
            s0 = sc[0];
        } else if (owner) {
#pragma unroll
            for (int m = 0; m < M; ++m) lk[m] = i0 + m < n ? __ldg(a.lik_table + i * (long long)n + i0 + m) : 0.0;
        }
        const double *A;
        if (staged) {
            mbar_wait(&bars[sb], ph[sb]);
            ph[sb] ^= 1u;
            A = S[sb];
        } else {
            A = src + i * (long long)n;
        }
        double pu[M];
        double spu = 0.0, sbeta = 0.0, sql = 0.0;
        if (owner) {
            if (!table) lik_cells<M>(a, s.tb, s0, sc, i, i0, n, lk);  // core.py:455
            double ql[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = i0 + m;
                const double al = li < n ? A[li] : 0.0;
                pu[m] = al * beta[m];                                      // posterior ~ alpha*beta   core.py:436
                ql[m] = li < n ? fast_div(pu[m], lk[m]) : 0.0;             // core.py:463
                if (li < n) store_mirrored(nxt, li, n, halo, beta[m] * kb * lk[m]);  // beta*likelihood  core.py:467
            }
            spu = tree_sum<M>(pu);
            sbeta = tree_sum<M>(beta);
            sql = tree_sum<M>(ql);
        }
        spu = warp_sum(spu);
        sbeta = warp_sum(sbeta);
        sql = warp_sum(sql);
        double *P = s.P + (i & 1) * 3 * kMaxWarps;
        if (lane == 0) {
            P[warp] = spu;
            P[kMaxWarps + warp] = sbeta;
            P[2 * kMaxWarps + warp] = sql;
        }
        __syncthreads();  // the only barrier of the step
        spu = 0.0;
        sbeta = 0.0;
        sql = 0.0;
        for (int w = 0; w < nw; ++w) {
            spu += P[w];
            sbeta += P[kMaxWarps + w];
            sql += P[2 * kMaxWarps + w];
        }
        if (!(spu > 0.0) || !(sbeta > 0.0)) {  // core.py:440-452
            dead = true;
            break;
        }
        if (staged && service && i >= 2)  // everybody is past the barrier: the staging slot is free again
            bulk_load(S[sb], src + (i - 2) * (long long)n, rowBytes, &bars[sb]);
        const double inv = fast_rcp(spu);  // posterior = alpha*beta / sum(alpha*beta)   core.py:439-441
        kb = fast_rcp(sbeta);              // core.py:470, applied lazily (beta only enters scale-free expressions)
        if (owner) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int li = i0 + m;
                if (li < n) {
                    const double p = pu[m] * inv;
                    if (acc) {
                        if (wgt > 0.0) atomicAdd(a.avg + i * (long long)n + li, wgt * (p < kTiny ? kTiny : p));
                    } else {
                        __stcs(seq + i * (long long)n + li, p);
                    }
                }
            }
        }
        if (service && a.local) a.local[b * a.row_stride + i] = fast_div(spu, sql * pb.lc_prod);  // 1/(sum(post/lik)*lc)  core.py:463
        double *tmp = cur;
        cur = nxt;
        nxt = tmp;
        const bool trans = (i >= s.b_lo) && (i < s.b_hi);
        if (owner) {
            if (trans && s.R > 0) {
                conv_item<M>(cur, i0, s.R, s.W, beta);  // transitionModels.py:117-118
            } else {
#pragma unroll
                for (int m = 0; m < M; ++m) beta[m] = i0 + m < n ? cur[i0 + m] : 0.0;
            }
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (i0 + m >= n) beta[m] = 0.0;
        }
    }
    if (dead && staged && i >= 1) mbar_wait(&bars[(i - 1) & 1], ph[(i - 1) & 1]);  // drain the prefetch in flight
    if (dead && service) {
        a.logE[b] = -INFINITY;
        if (a.alive) a.alive[b] = -1;
    }
    trace_end(a);
}

}  // namespace blg
