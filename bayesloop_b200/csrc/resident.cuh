// resident.cuh -- K1/K2: persistent forward / backward kernels with the grid state resident in shared memory.
//
// One CTA owns one hyper-parameter combination for ALL T time steps (the time recursion core.py:372-411 /
// :434-470 is strictly sequential; combos are independent, so no grid-wide synchronisation exists anywhere).
// Per step the CTA fuses: likelihood evaluation -> prior*likelihood -> block reduction (warp shuffles) ->
// normalisation -> log-evidence / local-evidence accumulation -> bulk-async (TMA) store of alpha[t] -> transition
// program (register-blocked reflect convolution for GaussianRandomWalk, clamp+renormalise for RegimeSwitch, prior
// reset for ChangePoint/Independent).  Compulsory HBM traffic: forward 8 B per cell-update (the alpha[t] store);
// backward 8 B read of alpha[t] (bulk-async prefetch, double buffered, mbarrier) + 8 B posterior store (Study) or
// one fp64 reduction into the running average (HyperStudy).
#pragma once

#include "common.cuh"

namespace blg {

struct CellIter {
    int g, i0, i1, dq, dr;
    __device__ __forceinline__ void start(int n1) {
        g = threadIdx.x;
        i0 = g / n1;
        i1 = g - i0 * n1;
        dq = blockDim.x / n1;
        dr = blockDim.x - dq * n1;
    }
    __device__ __forceinline__ void next(int n1) {
        g += blockDim.x;
        i0 += dq;
        i1 += dr;
        if (i1 >= n1) {
            i1 -= n1;
            ++i0;
        }
    }
};

struct Resident {
    double *cur, *oth;      // state buffer and convolution target
    const double *W;        // weight region
    const double *par;      // per-op parameter of this combo
    const int *rad, *win;   // per-op radius / windows of this combo
    RedScratch rs;
    LikTables tb;
    bool pending;           // a bulk store is still reading `cur`
};

// Convolution for grids that do not fit in shared memory (stream kernels): the state lives in global memory (L2),
// tiles that contain COMPLETE lines of the convolution axis are staged in shared memory, so no halo is exchanged:
//   axis 1 (contiguous): tile = a band of rows;   axis 0: tile = a strip of columns (all n0 rows).
template <int M>
__device__ __forceinline__ void conv_global(const double *src, double *dst, const double *W, int R, int n0, int n1,
                                            int axis, double *tile, int tileDoubles) {
    if (axis == 1 || n1 == 1) {
        const int n = n1 == 1 ? n0 : n1, rowsAll = n1 == 1 ? 1 : n0;
        const int rowsPerTile = max(1, tileDoubles / n);
        for (int r0 = 0; r0 < rowsAll; r0 += rowsPerTile) {
            const int rows = min(rowsPerTile, rowsAll - r0);
            const double *g = src + (size_t)r0 * n;
            for (int e = threadIdx.x; e < rows * n; e += blockDim.x) tile[e] = g[e];
            __syncthreads();
            if (R + 2 * M <= n)
                conv_lines<M, true>(tile, dst + (size_t)r0 * n, W, R, n, 1, rows, n);
            else
                conv_lines<M, false>(tile, dst + (size_t)r0 * n, W, R, n, 1, rows, n);
            __syncthreads();
        }
    } else {
        int cols = tileDoubles / n0;
        if (cols > 32) cols = cols / 32 * 32;
        cols = max(1, min(cols, n1));
        for (int c0 = 0; c0 < n1; c0 += cols) {
            const int w = min(cols, n1 - c0);
            for (int e = threadIdx.x; e < n0 * w; e += blockDim.x) {
                const int rr = e / w, cc = e - rr * w;
                tile[e] = src[(size_t)rr * n1 + c0 + cc];
            }
            __syncthreads();
            if (R + 2 * M <= n0)
                conv_lines<M, true>(tile, dst + c0, W, R, n0, w, w, 1, n1, 1);
            else
                conv_lines<M, false>(tile, dst + c0, W, R, n0, w, w, 1, n1, 1);
            __syncthreads();
        }
    }
}

// Transition program of one step (forward: index of the step just processed; backward: current step).
// Every branch is CTA-uniform.  Returns with all threads synchronised on the new state in r.cur.
template <bool STREAM>
__device__ __forceinline__ void apply_ops(const PassArgs &a, Resident &r, long long idx, bool backward, long long b,
                                          double *sm) {
    const DevProblem &pb = a.pb;
    const int G = pb.G;
    int applied = 0;
    for (int k = 0; k < a.pg.n_ops; ++k) {
        const int lo = r.win[4 * k + (backward ? 2 : 0)], hi = r.win[4 * k + (backward ? 3 : 1)];
        if (idx < (long long)lo || idx >= (long long)hi) continue;
        const int kind = a.pg.kind[k];
        const double par = r.par[k];
        if (kind == BLG_OP_GRW) {
            const int R = r.rad[k];
            if (!(par > 0.0) || R <= 0) continue;  // transitionModels.py:110-113; a single tap of weight 1
            const int ax = a.pg.axis[k];
            const int n = ax == 0 ? pb.n0 : pb.n1;
            const int es = ax == 0 ? pb.n1 : 1, nl = ax == 0 ? pb.n1 : pb.n0, ls = ax == 0 ? 1 : pb.n1;
            const double *W = r.W + a.pg.w_off[k];
            if (STREAM)
                conv_global<kConvM>(r.cur, r.oth, W, R, pb.n0, pb.n1, ax, sm + a.off_tile, a.tile_doubles);
            else if (R + 2 * kConvM <= n)
                conv_lines<kConvM, true>(r.cur, r.oth, W, R, n, es, nl, ls);
            else
                conv_lines<kConvM, false>(r.cur, r.oth, W, R, n, es, nl, ls);
            if (r.pending) {  // the alpha[t] bulk store may read r.cur until here
                if (threadIdx.x == 0) bulk_wait_read<0>();
                r.pending = false;
            }
            __syncthreads();
            double *t = r.cur;
            r.cur = r.oth;
            r.oth = t;
            ++applied;
            continue;
        }
        if (r.pending) {  // in-place operators overwrite the buffer the bulk store reads
            if (threadIdx.x == 0) bulk_wait_read<0>();
            r.pending = false;
            __syncthreads();
        }
        ++applied;
        if (kind == BLG_OP_REGIME) {  // transitionModels.py:405-410
            double part = 0.0;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                double v = r.cur[g];
                v = v < par ? par : v;
                r.cur[g] = v;
                part += v;
            }
            const double inv = fast_rcp(block_sum(part, r.rs));
            for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] *= inv;
            __syncthreads();
        } else if (kind == BLG_OP_RESET) {  // transitionModels.py:300-312, :350-360, :801-813
            const double *base = a.reset_base;
            for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] = __ldg(base + g) * par;
            __syncthreads();
        } else if (kind == BLG_OP_NOTEQUAL) {  // transitionModels.py:461-469
            double mx = -INFINITY;
            for (int g = threadIdx.x; g < G; g += blockDim.x) mx = fmax(mx, r.cur[g]);
            mx = block_max(mx, r.rs);
            double part = 0.0;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                const double v = mx - r.cur[g];
                r.cur[g] = v;
                part += v;
            }
            double inv = 1.0 / block_sum(part, r.rs);
            part = 0.0;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                double v = r.cur[g] * inv;
                v = v < par ? par : v;
                r.cur[g] = v;
                part += v;
            }
            inv = 1.0 / block_sum(part, r.rs);
            for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] *= inv;
            __syncthreads();
        }
    }
    if (r.pending) {  // nothing consumed the wait: the next phase writes r.cur in place
        if (threadIdx.x == 0) bulk_wait_read<0>();
        r.pending = false;
        __syncthreads();
    }
    (void)applied;
    (void)b;
}

// Shared set-up of both passes: tables, per-combo parameters, convolution weights.  Returns false (uniformly) if
// the combo's radius exceeds the table the host sized from blg_program.max_radius.
__device__ __forceinline__ bool resident_setup(const PassArgs &a, double *sm, long long b, Resident &r) {
    const DevProblem &pb = a.pb;
    double *tab = sm + a.off_tab;
    double *A0 = tab, *A1 = A0 + a.n0p, *A2 = A1 + a.n0p, *B0 = A2 + a.n0p, *B1 = B0 + a.n1p;
    for (int i = threadIdx.x; i < pb.n0; i += blockDim.x) {
        A0[i] = pb.tabA[0] ? pb.tabA[0][i] : 0.0;
        A1[i] = pb.tabA[1] ? pb.tabA[1][i] : 0.0;
        A2[i] = pb.tabA[2] ? pb.tabA[2][i] : 0.0;
    }
    for (int i = threadIdx.x; i < pb.n1; i += blockDim.x) {
        B0[i] = pb.tabB[0] ? pb.tabB[0][i] : 0.0;
        B1[i] = pb.tabB[1] ? pb.tabB[1][i] : 0.0;
    }
    r.tb.A0 = A0;
    r.tb.A1 = A1;
    r.tb.A2 = A2;
    r.tb.B0 = B0;
    r.tb.B1 = B1;
    double *misc = sm + a.off_misc;
    r.rs.buf = misc;
    r.rs.phase = 0;
    double *par = misc + 4 * kMaxWarps;
    int *ip = reinterpret_cast<int *>(par + BLG_MAX_OPS);
    int *rad = ip, *win = ip + BLG_MAX_OPS;
    const int K = a.pg.n_ops;
    if ((int)threadIdx.x < K) {
        const int k = threadIdx.x;
        par[k] = a.pg.param[b * K + k];
        rad[k] = a.pg.radius[b * K + k];
        for (int q = 0; q < 4; ++q) win[4 * k + q] = a.pg.window[(b * K + k) * 4 + q];
    }
    __syncthreads();
    r.par = par;
    r.rad = rad;
    r.win = win;
    double *W = sm + a.off_w;
    r.W = W;
    bool ok = true;
    for (int k = 0; k < K; ++k) {
        if (a.pg.kind[k] != BLG_OP_GRW) continue;
        const int R = rad[k];
        if (2 * R + 1 + kConvM > a.pg.w_len[k]) {
            ok = false;
            continue;
        }
        if (par[k] > 0.0 && R > 0) build_weights(W + a.pg.w_off[k], a.pg.w_len[k], par[k], R, r.rs);
    }
    r.pending = false;
    return ok;
}

// ------------------------------------------------------------------------------------------------ K1 forward
template <int NT, int MINB, bool STREAM>
__global__ void __launch_bounds__(NT, MINB) fwd_resident_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
  for (long long slot = STREAM ? blockIdx.x : combo_of_block(a); slot < a.B; slot += STREAM ? gridDim.x : a.B) {
    const long long b = a.order ? a.order[slot] : slot;
    const int G = pb.G, n1 = pb.n1;
    const long long T = a.T;
    Resident r;
    r.cur = STREAM ? a.scratch + (size_t)blockIdx.x * 2 * a.Gp : sm;
    r.oth = r.cur + a.Gp;
    if (STREAM) __syncthreads();  // previous combo of this CTA is completely done with shared memory
    const bool ok = resident_setup(a, sm, b, r);
    if (!ok) {
        if (threadIdx.x == 0) {
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
        }
        continue;
    }
    {
        const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)G : a.prior;
        for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] = init[g];
    }
    __syncthreads();

    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    const bool bulk = store && a.use_bulk;
    double *seq = store ? a.alpha_seq + b * a.seq_stride : nullptr;
    const int nce = pb.ncols_eff;
    double logE = 0.0;
    bool dead = false;

    for (long long t = 0; t < T; ++t) {
        if (t > 0 || (a.flags & BLG_F_TRANSITION_FIRST)) apply_ops<STREAM>(a, r, t - 1, false, b, sm);

        // alpha <- prior * likelihood, norm = sum(alpha)          core.py:375-385
        const StepC *sc = a.steps + t * nce;
        double part = 0.0;
        if (pb.om_kind == BLG_OM_TABLE) {
            const double *lt = a.lik_table + t * (long long)G;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                const double v = r.cur[g] * __ldg(lt + g);
                r.cur[g] = v;
                part += v;
            }
        } else if (nce == 1) {
            const StepC s0 = sc[0];
            if (s0.skip == 0.0) {
                CellIter c;
                for (c.start(n1); c.g < G; c.next(n1)) {
                    const double v = r.cur[c.g] * lik_column(pb.om_kind, r.tb, c.i0, c.i1, s0);
                    r.cur[c.g] = v;
                    part += v;
                }
            } else {
                for (int g = threadIdx.x; g < G; g += blockDim.x) part += r.cur[g];
            }
        } else {
            CellIter c;
            for (c.start(n1); c.g < G; c.next(n1)) {
                const double v = r.cur[c.g] * lik_cell(pb, r.tb, sc, c.i0, c.i1);
                r.cur[c.g] = v;
                part += v;
            }
        }
        const double norm = block_sum(part, r.rs);
        if (!(norm > 0.0)) {  // core.py:388-400
            dead = true;
            break;
        }
        const double inv = fast_rcp(norm);
        if (store && !bulk) {
            double *row = seq + t * (long long)G;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                const double v = r.cur[g] * inv;
                r.cur[g] = v;
                __stcs(row + g, v);
            }
        } else {
            for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] *= inv;
        }
        if (threadIdx.x == 0) {
            logE += log(norm);                                          // core.py:403
            if (a.local) a.local[b * a.row_stride + t] = norm * pb.lc_prod;        // core.py:404
        }
        if (bulk) fence_proxy_async();
        __syncthreads();
        if (bulk) {  // core.py:408 -- alpha[t] leaves through the bulk-async copy engine while the transition runs
            if (threadIdx.x == 0) bulk_store(seq + t * (long long)G, r.cur, (uint32_t)(G * sizeof(double)));
            r.pending = true;
        }
    }
    if (bulk && threadIdx.x == 0) bulk_wait_all();
    if (!dead && (a.flags & BLG_F_SAVE_STATE) && a.final_state) {
        double *fs = a.final_state + b * (long long)G;
        for (int g = threadIdx.x; g < G; g += blockDim.x) fs[g] = r.cur[g];
    }
    if (threadIdx.x == 0) {
        if (dead)
            logE = -INFINITY;
        else if (!(a.flags & BLG_F_INIT_STATE))
            logE += log(pb.lc_prod);  // core.py:417
        a.logE[b] = logE;
        if (a.alive) a.alive[b] = dead ? 0 : 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K2 backward
template <int NT, int MINB, bool STREAM>
__global__ void __launch_bounds__(NT, MINB) bwd_resident_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
  for (long long slot = STREAM ? blockIdx.x : combo_of_block(a); slot < a.B; slot += STREAM ? gridDim.x : a.B) {
    const long long b = a.order ? a.order[slot] : slot;
    if (a.alive && a.alive[b] != 1) continue;  // the forward pass aborted (core.py:400)
    const int G = pb.G, n1 = pb.n1;
    const long long T = a.T;
    Resident r;
    r.cur = STREAM ? a.scratch + (size_t)blockIdx.x * 2 * a.Gp : sm;
    r.oth = r.cur + a.Gp;
    if (STREAM) __syncthreads();
    if (!resident_setup(a, sm, b, r)) continue;
    const bool acc = (a.flags & BLG_F_ACCUMULATE) != 0;
    const double wgt = acc ? exp(a.log_weight[b]) : 0.0;
    const double beta0 = 1.0 / (double)G;  // core.py:424-425
    for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] = beta0;

    double *seq = a.alpha_seq + b * a.seq_stride;
    const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : seq;  // filtering rows (out-of-place smoothing)
    const bool staged = !STREAM && a.off_stage >= 0 && a.use_bulk;
    double *S[2] = {sm + (staged ? a.off_stage : 0), sm + (staged ? a.off_stage + a.Gp : 0)};
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + a.off_misc + kMiscBarrierOffset);
    uint32_t ph[2] = {0u, 0u};
    const uint32_t rowBytes = (uint32_t)(G * sizeof(double));
    if (staged) {
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_load(S[(T - 1) & 1], src + (T - 1) * (long long)G, rowBytes, &bars[(T - 1) & 1]);
            if (T >= 2) bulk_load(S[(T - 2) & 1], src + (T - 2) * (long long)G, rowBytes, &bars[(T - 2) & 1]);
        }
    } else {
        __syncthreads();
    }
    const int nce = pb.ncols_eff;
    bool dead = false;
    long long i = T - 1;

    for (; i >= 0; --i) {
        const int sb = (int)(i & 1);
        const double *A;
        if (staged) {
            mbar_wait(&bars[sb], ph[sb]);
            ph[sb] ^= 1u;
            A = S[sb];
        } else {
            A = src + i * (long long)G;
        }
        // posterior ~ alpha * beta                                  core.py:436-441
        double part = 0.0;
        for (int g = threadIdx.x; g < G; g += blockDim.x) part += A[g] * r.cur[g];
        const double norm = block_sum(part, r.rs);
        if (!(norm > 0.0)) {  // core.py:440-452
            dead = true;
            break;
        }
        const double inv = fast_rcp(norm);
        const StepC *sc = a.steps + i * nce;
        double *row = seq + i * (long long)G;
        double *av = acc ? a.avg + i * (long long)G : nullptr;
        double q = 0.0;
        CellIter c;
        for (c.start(n1); c.g < G; c.next(n1)) {
            const double beta = r.cur[c.g];
            const double p = A[c.g] * beta * inv;
            const double lik = pb.om_kind == BLG_OM_TABLE ? __ldg(a.lik_table + i * (long long)G + c.g)
                                                          : lik_cell(pb, r.tb, sc, c.i0, c.i1);  // core.py:455
            q += fast_div(p, lik);                                                               // core.py:463
            if (acc) {
                if (wgt > 0.0) atomicAdd(av + c.g, wgt * (p < kTiny ? kTiny : p));  // core.py:1362-1366
            } else {
                __stcs(row + c.g, p);
            }
            r.cur[c.g] = beta * lik;  // core.py:467 (beta * likelihood)
        }
        q = block_sum(q, r.rs);
        if (staged && threadIdx.x == 0 && i >= 2)  // everybody is past the barrier: S[sb] is free again
            bulk_load(S[sb], src + (i - 2) * (long long)G, rowBytes, &bars[sb]);
        if (threadIdx.x == 0 && a.local) a.local[b * a.row_stride + i] = 1.0 / (q * pb.lc_prod);
        apply_ops<STREAM>(a, r, i, true, b, sm);
        part = 0.0;
        for (int g = threadIdx.x; g < G; g += blockDim.x) part += r.cur[g];
        const double binv = fast_rcp(block_sum(part, r.rs));  // core.py:470
        for (int g = threadIdx.x; g < G; g += blockDim.x) r.cur[g] *= binv;
    }
    if (dead && staged && i >= 1) mbar_wait(&bars[(i - 1) & 1], ph[(i - 1) & 1]);  // drain the prefetch in flight
    if (dead && threadIdx.x == 0) {
        a.logE[b] = -INFINITY;
        if (a.alive) a.alive[b] = -1;
    }
  }
}

}  // namespace blg
