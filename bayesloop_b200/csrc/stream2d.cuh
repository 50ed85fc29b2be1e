// stream2d.cuh -- K3f/K4f: forward / backward passes for 2-D grids that do not fit in shared memory (200^2, 256^2,
// 512^2 ...) when the transition program consists of GaussianRandomWalks (at most one per axis) and prior resets
// (BASELINE.json configs[2], [3], [4]: Gaussian / ScaledAR1 grids, GRW on both parameters, change-points).
//
// One persistent CTA per SM loops over combos.  The state lives in a global scratch that stays in L2 and is kept
// UNNORMALISED; per time step it is touched by at most two fused tile phases:
//   phase 1  tiles with complete lines of the first convolution axis are staged in shared memory; on the way in the
//            previous normaliser is applied (lazily) and the normalised row of the PREVIOUS step is flushed to HBM
//            (forward) / the smoothed posterior is written and beta*likelihood formed (backward); convolution -> B
//   phase 2  tiles along the second axis; the epilogue multiplies with the likelihood row of this step, accumulates
//            the evidence increment (forward) or sum(alpha*beta), sum(beta) for the NEXT step (backward) and writes
//            the new state.
// So the only full-grid traffic is: state read/write per phase (L2), the likelihood row (L2, shared by all
// combos), and the compulsory HBM rows (forward: one store; backward: one load + one store).  Semantics are the
// same as the generic kernels (core.py:372-417, :434-470; transitionModels.py:96-115, :300-312).
#pragma once

#include "common.cuh"
#include "fast1d.cuh"

namespace blg {

constexpr int kM2d = 9;  // outputs per work item of the 2-D tile convolution

struct Stream2dOps {
    int k0, k1;        // program index of the GRW acting on axis 0 / axis 1 (-1: none)
    int pre, post;     // program index of a RESET before all / after all GRWs (-1: none)
    bool ok;
};

// Host and device share this classification: which programs the fused 2-D kernels understand.
__host__ __device__ inline Stream2dOps classify2d(int n_ops, const int *kind, const int *axis) {
    Stream2dOps o;
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

struct S2d {
    double *A, *Bf, *tile, *W0, *W1;
    RedScratch rs;
    double sig0, sig1, parPre, parPost;
    int R0, R1;
    int w0[4], w1[4], wPre[4], wPost[4];
    Stream2dOps ops;
};

__device__ __forceinline__ bool s2d_setup(const PassArgs &a, double *sm, long long b, S2d &s) {
    s.A = a.scratch + (size_t)blockIdx.x * 2 * a.Gp;
    s.Bf = s.A + a.Gp;
    s.tile = sm + a.off_tile;
    s.rs.buf = sm + a.off_misc;
    s.rs.phase = 0;
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
    if (!(s.sig0 > 0.0) || s.R0 <= 0) s.R0 = 0;
    if (!(s.sig1 > 0.0) || s.R1 <= 0) s.R1 = 0;
    bool ok = true;
    s.W0 = sm + a.off_w + (s.ops.k0 >= 0 ? a.pg.w_off[s.ops.k0] : 0);
    s.W1 = sm + a.off_w + (s.ops.k1 >= 0 ? a.pg.w_off[s.ops.k1] : 0);
    __syncthreads();  // previous combo is done with the weight tables
    if (s.ops.k0 >= 0) {
        if ((2 * s.R0 + kM2d) / kM2d * (kM2d + 1) > a.pg.w_len[s.ops.k0])
            ok = false;
        else if (s.R0 > 0)
            build_weights_chunked<kM2d>(s.W0, a.pg.w_len[s.ops.k0], s.sig0, s.R0, s.rs);
    }
    if (s.ops.k1 >= 0) {
        if ((2 * s.R1 + kM2d) / kM2d * (kM2d + 1) > a.pg.w_len[s.ops.k1])
            ok = false;
        else if (s.R1 > 0)
            build_weights_chunked<kM2d>(s.W1, a.pg.w_len[s.ops.k1], s.sig1, s.R1, s.rs);
    }
    return ok;
}

__device__ __forceinline__ bool in_window(const int *w, long long idx, bool backward) {
    return idx >= (long long)w[backward ? 2 : 0] && idx < (long long)w[backward ? 3 : 1];
}

// exact e / d for 0 <= e < 2^32 / d with one multiply-high (d fixed per stage)
struct FastDiv {
    unsigned m, d;
    __device__ __forceinline__ explicit FastDiv(int dd) : m((unsigned)((0x100000000ull + (unsigned)dd - 1) / (unsigned)dd)), d((unsigned)dd) {}
    __device__ __forceinline__ int div(int e) const { return d == 1 ? e : (int)__umulhi((unsigned)e, m); }
};

struct Raw3 {
    double x, y, z;
};

// Grid sweep with memory-level parallelism: `load(g)` only reads (up to three values per cell), `apply(g, raw)`
// transforms / stores.  Four cells per thread are in flight before the first dependent instruction, which is what
// hides the L2 latency of the streamed state (one CTA per SM, 32 warps).
template <typename Cell, typename Load, typename Apply, typename Sink>
__device__ __forceinline__ void sweep4(int count, Cell cell, Load load, Apply apply, Sink sink) {
    const int nt = blockDim.x;
    for (int e0 = threadIdx.x; e0 < count; e0 += 4 * nt) {
        Raw3 raw[4];
        int g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * nt;
            g[u] = e < count ? cell(e) : -1;
            if (g[u] >= 0) raw[u] = load(g[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (g[u] >= 0) sink(e0 + u * nt, g[u], apply(g[u], raw[u]));  // apply may return a value or the raw triple
    }
}

// conv_item of fast1d.cuh with an element stride: M outputs of one line, inputs es doubles apart (es = 1 along rows,
// es = tile width along columns), halo already in the tile, weights in the chunk-padded layout.
template <int M>
__device__ __forceinline__ void conv_item_strided(const double *__restrict__ line, int es, int i0, int R,
                                                  const double *__restrict__ W, double (&acc)[M]) {
    constexpr int MP = M + 1;
    const double *p = line + (long long)(i0 - R) * es;
    double win[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        win[m] = p[m * es];
        acc[m] = 0.0;
    }
    p += M * es;
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
            for (int m = 0; m < M; ++m) acc[m] = fma(w[u], win[(u + m) % M], acc[m]);
            win[u] = p[u * es];
        }
        p += M * es;
        wc += MP;
    }
}

// One convolution stage over the whole grid, tile by tile.  A tile holds L complete lines of the convolution axis
// WITH a reflected halo of H = R + 2M cells on both ends (so the inner loop has no boundary logic), plus an output
// tile of the same L lines:
//   1. sweep4: coalesced, 4-deep batched loads of the tile cells (`load`/`apply`; apply may flush a row to HBM)
//   2. halo fill from the tile interior (general reflect, valid for R >= n)
//   3. register-blocked convolution (conv_item_strided) into the output tile
//   4. sweep4 over the output tile: `eload(g)` (likelihood row, next alpha row ...) batched, `eapply` -> dst[g]
template <typename Load, typename Apply, typename ELoad, typename EApply>
__device__ __forceinline__ void conv_stage(const PassArgs &a, const S2d &s, int axis, int R, const double *W, double *dst,
                                           Load load, Apply apply, ELoad eload, EApply eapply) {
    constexpr int M = kM2d;
    const int n0 = a.pb.n0, n1 = a.pb.n1;
    const int n = axis == 1 ? n1 : n0, other = axis == 1 ? n0 : n1;
    const int H = R + 2 * M;
    const int ext = n + 2 * H;
    int L = a.tile_doubles / (ext + n);  // lines per tile (input with halo + output)
    if (axis == 0 && L > 32) L = L / 32 * 32;
    L = max(1, min(L, other));
    double *tin = s.tile;
    const int S = (n + M - 1) / M;
    const FastDiv dn(n), dS(S), d2H(2 * H);
    for (int l0 = 0; l0 < other; l0 += L) {
        const int nl = min(L, other - l0);
        const FastDiv dnl(nl);
        double *tout = tin + (size_t)nl * ext;
        // 1. interior cells.  axis 1: tin[l][H + j] (pitch ext);  axis 0: tin[(H + i)][c] (pitch nl)
        if (axis == 1) {
            sweep4(nl * n, [&](int e) { return l0 * n1 + e; }, load, apply,  // n == n1: rows are contiguous
                   [&](int e, int, double v) { tin[e + dn.div(e) * 2 * H + H] = v; });
        } else {
            sweep4(n * nl,
                   [&](int e) {
                       const int i = dnl.div(e);
                       return i * n1 + l0 + (e - i * nl);
                   },
                   load, apply, [&](int e, int, double v) { tin[H * nl + e] = v; });
        }
        __syncthreads();
        // 2. reflected halo (NI_EXTEND_REFLECT, any R)
        for (int e = threadIdx.x; e < 2 * H * nl; e += blockDim.x) {
            int l, k;
            if (axis == 1) {
                l = d2H.div(e);
                k = e - l * 2 * H;
            } else {
                k = dnl.div(e);
                l = e - k * nl;
            }
            const int pos = k < H ? k - H : n + (k - H);  // extended index in [-H, 0) or [n, n+H)
            const int src = reflect_any(pos, n);
            if (axis == 1)
                tin[l * ext + H + pos] = tin[l * ext + H + src];
            else
                tin[(H + pos) * nl + l] = tin[(H + src) * nl + l];
        }
        __syncthreads();
        // 3. convolution: item = (line, segment of M outputs)
        const int items = nl * S;
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            int l, sg;
            if (axis == 1) {
                l = dS.div(w);
                sg = w - l * S;
            } else {
                sg = dnl.div(w);
                l = w - sg * nl;
            }
            const int i0 = sg * M;
            double acc[M];
            if (axis == 1) {
                conv_item_strided<M>(tin + (size_t)l * ext + H, 1, i0, R, W, acc);
#pragma unroll
                for (int m = 0; m < M; ++m)
                    if (i0 + m < n) tout[l * n + i0 + m] = acc[m];
            } else {
                conv_item_strided<M>(tin + (size_t)H * nl + l, nl, i0, R, W, acc);
#pragma unroll
                for (int m = 0; m < M; ++m)
                    if (i0 + m < n) tout[(i0 + m) * nl + l] = acc[m];
            }
        }
        __syncthreads();
        // 4. coalesced epilogue + store
        if (axis == 1) {
            sweep4(nl * n, [&](int e) { return l0 * n1 + e; }, eload, [&](int, const Raw3 &r) { return r; },
                   [&](int e, int g, const Raw3 &r) { dst[g] = eapply(g, r, tout[e]); });
        } else {
            sweep4(n * nl,
                   [&](int e) {
                       const int i = dnl.div(e);
                       return i * n1 + l0 + (e - i * nl);
                   },
                   eload, [&](int, const Raw3 &r) { return r; },
                   [&](int e, int g, const Raw3 &r) { dst[g] = eapply(g, r, tout[e]); });
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double lik_at(const PassArgs &a, const LikTables &tb, long long t, int g) {
    if (a.pb.om_kind == BLG_OM_TABLE) return __ldg(a.lik_table + t * (long long)a.pb.G + g);
    const int i0 = g / a.pb.n1, i1 = g - i0 * a.pb.n1;
    return lik_cell(a.pb, tb, a.steps + t * a.pb.ncols_eff, i0, i1);
}

// ------------------------------------------------------------------------------------------------ K3f forward
template <int NT>
__global__ void __launch_bounds__(NT, 1) fwd_stream2d_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const int G = pb.G;
    const long long T = a.T;
    LikTables tb;  // only used when the likelihood is not tabulated (few combos): tables straight from global memory
    tb.A0 = pb.tabA[0];
    tb.A1 = pb.tabA[1];
    tb.A2 = pb.tabA[2];
    tb.B0 = pb.tabB[0];
    tb.B1 = pb.tabB[1];
    for (long long slot = blockIdx.x; slot < a.B; slot += gridDim.x) {
        const long long b = a.order ? a.order[slot] : slot;
        S2d s;
        if (!s2d_setup(a, sm, b, s)) {
            if (threadIdx.x == 0) {
                a.logE[b] = NAN;
                if (a.alive) a.alive[b] = -2;
            }
            continue;
        }
        double *A = s.A, *Bf = s.Bf;
        {
            const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)G : a.prior;
            for (int g = threadIdx.x; g < G; g += blockDim.x) A[g] = init[g];
        }
        __syncthreads();
        const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
        double *seq = store ? a.alpha_seq + b * T * (long long)G : nullptr;
        LogProduct lp;
        lp.init();
        bool dead = false;
        double kappa = 1.0;

        for (long long t = 0; t < T; ++t) {
            const bool trans = t > 0 || (a.flags & BLG_F_TRANSITION_FIRST);
            const long long idx = t - 1;
            const bool act0 = trans && s.ops.k0 >= 0 && s.R0 > 0 && in_window(s.w0, idx, false);
            const bool act1 = trans && s.ops.k1 >= 0 && s.R1 > 0 && in_window(s.w1, idx, false);
            const bool pre = trans && s.ops.pre >= 0 && in_window(s.wPre, idx, false);
            const bool post = trans && s.ops.post >= 0 && in_window(s.wPost, idx, false);
            double *row = (store && t > 0) ? seq + (t - 1) * (long long)G : nullptr;
            const double kap = kappa, pPre = s.parPre, pPost = s.parPost;
            const double *rb = a.reset_base;
            // input of the first stage: previous state with its lazy normaliser (and the row flush of alpha[t-1])
            const double *Ard = A;
            auto load = [&](int g) {
                Raw3 r;
                r.x = Ard[g];
                r.y = pre ? __ldg(rb + g) : 0.0;
                r.z = 0.0;
                return r;
            };
            auto fill = [&](int g, const Raw3 &r) {
                const double x = r.x * kap;
                if (row) __stcs(row + g, x);  // core.py:389, :408: normalised filtering distribution of step t-1
                return pre ? r.y * pPre : x;
            };
            double part = 0.0;
            auto eload = [&](int g) {
                Raw3 r;
                r.x = lik_at(a, tb, t, g);
                r.y = post ? __ldg(rb + g) : 0.0;
                r.z = 0.0;
                return r;
            };
            auto eapply = [&](int, const Raw3 &r, double v) {  // core.py:375-382: prior * likelihood
                if (post) v = r.y * pPost;
                const double y = v * r.x;
                part += y;
                return y;
            };
            auto nol = [&](int) { return Raw3(); };
            auto pass = [&](int, const Raw3 &, double v) { return v; };
            // order of the two convolutions = program order
            const bool first0 = s.ops.k0 >= 0 && (s.ops.k1 < 0 || s.ops.k0 < s.ops.k1);
            const int nact = (act0 ? 1 : 0) + (act1 ? 1 : 0);
            double *Bw = Bf;
            const double *Brd = Bf;
            auto rdl = [&](int g) {
                Raw3 r;
                r.x = Brd[g];
                r.y = r.z = 0.0;
                return r;
            };
            auto rda = [&](int, const Raw3 &r) { return r.x; };
            if (post || nact == 0) {  // no convolution (or its result is discarded by a trailing reset): one sweep
                sweep4(G, [&](int e) { return e; },
                       [&](int g) {
                           Raw3 r = load(g);
                           const Raw3 e = eload(g);
                           r.z = e.x;
                           if (post) r.y = e.y;
                           return r;
                       },
                       [&](int g, const Raw3 &r) {
                           const double v = fill(g, r);
                           Raw3 e;
                           e.x = r.z;
                           e.y = r.y;
                           e.z = 0.0;
                           return eapply(g, e, v);
                       },
                       [&](int, int g, double v) { Bw[g] = v; });
            } else if (nact == 1) {
                if (act0)
                    conv_stage(a, s, 0, s.R0, s.W0, Bf, load, fill, eload, eapply);
                else
                    conv_stage(a, s, 1, s.R1, s.W1, Bf, load, fill, eload, eapply);
            } else {
                if (first0) {
                    conv_stage(a, s, 0, s.R0, s.W0, Bf, load, fill, nol, pass);
                    __syncthreads();
                    conv_stage(a, s, 1, s.R1, s.W1, A, rdl, rda, eload, eapply);
                } else {
                    conv_stage(a, s, 1, s.R1, s.W1, Bf, load, fill, nol, pass);
                    __syncthreads();
                    conv_stage(a, s, 0, s.R0, s.W0, A, rdl, rda, eload, eapply);
                }
            }
            if (post || nact < 2) {  // new state sits in Bf
                double *tmp = A;
                A = Bf;
                Bf = tmp;
            }
            const double norm = block_sum(part, s.rs);  // core.py:385 (also makes the new state visible)
            if (!(norm > 0.0)) {                         // core.py:388-400
                dead = true;
                break;
            }
            kappa = fast_rcp(norm);
            if (threadIdx.x == 0) {
                lp.mul(norm);                                         // core.py:403
                if (a.local) a.local[b * T + t] = norm * pb.lc_prod;  // core.py:404
            }
        }
        if (!dead) {
            if (store) {
                double *row = seq + (T - 1) * (long long)G;
                for (int g = threadIdx.x; g < G; g += blockDim.x) __stcs(row + g, A[g] * kappa);
            }
            if ((a.flags & BLG_F_SAVE_STATE) && a.final_state) {
                double *fs = a.final_state + b * (long long)G;
                for (int g = threadIdx.x; g < G; g += blockDim.x) fs[g] = A[g] * kappa;
            }
        }
        if (threadIdx.x == 0) {
            double logE = lp.log_value();
            if (dead)
                logE = -INFINITY;
            else if (!(a.flags & BLG_F_INIT_STATE))
                logE += log(pb.lc_prod);  // core.py:417
            a.logE[b] = logE;
            if (a.alive) a.alive[b] = dead ? 0 : 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4f backward
template <int NT>
__global__ void __launch_bounds__(NT, 1) bwd_stream2d_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const int G = pb.G;
    const long long T = a.T;
    LikTables tb;
    tb.A0 = pb.tabA[0];
    tb.A1 = pb.tabA[1];
    tb.A2 = pb.tabA[2];
    tb.B0 = pb.tabB[0];
    tb.B1 = pb.tabB[1];
    const bool acc = (a.flags & BLG_F_ACCUMULATE) != 0;
    for (long long slot = blockIdx.x; slot < a.B; slot += gridDim.x) {
        const long long b = a.order ? a.order[slot] : slot;
        if (a.alive && a.alive[b] != 1) continue;  // the forward pass aborted (core.py:400)
        S2d s;
        if (!s2d_setup(a, sm, b, s)) continue;
        double *A = s.A, *Bf = s.Bf;
        double *seq = a.alpha_seq + b * T * (long long)G;
        const double wgt = acc ? exp(a.log_weight[b]) : 0.0;
        // beta = 1/G (core.py:424-425); sums for the first posterior
        double sab = 0.0, sbb = 0.0;
        {
            const double beta0 = 1.0 / (double)G;
            const double *al = seq + (T - 1) * (long long)G;
            for (int g = threadIdx.x; g < G; g += blockDim.x) {
                A[g] = beta0;
                sab += al[g] * beta0;
                sbb += beta0;
            }
        }
        block_sum2(sab, sbb, s.rs);
        bool dead = false;

        for (long long i = T - 1; i >= 0; --i) {
            if (!(sab > 0.0) || !(sbb > 0.0)) {  // core.py:440-452
                dead = true;
                break;
            }
            const double inv = fast_rcp(sab);  // posterior = alpha*beta / sum(alpha*beta)  core.py:436-441
            const double kb = fast_rcp(sbb);   // core.py:470 (beta only enters scale-free expressions)
            const bool act0 = s.ops.k0 >= 0 && s.R0 > 0 && in_window(s.w0, i, true);
            const bool act1 = s.ops.k1 >= 0 && s.R1 > 0 && in_window(s.w1, i, true);
            const bool pre = s.ops.pre >= 0 && in_window(s.wPre, i, true);
            const bool post = s.ops.post >= 0 && in_window(s.wPost, i, true);
            double *row = seq + i * (long long)G;
            const double *nextRow = i > 0 ? seq + (i - 1) * (long long)G : nullptr;
            double *av = acc ? a.avg + i * (long long)G : nullptr;
            const double pPre = s.parPre, pPost = s.parPost;
            const double *rb = a.reset_base;
            double q = 0.0;
            // input of the first stage: beta*likelihood (core.py:467); on the way the smoothed posterior of step i is
            // written (core.py:441) and sum(post/lik) accumulated (core.py:463)
            const double *Ard = A;
            auto load = [&](int g) {
                Raw3 r;
                r.x = Ard[g];
                r.y = row[g];
                r.z = lik_at(a, tb, i, g);
                return r;
            };
            auto fill = [&](int g, const Raw3 &r) {
                const double beta = r.x;
                const double lk = r.z;
                const double p = r.y * beta * inv;
                q += fast_div(p, lk);
                if (acc) {
                    if (wgt > 0.0) atomicAdd(av + g, wgt * (p < kTiny ? kTiny : p));
                } else {
                    __stcs(row + g, p);
                }
                return (pre || post) ? (post ? 0.0 : __ldg(rb + g) * pPre) : beta * kb * lk;
            };
            sab = 0.0;
            sbb = 0.0;
            auto eload = [&](int g) {
                Raw3 r;
                r.x = nextRow ? nextRow[g] : 0.0;
                r.y = post ? __ldg(rb + g) : 0.0;
                r.z = 0.0;
                return r;
            };
            auto eapply = [&](int, const Raw3 &r, double v) {  // new beta (unnormalised) + the sums the next step needs
                if (post) v = r.y * pPost;
                sab = fma(r.x, v, sab);
                sbb += v;
                return v;
            };
            auto nol = [&](int) { return Raw3(); };
            auto pass = [&](int, const Raw3 &, double v) { return v; };
            const bool first0 = s.ops.k0 >= 0 && (s.ops.k1 < 0 || s.ops.k0 < s.ops.k1);
            const int nact = (act0 ? 1 : 0) + (act1 ? 1 : 0);
            double *Bw = Bf;
            const double *Brd = Bf;
            auto rdl = [&](int g) {
                Raw3 r;
                r.x = Brd[g];
                r.y = r.z = 0.0;
                return r;
            };
            auto rda = [&](int, const Raw3 &r) { return r.x; };
            if (post || nact == 0) {
                sweep4(G, [&](int e) { return e; }, load, fill,
                       [&](int, int g, double v) { Bw[g] = eapply(g, eload(g), v); });
            } else if (nact == 1) {
                if (act0)
                    conv_stage(a, s, 0, s.R0, s.W0, Bf, load, fill, eload, eapply);
                else
                    conv_stage(a, s, 1, s.R1, s.W1, Bf, load, fill, eload, eapply);
            } else {
                if (first0) {
                    conv_stage(a, s, 0, s.R0, s.W0, Bf, load, fill, nol, pass);
                    __syncthreads();
                    conv_stage(a, s, 1, s.R1, s.W1, A, rdl, rda, eload, eapply);
                } else {
                    conv_stage(a, s, 1, s.R1, s.W1, Bf, load, fill, nol, pass);
                    __syncthreads();
                    conv_stage(a, s, 0, s.R0, s.W0, A, rdl, rda, eload, eapply);
                }
            }
            if (post || nact < 2) {
                double *tmp = A;
                A = Bf;
                Bf = tmp;
            }
            q = block_sum(q, s.rs);
            block_sum2(sab, sbb, s.rs);
            if (threadIdx.x == 0 && a.local) a.local[b * T + i] = fast_div(1.0, q * pb.lc_prod);  // core.py:463
            if (i == 0) break;
        }
        if (dead && threadIdx.x == 0) {
            a.logE[b] = -INFINITY;
            if (a.alive) a.alive[b] = -1;
        }
    }
}

}  // namespace blg
