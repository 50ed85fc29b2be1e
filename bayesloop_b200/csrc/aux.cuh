// aux.cuh -- small streaming kernels around the persistent passes: per-step data constants, weighted row sums,
// row normalisation + posterior means, running-average maintenance.  All are HBM/L2-bound elementwise or row
// reductions with coalesced 8-byte accesses; none sits on the critical path of a sweep (O(T*G) once per sweep
// against O(B*T*G) for the passes).
#pragma once

#include "common.cuh"

namespace blg {

// Per (time step, data column) constants: removes lgamma/log/div and the NaN test from the per-cell loops.
// Implements the segment handling of preprocessing.py:14-26 (two-point segments) and the missing-data rule of
// observationModels.py:53-54.
__global__ void prep_steps_kernel(const double *__restrict__ data, StepC *__restrict__ out, long long T, int om, int nc,
                                  int nce) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * nce) return;
    const long long t = e / nce;
    const int c = (int)(e - t * nce);
    StepC s;
    s.d0 = 0.0;
    s.d1 = 0.0;
    s.c = 0.0;
    s.skip = 0.0;
    if (om == BLG_OM_GAUSSIAN_MEAN) {
        const double v = data[t * nc + 0], err = data[t * nc + 1];
        if (isnan(v) || isnan(err)) s.skip = 1.0;
        s.d0 = v;
        s.d1 = 1.0 / (2.0 * err * err);
        s.c = -0.5 * log(2.0 * M_PI * err * err);
    } else if (om == BLG_OM_AR1 || om == BLG_OM_SCALED_AR1) {
        s.d0 = data[t * nc + c];
        s.d1 = data[(t + 1) * nc + c];
        if (isnan(s.d0) || isnan(s.d1)) s.skip = 1.0;
    } else {
        s.d0 = data[t * nc + c];
        if (isnan(s.d0)) s.skip = 1.0;
        if (om == BLG_OM_POISSON) s.c = lgamma(s.d0 + 1.0);  // log(k!)  (observationModels.py:502)
    }
    out[e] = s;
}

// Likelihood table lik[t][g] = processedPdf(grid, segment_t) evaluated ONCE per sweep: the likelihood does not depend
// on the hyper-parameter combination (core.py:375 recomputes it per combo), so the B combos of a call share one
// table through L2 instead of each evaluating T*G exponentials.
__global__ void lik_table_kernel(DevProblem pb, const StepC *__restrict__ steps, long long T, double *__restrict__ table) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * pb.G) return;
    const long long t = e / pb.G;
    const int g = (int)(e - t * pb.G);
    const int i0 = g / pb.n1, i1 = g - i0 * pb.n1;
    LikTables tb;
    tb.A0 = pb.tabA[0];
    tb.A1 = pb.tabA[1];
    tb.A2 = pb.tabA[2];
    tb.B0 = pb.tabB[0];
    tb.B1 = pb.tabB[1];
    table[e] = lik_cell(pb, tb, steps + t * pb.ncols_eff, i0, i1);
}

// The same table in "owner order" for the warp-specialised 1-D kernels (fast1d_ws.cuh): row t holds M planes of NC
// entries, plane m entry c = likelihood of the m-th cell of compute thread c (0 beyond the grid or beyond the
// thread's cells), so the loads of a compute thread are coalesced across the warp.  Threads own M consecutive cells,
// those of the last warp (c >= NC - 32) ML <= M cells (uneven split).  `src` != NULL: permute a caller-supplied table
// (BLG_OM_TABLE) instead of evaluating.
__global__ void lik_table_perm_kernel(DevProblem pb, const StepC *__restrict__ steps, const double *__restrict__ src,
                                      long long T, int M, int ML, int NC, double *__restrict__ table) {
    const long long pitch = (long long)M * NC;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * pitch) return;
    const long long t = e / pitch;
    const int q = (int)(e - t * pitch);
    const int m = q / NC, c = q - m * NC;
    const int cs = NC - 32;  // first thread of the short warp
    const int g = c < cs ? c * M + m : (m < ML ? cs * M + (c - cs) * ML + m : pb.G);
    double v = 0.0;
    if (g < pb.G) {
        if (src) {
            v = src[t * pb.G + g];
        } else {
            LikTables tb;
            tb.A0 = pb.tabA[0];
            tb.A1 = pb.tabA[1];
            tb.A2 = pb.tabA[2];
            tb.B0 = pb.tabB[0];
            tb.B1 = pb.tabB[1];
            v = lik_cell(pb, tb, steps + t * pb.ncols_eff, g, 0);
        }
    }
    table[e] = v;
}

// out[j] = sum_k weight[k] * state[k][j]   (core.py:1410, :2195-2197, :2212); fixed summation order
__global__ void mix_kernel(const double *__restrict__ state, const double *__restrict__ weight, long long K, long long n,
                           double *__restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (long long k = 0; k < K; ++k) s = fma(__ldg(weight + k), __ldg(state + k * n + j), s);
    out[j] = s;
}

__global__ void scale_kernel(double *__restrict__ x, long long count, double f) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) x[i] *= f;
}

__global__ void weights_kernel(const double *__restrict__ logw, const int *__restrict__ alive, long long B,
                               double *__restrict__ w) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) w[b] = (!alive || alive[b] == 1) ? exp(logw[b]) : 0.0;
}

// avg[e] += sum_b w[b] * max(scale[b][t] * seq[b][e], 1e-300)   (core.py:1362-1366 applied to stored sequences;
// scale = NULL: rows are normalised already; otherwise the per-row factor left by a BLG_F_RAW_POSTERIOR backward pass)
template <bool SCALED>
__global__ void accumulate_kernel(const double *__restrict__ seq, const double *__restrict__ w, long long B,
                                  long long count, double *__restrict__ avg, const double *__restrict__ scale, long long T,
                                  int G, long long seqStride, long long rowStride) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const long long t = SCALED ? e / G : 0;
    double s = 0.0;
    long long b = 0;
    for (; b + 8 <= B; b += 8) {  // 8 independent streaming loads in flight per thread
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(seq + (b + u) * seqStride + e);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double wb = __ldg(w + b + u);
            const double p = SCALED ? v[u] * __ldg(scale + (b + u) * rowStride + t) : v[u];
            if (wb > 0.0) s = fma(wb, p < kTiny ? kTiny : p, s);  // wb == 0: combo not alive (rows may hold NaN)
        }
    }
    for (; b < B; ++b) {
        const double wb = __ldg(w + b);
        const double v = __ldcs(seq + b * seqStride + e);
        const double p = SCALED ? v * __ldg(scale + b * rowStride + t) : v;
        if (wb > 0.0) s = fma(wb, p < kTiny ? kTiny : p, s);
    }
    avg[e] += s;
}

// One block: averaging weights of a wave relative to the running reference log-weight (blg_wave_weights).
// factor[0] receives exp(old shift - new shift), the re-base of the running sum (1 when nothing moves).
__global__ void __launch_bounds__(1024) wave_weights_kernel(const double *__restrict__ logE, const double *__restrict__ logPrior,
                                                            long long B, double *__restrict__ shift,
                                                            double *__restrict__ factor, double *__restrict__ logw) {
    __shared__ double scratch[2 * kMaxWarps];
    RedScratch rs;
    rs.buf = scratch;
    rs.phase = 0;
    double top = -INFINITY;
    for (long long b = threadIdx.x; b < B; b += blockDim.x) {
        const double lw = logE[b] + logPrior[b];
        if (isfinite(lw)) top = fmax(top, lw);
    }
    top = block_max(top, rs);
    const double old = shift[0];
    const double now = fmax(old, top);  // fmax ignores nothing here: both are -inf or finite
    for (long long b = threadIdx.x; b < B; b += blockDim.x) {
        const double lw = logE[b] + logPrior[b];
        logw[b] = isfinite(lw) ? lw - now : -INFINITY;
    }
    __syncthreads();  // everybody has read shift[0]
    if (threadIdx.x == 0) {
        factor[0] = (isfinite(old) && now > old) ? exp(old - now) : 1.0;
        shift[0] = now;
    }
}

// x[i] *= f[0] with a device scalar; a factor of exactly 1 leaves the array untouched (no traffic)
__global__ void scale_dev_kernel(double *__restrict__ x, long long count, const double *__restrict__ f) {
    const double v = __ldg(f);
    if (v == 1.0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) x[i] *= v;
}

__global__ void rebase_factor_kernel(const double *__restrict__ from, const double *__restrict__ to, double *__restrict__ f) {
    const double a = from[0], b = to[0];
    f[0] = (isfinite(a) && isfinite(b) && b != a) ? exp(a - b) : 1.0;
}

// Marginal over the other axis of every row of a [T][n0][n1] sequence (core.py:915, :979-980).  One warp per output
// element group: axis 0 -> out[t][i0] = sum_j seq[t][i0][j] (contiguous line, lanes stride the line, shuffle sum);
// axis 1 -> out[t][j] = sum_i seq[t][i][j] (one thread per column, coalesced across the warp).  Fixed order.
__global__ void marginal_kernel(const double *__restrict__ seq, long long T, int n0, int n1, int axis,
                                double *__restrict__ out) {
    if (axis == 0) {
        const long long line = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // (t, i0)
        const int lane = threadIdx.x & 31;
        if (line >= T * n0) return;
        const double *p = seq + line * n1;
        double s = 0.0;
        for (int j = lane; j < n1; j += 32) s += p[j];
        s = warp_sum(s);
        if (lane == 0) out[line] = s;
    } else {
        const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (t, j)
        if (e >= T * n1) return;
        const long long t = e / n1;
        const int j = (int)(e - t * n1);
        const double *p = seq + t * (long long)n0 * n1 + j;
        double s = 0.0;
        for (int i = 0; i < n0; ++i) s += p[(long long)i * n1];
        out[e] = s;
    }
}

// out[g] = (1/T) sum_t seq[t][g]   (core.py:886)
__global__ void time_average_kernel(const double *__restrict__ seq, long long T, long long G, double *__restrict__ out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (long long t = 0; t < T; ++t) s += __ldcs(seq + t * G + g);
    out[g] = s / (double)T;
}

// Change-point prefix sharing: one CTA per (group g, row t).  The backward message ratio[g][t][.] and the likelihood
// row lik[t][.] are read once; for every change-point k of the call with cp[k] < t the row of combo k * nG + g becomes
// u = alpha * ratio in place, row_scale = 1 / sum(u), local evidence from sum(u / lik).  Up to kShareK change-points
// per launch (per-thread partial sums live in registers).
constexpr int kShareK = 8;
template <bool VEC>  // VEC: G even and 16-byte aligned rows -> double2 accesses (twice the bytes in flight per request)
__global__ void __launch_bounds__(256, 2) share_apply_kernel(double *__restrict__ seq, long long seqStride,
                                                             const double *__restrict__ ratio, long long ratioStride,
                                                             const double *__restrict__ lik, long long T, int G,
                                                             double lcProd, double *__restrict__ rowScale,
                                                             double *__restrict__ local, long long rowStride,
                                                             int *__restrict__ alive, long long nG,
                                                             const int *__restrict__ cp, int k0, int nK) {
    __shared__ double scratch[6 * kMaxWarps];
    RedScratch rs;
    rs.buf = scratch;
    rs.phase = 0;
    const long long g = blockIdx.x / T, t = blockIdx.x - g * T;
    unsigned on = 0u;  // bit k: change-point k0 + k lies before row t and its combo is alive
    for (int k = 0; k < nK; ++k)
        if (__ldg(cp + k0 + k) < t && (!alive || alive[(k0 + k) * nG + g] == 1)) on |= 1u << k;
    if (!on) return;
    const double *rr = ratio + g * ratioStride + t * G, *lk = lik + t * (long long)G;
    double *rows = seq + (k0 * nG + g) * seqStride + t * G;  // row of change-point k0; the next one is nG sequences on
    const long long kStride = nG * seqStride;
    double s[kShareK], q[kShareK];
#pragma unroll
    for (int k = 0; k < kShareK; ++k) s[k] = q[k] = 0.0;
    // core.py:463 without a division per change-point: one reciprocal of the likelihood per cell (see fast_div_pos)
    auto cell = [&](double r, double l, double &x, int k) {
        const bool small = l < 1e-290;
        const double rl = fast_rcp(small ? l * 0x1p600 : l);
        const double u = x * r;
        x = u;
        s[k] += u;
        const double qv = u * rl;
        q[k] += small ? qv * 0x1p600 : qv;
    };
    if (VEC) {
        for (int c2 = threadIdx.x; c2 < G / 2; c2 += blockDim.x) {
            const double2 r = __ldcs(reinterpret_cast<const double2 *>(rr) + c2);
            const double2 l = __ldg(reinterpret_cast<const double2 *>(lk) + c2);
            double2 v[kShareK];
#pragma unroll
            for (int k = 0; k < kShareK; ++k)
                if (on >> k & 1u) v[k] = reinterpret_cast<const double2 *>(rows + k * kStride)[c2];
#pragma unroll
            for (int k = 0; k < kShareK; ++k)
                if (on >> k & 1u) {
                    cell(r.x, l.x, v[k].x, k);
                    cell(r.y, l.y, v[k].y, k);
                    reinterpret_cast<double2 *>(rows + k * kStride)[c2] = v[k];
                }
        }
    } else {
        for (int c = threadIdx.x; c < G; c += blockDim.x) {
            const double r = __ldcs(rr + c), l = __ldg(lk + c);
#pragma unroll
            for (int k = 0; k < kShareK; ++k)
                if (on >> k & 1u) {
                    double x = rows[k * kStride + c];
                    cell(r, l, x, k);
                    rows[k * kStride + c] = x;
                }
        }
    }
#pragma unroll
    for (int k = 0; k < kShareK; ++k)
        if (on >> k & 1u) {  // uniform over the block
            double sk = s[k], qk = q[k];
            block_sum2(sk, qk, rs);
            if (threadIdx.x == 0) {
                const long long b = (k0 + k) * nG + g;
                if (!(sk > 0.0)) {  // core.py:440-452
                    if (alive) alive[b] = -1;
                } else {
                    rowScale[b * rowStride + t] = 1.0 / sk;
                    if (local) local[b * rowStride + t] = sk / (qk * lcProd);
                }
            }
        }
}

__global__ void fill_kernel(double *__restrict__ x, long long count, double value) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < count) x[e] = value;
}

// x[b * stride + t] = value for b < B, t < T (rows of a strided [B][T] array)
__global__ void fill_rows_kernel(double *__restrict__ x, long long B, long long T, long long stride, double value) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B * T) x[(e / T) * stride + (e % T)] = value;
}

// One CTA per row t: optional normalisation by the row sum (core.py:1379-1382) and posterior means
// mean_p[t] = sum_g post[t][g] * coord_p[g]  (core.py:480-483, :1416-1419)
__global__ void __launch_bounds__(256) finalize_kernel(double *__restrict__ seq, long long T, int G, int n1, int ndim,
                                                       const double *__restrict__ c0, const double *__restrict__ c1,
                                                       double *__restrict__ means, int normalize) {
    __shared__ double scratch[6 * kMaxWarps];
    RedScratch rs;
    rs.buf = scratch;
    rs.phase = 0;
    for (long long t = blockIdx.x; t < T; t += gridDim.x) {
        double *row = seq + t * (long long)G;
        double inv = 1.0;
        if (normalize) {
            double part = 0.0;
            for (int g = threadIdx.x; g < G; g += blockDim.x) part += row[g];
            inv = 1.0 / block_sum(part, rs);
        }
        double m0 = 0.0, m1 = 0.0;
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            double v = row[g];
            if (normalize) {
                v *= inv;
                row[g] = v;
            }
            const int i0 = g / n1;
            m0 = fma(v, __ldg(c0 + i0), m0);
            if (ndim == 2) m1 = fma(v, __ldg(c1 + (g - i0 * n1)), m1);
        }
        if (means) {
            block_sum2(m0, m1, rs);
            if (threadIdx.x == 0) {
                means[t] = m0;
                if (ndim == 2) means[T + t] = m1;
            }
        }
        __syncthreads();
    }
}

}  // namespace blg
