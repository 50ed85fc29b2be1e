// common.cuh -- device-side building blocks shared by the sm_100a kernels of libblgrid.so.
//
//   * observation-model likelihoods evaluated in registers from small per-axis tables (no per-step log/div)
//   * single-barrier block reductions (warp shuffles + one shared-memory exchange, double-buffered scratch)
//   * the GaussianRandomWalk operator: register-blocked sliding-window convolution with the reflect boundary of
//     scipy.ndimage.gaussian_filter1d, FP64 FMA bound (M outputs per thread, one shared-memory load per M FMAs)
//   * bulk-async (TMA) helpers: cp.async.bulk shared->global stores of alpha[t], global->shared prefetch of
//     alpha[t] for the backward pass (mbarrier completion)
//
// Reference semantics implemented here (file:line relative to the reference repository):
//   bayesloop/observationModels.py:35-56, :430-439, :502, :566-567, :635, :705-706, :767, :830-831, :892-896
//   bayesloop/transitionModels.py:96-115, :300-314, :339-360, :394-412, :450-471
//   scipy/ndimage/_filters.py:656-666, :747 (gaussian kernel, radius), NI_EXTEND_REFLECT
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/blgrid.h"

namespace blg {

constexpr int kConvM = 5;        // outputs per work item of the convolution; odd => conflict-free 64-bit LDS
constexpr int kMaxWarps = 32;
constexpr double kTiny = 1e-300; // clip of core.py:1362
// "misc" region of the resident kernels (doubles): [0,128) reduction scratch, [128,144) per-op parameters,
// [144,152) radii (16 ints), [152,184) windows (64 ints), [184,186) two mbarriers
constexpr int kMiscBarrierOffset = 184;
constexpr int kMiscPartialOffset = 192;  // fast 1-D kernels: [2][3][kMaxWarps] partial sums (192 doubles)

// ------------------------------------------------------------------------------------------------ structures
struct DevProblem {
    int ndim, n0, n1, G;
    int om_kind, seg, ncols, ncols_eff;  // ncols_eff: columns multiplied together (1 for GAUSSIAN_MEAN)
    double lc_prod;
    const double *tabA[3];  // axis-0 tables, each [n0]
    const double *tabB[2];  // axis-1 tables, each [n1]
    const double *c0, *c1;  // coordinates
};

struct DevProgram {
    int n_ops;
    int kind[BLG_MAX_OPS];
    int axis[BLG_MAX_OPS];
    int w_off[BLG_MAX_OPS];  // offset (doubles) of the weight table of op k inside the weight region
    int w_len[BLG_MAX_OPS];  // padded length of that table
    const double *param;
    const int *radius;
    const int *window;
};

// per (time step, data column) constants prepared once per call by prep_steps_kernel
struct StepC {
    double d0, d1, c, skip;  // skip != 0 -> missing data in this column's segment -> factor 1
};

struct PassArgs {
    DevProblem pb;
    DevProgram pg;
    long long T, B;
    const double *prior, *reset_base, *lik_table, *log_weight, *init_state;
    double *logE, *local, *alpha_seq, *avg, *final_state;
    double *row_scale;   // [B][T] normalising factor of each smoothed row (BLG_F_RAW_POSTERIOR) or NULL
    long long seq_stride;  // doubles between consecutive combos in alpha_seq (T * G when packed)
    const double *alpha_src;  // backward: filtering rows are read here (out-of-place smoothing) or NULL = alpha_seq
    long long src_stride;
    long long row_stride;  // doubles between consecutive combos in local / row_scale (T when packed)
    int *alive;
    const StepC *steps;  // [T][ncols_eff]
    unsigned flags;
    int Gp;              // G rounded up to an even number of doubles
    int off_stage;       // backward: offset (doubles) of the 2 staging buffers, or -1
    int off_tab;         // offset of the likelihood tables (3*n0p + 2*n1p)
    int off_w;           // offset of the weight region
    int off_misc;        // offset of reduction scratch / per-combo parameters
    int n0p, n1p;
    int use_bulk;        // 1: bulk-async copies allowed (G even, 16-byte aligned rows)
    int serpentine;      // 1: blockIdx -> combo mapping alternates direction per 148-block wave
    int num_sms;
    const int *order;    // device [B] launch order (descending cost) or NULL
    const int *sm_assign;  // device [sm_count][sm_slots] per-SM combo table or NULL
    int *sm_state;       // device [sm_count + B]: per-SM arrival counters, then per-combo claim flags (zeroed per launch)
    int sm_count, sm_slots;
    double *scratch;     // stream kernels: [gridDim.x][2][Gp] state buffers in global memory (L2 resident)
    int off_tile;        // stream kernels: offset (doubles) and size of the shared-memory convolution tile
    int tile_doubles;
    long long *trace;    // debugging: per-CTA {smid, combo, start, end} or NULL
    int halo;            // fast 1-D kernels: reflected halo cells on each side of the state (0 = generic kernels)
    int ws_part, ws_ctl; // warp-specialised 1-D kernels: offsets (doubles) of the partial sums / control block
    int ws_w2, ws_w2_len;  // ... and of the weights in the chunk layout of the short last compute warp (uneven split)
    // ... pacing of the chains that share an SM (ws_pace in fast1d_ws.cuh): device [B] steps done per combo (-1: not
    // started, INT_MAX: finished) or NULL; a chain holds back while it is more than pace_skew steps ahead of a peer
    int *ws_progress;
    int pace_every, pace_skew;
    int il_stride, il_ctl;  // interleaved 1-D kernels (fast1d_il.cuh): doubles per chain slot, offset of the control blocks
    int mma_pitch;       // DMMA 1-D kernels (fast1d_mma.cuh): doubles per swizzled state buffer (halo + tiles + halo + 8)
    long long lik_pitch; // row pitch (doubles) of lik_table: G, or M*threads for the owner-order table
    // cluster-resident 2-D kernels (cluster2d.cuh): rows per band, halo rows per side, offsets (doubles) of the state
    // buffer and the staging band, size of the state buffer
    int c2_nb, c2_h0, c2_off_x, c2_off_s, c2_x_doubles, c2_rows;
};

__device__ __forceinline__ long long combo_of_block(const PassArgs &a) {
    long long j = blockIdx.x;
    if (a.serpentine) {  // alternate direction per wave of num_sms blocks so cheap and expensive combos share an SM
        const long long S = a.num_sms, wave = j / S, pos = j - wave * S;
        if (wave & 1) {
            const long long left = a.B - wave * S;
            const long long cnt = left < S ? left : S;
            j = wave * S + (cnt - 1 - pos);
        }
    }
    return j;
}

// Combo of this CTA when the caller supplied a per-SM assignment (blg_program.sm_assign): the CTA looks up the SM it
// runs on, takes the next slot of that SM's list, and claims the combo.  CTAs that arrive after an SM's list is
// exhausted (only possible if the hardware did not co-schedule the whole grid) adopt any combo still unclaimed, so
// every combo is processed exactly once whatever the placement.  Returns -1 when there is nothing to do.
__device__ __forceinline__ long long combo_of_sm(const PassArgs &a, int *shared_slot) {
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int b = -1;
        int *counters = a.sm_state, *claimed = a.sm_state + a.sm_count;
        if ((int)smid < a.sm_count) {
            const int k = atomicAdd(&counters[smid], 1);
            if (k < a.sm_slots) {
                b = a.sm_assign[smid * a.sm_slots + k];
                if (b >= 0 && atomicCAS(&claimed[b], 0, 1) != 0) b = -1;
            } else {
                b = -2;  // late arrival: adopt an orphan
            }
        } else {
            b = -2;
        }
        if (b == -2) {
            b = -1;
            for (long long j = 0; j < a.B; ++j) {
                const int cand = a.order ? a.order[j] : (int)j;
                if (atomicCAS(&claimed[cand], 0, 1) == 0) {
                    b = cand;
                    break;
                }
            }
        }
        *shared_slot = b;
    }
    __syncthreads();
    return *shared_slot;
}

// ------------------------------------------------------------------------------------------------ reductions
struct RedScratch {
    double *buf;  // 2 * kMaxWarps doubles
    int phase;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One barrier per reduction: every warp publishes its partial, everybody sums all partials in the same order.
__device__ __forceinline__ double block_sum(double v, RedScratch &rs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double *slot = rs.buf + (rs.phase & 1) * kMaxWarps;
    rs.phase ^= 1;
    v = warp_sum(v);
    if (lane == 0) slot[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += slot[w];
    return s;
}

__device__ __forceinline__ void block_sum2(double &a, double &b, RedScratch &rs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double *slot = rs.buf + (rs.phase & 1) * kMaxWarps;
    double *slot2 = rs.buf + 2 * kMaxWarps + (rs.phase & 1) * kMaxWarps;
    rs.phase ^= 1;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
        slot[warp] = a;
        slot2[warp] = b;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < nw; ++w) {
        sa += slot[w];
        sb += slot2[w];
    }
    a = sa;
    b = sb;
}

__device__ __forceinline__ double block_max(double v, RedScratch &rs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double *slot = rs.buf + (rs.phase & 1) * kMaxWarps;
    rs.phase ^= 1;
    v = warp_max(v);
    if (lane == 0) slot[warp] = v;
    __syncthreads();
    double s = slot[0];
    for (int w = 1; w < nw; ++w) s = fmax(s, slot[w]);
    return s;
}

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void trace_begin(const PassArgs &a, long long b) {
    if (a.trace && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.trace[4 * (long long)blockIdx.x + 0] = smid;
        a.trace[4 * (long long)blockIdx.x + 1] = b;
        a.trace[4 * (long long)blockIdx.x + 2] = global_ns();
    }
}

__device__ __forceinline__ void trace_end(const PassArgs &a) {
    if (a.trace && threadIdx.x == 0) a.trace[4 * (long long)blockIdx.x + 3] = global_ns();
}

// ------------------------------------------------------------------------------------------------ reciprocal
// 1/x without the IEEE division subroutine (a serial ~60-instruction call per use): hardware seed (>= 20 bits) plus
// two Newton steps, <= 2 ulp.  Subnormal x is flushed by the seed, so callers guard tiny operands.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// a / b; exact division only when b is too small for the fast reciprocal (keeps 0/tiny = 0, x/0 = inf semantics)
__device__ __forceinline__ double fast_div(double a, double b) {
    return fabs(b) > 1e-290 ? a * fast_rcp(b) : a / b;
}

// a / b for b >= 0 with neither the division subroutine nor a branch (the unrolled per-cell epilogues keep their
// instruction-level parallelism): operands below the range of the fast reciprocal are rescaled by 2^600 (exact).
// b == 0 gives NaN; that is what 0 / 0 gives, and the callers' numerators vanish with b (post = alpha*beta carries the
// same likelihood factor, core.py:463).
__device__ __forceinline__ double fast_div_pos(double a, double b) {
    const bool small = b < 1e-290;
    const double q = a * fast_rcp(small ? b * 0x1p600 : b);
    return small ? q * 0x1p600 : q;
}

// The same with ONE Newton step on the hardware seed (rcp.approx.ftz.f64, ~2^-20): relative error ~1e-12, two dependent
// FP64 instructions fewer per cell.  For sum(post / lik), which only feeds the local evidence of the backward pass
// (core.py:463; the reference's own tests pin it to 5 decimals).
__device__ __forceinline__ double fast_rcp_pos1(double b) {  // 1 / b for b >= 0 (inf for b below ~1e-290: see the caller)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return fma(r, fma(-b, r, 1.0), r);
}

__device__ __forceinline__ double fast_div_pos1(double a, double b) {
    // the range test is an INTEGER compare of the high word (b >= 0): FP64 compares and selects would go through the FP64
    // pipe like the arithmetic; operands below the range of the hardware seed take the rescaled path (rare, divergent)
    if (__double2hiint(b) >= 0x03d00000) return a * fast_rcp_pos1(b);
    return a * fast_rcp_pos1(b * 0x1p600) * 0x1p600;
}

// ------------------------------------------------------------------------------------------------ log-evidence
// logE = sum_t log(norm_t) (core.py:403) without a log() on the per-step critical path: the product of the norms is
// carried as mantissa * 2^exponent (two frexp per step, ~10 instructions) and one log() is taken at the end.
struct LogProduct {
    double mant;
    long long expo;
    __device__ __forceinline__ void init() {
        mant = 1.0;
        expo = 0;
    }
    __device__ __forceinline__ void mul(double x) {
        int e1, e2;
        const double m = frexp(x, &e1);
        mant = frexp(mant * m, &e2);
        expo += e1 + e2;
    }
    __device__ __forceinline__ double log_value() const { return log(mant) + (double)expo * 0.693147180559945309417232121458; }
};

// ------------------------------------------------------------------------------------------------ likelihood
// Likelihood of ONE data column at grid point (i0, i1).  Tables (shared memory):
//   POISSON        A0 = lambda, A1 = log(lambda)                      arg = k*A1 - A0 - lgamma(k+1)
//   GAUSSIAN       A0 = mean;  B0 = 1/(2 s^2), B1 = -0.5 log(2 pi s^2) arg = -(d-A0)^2 B0 + B1
//   SCALED_AR1     A0 = r, A1 = 1/(1-r^2), A2 = -0.5 log(1-r^2); B0, B1 as above
//                                                                     arg = -(d1 - r d0)^2 A1 B0 + A2 + B1
//   AR1            A0 = r; B0, B1                                      arg = -(d1 - r d0)^2 B0 + B1
//   WHITE_NOISE    A0 = 1/(2 s^2), A1 = -0.5 log(2 pi s^2)             arg = -d^2 A0 + A1
//   GAUSSIAN_MEAN  A0 = mean; per step d1 = 1/(2 e^2), c = -0.5 log(2 pi e^2)   arg = -(d0-A0)^2 d1 + c
//   LAPLACE        A0 = mean; B0 = 1/b, B1 = -log(2b)                  arg = -|d-A0| B0 + B1
//   BERNOULLI      A0 = p clipped to [0,1] (0 outside)                 lik = d ? A0 : 1 - A0
struct LikTables {
    const double *A0, *A1, *A2, *B0, *B1;
};

__device__ __forceinline__ double lik_column(int om, const LikTables &tb, int i0, int i1, const StepC &sc) {
    switch (om) {
        case BLG_OM_POISSON:
            return exp(fma(sc.d0, tb.A1[i0], -tb.A0[i0]) - sc.c);
        case BLG_OM_GAUSSIAN: {
            const double r = sc.d0 - tb.A0[i0];
            return exp(fma(-r * r, tb.B0[i1], tb.B1[i1]));
        }
        case BLG_OM_SCALED_AR1: {
            const double r = fma(-tb.A0[i0], sc.d0, sc.d1);
            return exp(fma(-r * r, tb.A1[i0] * tb.B0[i1], tb.A2[i0] + tb.B1[i1]));
        }
        case BLG_OM_AR1: {
            const double r = fma(-tb.A0[i0], sc.d0, sc.d1);
            return exp(fma(-r * r, tb.B0[i1], tb.B1[i1]));
        }
        case BLG_OM_WHITE_NOISE:
            return exp(fma(-sc.d0 * sc.d0, tb.A0[i0], tb.A1[i0]));
        case BLG_OM_GAUSSIAN_MEAN: {
            const double r = sc.d0 - tb.A0[i0];
            return exp(fma(-r * r, sc.d1, sc.c));
        }
        case BLG_OM_LAPLACE:
            return exp(fma(-fabs(sc.d0 - tb.A0[i0]), tb.B0[i1], tb.B1[i1]));
        case BLG_OM_BERNOULLI:
            return sc.d0 != 0.0 ? tb.A0[i0] : 1.0 - tb.A0[i0];
        default:
            return 1.0;
    }
}

// processedPdf (observationModels.py:35-56): product over data columns, missing data -> ones
__device__ __forceinline__ double lik_cell(const DevProblem &pb, const LikTables &tb, const StepC *sc, int i0, int i1) {
    double lik = 1.0;
    for (int c = 0; c < pb.ncols_eff; ++c) {
        const StepC s = sc[c];
        if (s.skip == 0.0) lik *= lik_column(pb.om_kind, tb, i0, i1, s);
    }
    return lik;
}

// ------------------------------------------------------------------------------------------------ convolution
__device__ __forceinline__ int reflect_any(int i, int n) {  // NI_EXTEND_REFLECT for any index
    if ((unsigned)i < (unsigned)n) return i;
    const int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m >= n ? p - 1 - m : m;
}

__device__ __forceinline__ int reflect_once(int i, int n) {  // valid while -n <= i < 2n
    const int lo = -1 - i, hi = 2 * n - 1 - i;
    int r = i < 0 ? lo : i;
    return i >= n ? hi : r;
}

// dst = correlate1d(src, W) along one axis with the reflect boundary.  `taps` = 2R+1; W is zero-padded to a multiple
// of M (+M).  Work item = M consecutive outputs of one line; the M+2R inputs stream through a register window, so
// each tap costs one shared-memory load and M FMAs.  Lane->item mapping keeps 64-bit accesses conflict-free:
// consecutive segments of a line when the axis is contiguous (stride M, M odd), consecutive lines otherwise.
struct EpiIdentity {
    __device__ __forceinline__ double operator()(int, int, double v) const { return v; }
};

// `epi(line, index, value)` is applied to every output before it is stored (likelihood multiply + partial sums of
// the 2-D stream kernels).
template <int M, bool SIMPLE, typename Epi = EpiIdentity>
__device__ __forceinline__ void conv_lines(const double *__restrict__ src, double *__restrict__ dst,
                                           const double *__restrict__ W, int R, int n, int elemStride, int nLines,
                                           int lineStride, int dElemStride = -1, int dLineStride = -1, Epi epi = Epi()) {
    if (dElemStride < 0) {
        dElemStride = elemStride;
        dLineStride = lineStride;
    }
    const int S = (n + M - 1) / M;
    const int nItems = S * nLines;
    const int taps = 2 * R + 1;
    for (int w = threadIdx.x; w < nItems; w += blockDim.x) {
        int l, s;
        if (elemStride == 1) {
            l = w / S;
            s = w - l * S;
        } else {
            s = w / nLines;
            l = w - s * nLines;
        }
        const double *line = src + (size_t)l * lineStride;
        const int i0 = s * M;
        int idx = i0 - R;
        double win[M], acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int q = SIMPLE ? reflect_once(idx + m, n) : reflect_any(idx + m, n);
            win[m] = line[(size_t)q * elemStride];
            acc[m] = 0.0;
        }
        idx += M;
        for (int j0 = 0; j0 < taps; j0 += M) {
#pragma unroll
            for (int u = 0; u < M; ++u) {
                const double wt = W[j0 + u];
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = fma(wt, win[(u + m) % M], acc[m]);
                const int q = SIMPLE ? reflect_once(idx, n) : reflect_any(idx, n);
                win[u] = line[(size_t)q * elemStride];
                ++idx;
            }
        }
        double *out = dst + (size_t)l * dLineStride;
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (i0 + m < n) out[(size_t)(i0 + m) * dElemStride] = epi(l, i0 + m, acc[m]);
    }
}

// Gaussian weights of scipy.ndimage._filters._gaussian_kernel1d: exp(-0.5/sigma^2 * x^2), x = -R..R, normalised
// to sum 1.  Block-cooperative; leaves W[j] = 0 for j > 2R (padding read by conv_lines).
__device__ __forceinline__ void build_weights(double *W, int len, double sigma, int R, RedScratch &rs) {
    const double h = -0.5 / (sigma * sigma);
    double part = 0.0;
    for (int j = threadIdx.x; j < len; j += blockDim.x) {
        double v = 0.0;
        if (j <= 2 * R) {
            const double x = (double)(j - R);
            v = exp(h * x * x);
        }
        W[j] = v;
        part += v;
    }
    const double total = block_sum(part, rs);
    const double inv = 1.0 / total;
    for (int j = threadIdx.x; j < len; j += blockDim.x) W[j] *= inv;
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ bulk async (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

}  // namespace blg
