// fast1d_il.cuh -- K1i/K2i: the DMMA 1-D kernels (fast1d_mma.cuh) with the chains of an SM INTERLEAVED inside one CTA.
//
// Why (B200, profiles/r2_dmma_variants.txt): with one CTA per chain and 3-4 chains per SM the FP64 matrix pipe is ~70 %
// busy whatever the instruction stream looks like -- every chain is a serial loop convolution -> epilogue -> CTA
// barrier, its 4 compute warps sit on 4 different sub-partitions, and the hardware interleaves the 3-4 resident chains
// as it likes: their convolutions bunch up (40 % of the time none of the warps of a sub-partition is inside its
// convolution, per-step event trace r2E), the heaviest chain of an SM is served like the light ones and then runs
// alone.  Here ONE CTA per SM (16 compute warps + 1 service warp) owns the SM's list of up to 4 chains and every
// compute warp works through the chains IN PROGRAM ORDER, one tile of 64 cells per chain and step:
//     for t:  for chain c of the list:  wait until all warps have finished step t-1 of c;
//                                       convolution of my tile (DMMA) -> x likelihood -> state, row, partial sum;
//                                       signal "my part of step t of c is done"
//   * all chains of an SM advance in lock step by construction (no tail with a single resident chain);
//   * the latency of a chain's step is hidden by the other chains in the SAME instruction stream: while the last warps
//     finish chain c the first ones are already inside chain c+1 -- there is no CTA-wide barrier, only a counter per
//     chain (shared-memory atomic + fences), and a warp returns to chain c a whole round later;
//   * the four warps of a sub-partition start a quarter of a round apart (clock spin at the start), so that their
//     epilogues fall into each other's convolutions; the skew persists because all warps do the same work;
//   * the likelihood row of a step is loaded ONCE per warp and reused for all chains (they are at the same step).
// The service warp follows one step behind: per chain it adds the 16 per-warp sums, keeps the evidence bookkeeping of
// fast1d_mma.cuh (power-of-two rescale on demand, telescoped log-evidence, deferred divisions) and, in the backward
// pass, runs the alpha ring (bulk-async loads / stores).  Semantics: core.py:372-417, :424-470; transitionModels.py:96-118.
#pragma once

#include "fast1d_mma.cuh"

namespace blg {

constexpr int kIlChains = 4;                   // chains interleaved per CTA (= blg_program.sm_slots)
constexpr int kIlCW = 16;                      // compute warps
constexpr int kIlNT = (kIlCW + 1) * 32;        // + the service warp
constexpr int kIlCtlDoubles = 8;               // control block per chain, see IlCtl
constexpr int kIlPL = kIlCW * 16;               // partial sums of one kind per chain-step: 16 per compute warp (lane pairs)
constexpr int kIlPP = 2 * 3 * kIlPL;           // partial sums per chain: [2 parities][3 sums][kIlPL]

// control block of one chain (shared memory, 8 doubles)
struct IlCtl {
    double kappa[2];       // rescale factor by step parity (written by the service warp)
    int dead, svc_done;    // step at which the service warp found a zero norm (-1: alive); steps it has processed
    int done;              // compute warps that have finished a step, summed over the steps (16 per step)
    int R;                 // radius, -1: slot unused / radius beyond the weight buffer
    int lo, hi;            // active window of the transition (time-step indices)
    int combo, pad;
    double sigma, pad2;
};
static_assert(sizeof(IlCtl) == kIlCtlDoubles * sizeof(double), "IlCtl layout");

__device__ __forceinline__ int ld_volatile_s32(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

// every lane of a warp waits until *p >= want (written by other warps of the CTA), then orders its later reads
__device__ __forceinline__ void il_wait_ge(const int *p, int want) {
    for (unsigned spins = 0; ld_volatile_s32(p) < want; ++spins)
        if (spins > (1u << 28)) asm volatile("trap;");  // seconds: a protocol error must end as an error, never as a hang
    __threadfence_block();
}

// the warp publishes "my part of this step is done": one arrival on the chain's mbarrier (the other compute warps wait
// on it with mbarrier.try_wait, which suspends the warp instead of spinning in the issue slots of the sub-partition) and
// one count for the service warp (which may lag more than one phase behind, where a parity wait would be ambiguous)
__device__ __forceinline__ void il_signal(int *p, uint64_t *bar, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        atomicAdd(p, 1);
    }
}

// the compute warps leave 16 partial sums each (one shuffle level); the service warp adds the 256 of a chain-step
// (four independent chains of adds, then the warp reduction)
__device__ __forceinline__ void il_partial(double *pp, double v, int warp, int lane) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 16) pp[warp * 16 + lane] = v;
}

__device__ __forceinline__ double il_reduce(const double *pp, int lane) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int j = 0; j < kIlPL / 32; j += 4) {
        s0 += pp[j * 32 + lane];
        s1 += pp[(j + 1) * 32 + lane];
        s2 += pp[(j + 2) * 32 + lane];
        s3 += pp[(j + 3) * 32 + lane];
    }
    return warp_sum((s0 + s1) + (s2 + s3));
}

template <int TPW>
__device__ __forceinline__ void il_conv(const double *__restrict__ line, const int (&base)[TPW], int ntw, int R,
                                        const double *__restrict__ wz, double2 (&acc)[TPW]) {
    // one accumulator per tile: consecutive matrix instructions of a warp depend on each other, the other three warps of
    // the sub-partition fill the pipe meanwhile (a split into even / odd accumulators costs a dependent FP64 add in the
    // epilogue, which is what this kernel cannot afford: every dependent FP64 level queues behind the matrix pipe)
#pragma unroll
    for (int k = 0; k < TPW; ++k) acc[k] = make_double2(0.0, 0.0);
    const int s0 = -(R + (R & 1)), groups = (2 * R + 15 + (R & 1)) >> 3;  // see mma_conv_body
    const double *x0 = line + base[0] + s0;
    const double *w = wz + s0;
#pragma unroll 4
    for (int j = 0; j < groups; ++j) {
        const double b0 = w[0], b1 = w[1];
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            if (k < ntw) {
                const double2 av = *reinterpret_cast<const double2 *>(x0 + (base[k] - base[0]));
                dmma884(acc[k].x, acc[k].y, av.x, b0);
                dmma884(acc[k].x, acc[k].y, av.y, b1);
            }
        }
        x0 += 8;
        w += 8;
    }
}

// common set-up of both passes: the CTA's chain list, zeroed state buffers, weights; returns the number of list entries
__device__ __forceinline__ int il_setup(const PassArgs &a, double *sm, bool backward) {
    IlCtl *ctl = reinterpret_cast<IlCtl *>(sm + a.il_ctl);
    const int slots = a.sm_slots < kIlChains ? a.sm_slots : kIlChains;
    if (threadIdx.x < kIlChains) {
        const int c = threadIdx.x;
        IlCtl &q = ctl[c];
        const int b = (c < slots && a.sm_assign) ? a.sm_assign[(long long)blockIdx.x * a.sm_slots + c] : -1;
        q.combo = b;
        q.dead = -1;
        q.svc_done = 0;
        q.done = 0;
        q.kappa[0] = q.kappa[1] = 1.0;
        q.R = -1;
        q.sigma = 0.0;
        q.lo = q.hi = 0;
        if (b >= 0) {
            const double sigma = a.pg.param[b];
            int R = a.pg.radius[b];
            if (!(sigma > 0.0) || R <= 0) R = 0;
            q.sigma = sigma;
            q.lo = a.pg.window[b * 4 + (backward ? 2 : 0)];
            q.hi = a.pg.window[b * 4 + (backward ? 3 : 1)];
            if (2 * R + 1 + 2 * kMmaWPad > a.pg.w_len[0] || mma_halo(R) > a.halo) {  // radius beyond blg_program.max_radius
                if (!backward) {
                    a.logE[b] = NAN;
                    if (a.alive) a.alive[b] = -2;
                }
            } else if (!backward || !a.alive || a.alive[b] == 1) {  // backward: the forward pass aborted (core.py:400)
                q.R = R;
            }
        }
    }
    __syncthreads();
    int n = 0;
    for (int c = 0; c < kIlChains; ++c)
        if (ctl[c].combo >= 0) n = c + 1;
    RedScratch rs;
    rs.buf = sm + a.off_misc;
    rs.phase = 0;
    for (int c = 0; c < n; ++c) {
        double *slot = sm + (long long)c * a.il_stride;
        for (int j = threadIdx.x; j < 2 * a.mma_pitch; j += kIlNT) slot[j] = 0.0;
        if (ctl[c].R >= 0) mma_build_weights(slot + a.off_w, a.pg.w_len[0], ctl[c].sigma, ctl[c].R, rs);  // CTA barriers inside
    }
    for (int j = threadIdx.x; j < kIlChains * kIlPP; j += kIlNT) (sm + a.ws_part)[j] = 0.0;
    __syncthreads();
    return n;
}

// ------------------------------------------------------------------------------------------------ K1i forward
template <int TPW>
__global__ void __launch_bounds__(kIlNT, 1) fwd_fast1d_il_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const int n = pb.G, halo = a.halo;
    const long long T = a.T;
    const int nc = il_setup(a, sm, false);
    if (nc == 0) return;
    IlCtl *ctl = reinterpret_cast<IlCtl *>(sm + a.il_ctl);
    double *PPall = sm + a.ws_part;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    const bool first = (a.flags & BLG_F_TRANSITION_FIRST) != 0;
    const bool rawRows = store && (a.flags & BLG_F_RAW_ALPHA);
    uint64_t *ready = reinterpret_cast<uint64_t *>(sm + a.ws_ctl) + 2 * kIlChains;  // one mbarrier per chain: 16 arrivals
    if (threadIdx.x == 0) {
        for (int c = 0; c < kIlChains; ++c) mbar_init(&ready[c], kIlCW);
        fence_proxy_async();
    }
    // initial state of every chain (core.py:363)
    for (int c = 0; c < nc; ++c) {
        const int b = ctl[c].combo;
        if (ctl[c].R < 0) continue;
        double *buf0 = sm + (long long)c * a.il_stride + halo;
        const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)n : a.prior;
        for (int g = threadIdx.x; g < n; g += kIlNT) {
            const double v = init[g];
            buf0[g] = v;
            if (g < halo) buf0[-1 - g] = v;
            if (g >= n - halo) buf0[2 * n - 1 - g] = v;
        }
    }
    __syncthreads();

    if (warp < kIlCW) {
        // ------------------------------------------------------------------ compute warps
        const int g8 = lane >> 2, u = lane & 3;
        int base[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) base[k] = (warp + k * kIlCW) * 64 + 8 * g8 + 2 * u;
        const int ntiles = (n + 63) >> 6;
        int ntw = (ntiles - warp + kIlCW - 1) / kIlCW;
        ntw = ntw < 0 ? 0 : (ntw > TPW ? TPW : ntw);
        bool live[kIlChains];
        int est = 0;  // rough cycles of one round of this CTA (for the stagger below)
#pragma unroll
        for (int c = 0; c < kIlChains; ++c) {
            live[c] = c < nc && ctl[c].R >= 0;
            if (live[c]) est += ((2 * ctl[c].R + 15 + (ctl[c].R & 1)) >> 3) * 128 * TPW + 500;
        }
        {   // the four warps of a sub-partition (warp mod 4) start a quarter of a round apart
            const long long until = clock64() + (long long)(warp >> 2) * (est >> 2);
            while (clock64() < until) {
            }
        }
        const double *likp = a.lik_table;
        const long long pitch = a.lik_pitch;
        double2 lk[TPW], lkn[TPW];  // likelihood of the lane's cell pairs (the same row for every chain of the CTA)
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            lk[k] = base[k] < n ? __ldg(reinterpret_cast<const double2 *>(likp + base[k])) : make_double2(0.0, 0.0);
            lkn[k] = lk[k];
        }
        for (long long t = 0; t < T; ++t) {
            if (t + 1 < T) {
#pragma unroll
                for (int k = 0; k < TPW; ++k)
                    if (base[k] < n) lkn[k] = __ldg(reinterpret_cast<const double2 *>(likp + (t + 1) * pitch + base[k]));
            }
#pragma unroll
            for (int c = 0; c < kIlChains; ++c) {
                if (!live[c]) continue;
                IlCtl &q = ctl[c];
                if (t >= 1) mbar_wait(&ready[c], (uint32_t)((t - 1) & 1));  // every warp has finished step t-1 of this chain
                if (t >= 2) il_wait_ge(&q.svc_done, (int)t - 1);            // ... and the service warp step t-2
                {
                    const int ds = ld_volatile_s32(&q.dead);
                    if (ds >= 0 && ds <= t - 2) {  // deaths up to step t-2 are visible to every warp at this point
                        live[c] = false;
                        continue;
                    }
                }
                double *slot = sm + (long long)c * a.il_stride;
                const double *cur = slot + (t & 1) * a.mma_pitch + halo;
                double *nxt = slot + ((t + 1) & 1) * a.mma_pitch + halo;
                const int R = q.R;
                const bool trans = (t > 0 || first) && (t - 1 >= q.lo) && (t - 1 < q.hi);
                double2 v[TPW];
                if (trans && R > 0) {
                    il_conv<TPW>(cur, base, ntw, R, slot + a.off_w + (2 * u - g8 + R + kMmaWPad), v);  // transitionModels.py:111
                } else {
#pragma unroll
                    for (int k = 0; k < TPW; ++k)
                        v[k] = base[k] < n ? *reinterpret_cast<const double2 *>(cur + base[k]) : make_double2(0.0, 0.0);
                }
                const double kappa = t >= 2 ? *reinterpret_cast<volatile double *>(&q.kappa[t & 1]) : 1.0;
                // alpha <- prior * likelihood (core.py:375-382); the factor is a power of two, 1 on most steps
                if (__double2hiint(kappa) == 0x3ff00000) {
#pragma unroll
                    for (int k = 0; k < TPW; ++k) {
                        v[k].x *= lk[k].x;
                        v[k].y *= lk[k].y;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < TPW; ++k) {
                        v[k].x *= kappa * lk[k].x;
                        v[k].y *= kappa * lk[k].y;
                    }
                }
                double part = 0.0;
                double *row = rawRows ? a.alpha_seq + q.combo * a.seq_stride + t * (long long)n : nullptr;
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    if (base[k] < n) {
                        mma_store_pair(nxt, base[k], n, halo, v[k]);
                        if (rawRows) __stcs(reinterpret_cast<double2 *>(row + base[k]), v[k]);
                        part += v[k].x + v[k].y;
                    }
                }
                il_partial(PPall + c * kIlPP + (int)(t & 1) * 3 * kIlPL, part, warp, lane);  // reduced by the service warp
                il_signal(&q.done, &ready[c], lane);
            }
#pragma unroll
            for (int k = 0; k < TPW; ++k) lk[k] = lkn[k];
        }
    } else {
        // ------------------------------------------------------------------ service warp: one step behind, chain by chain
        bool live[kIlChains], dead[kIlChains];
        double sPrev[kIlChains], sLast[kIlChains], myS[kIlChains], myPrev[kIlChains];
        int keNow[kIlChains], keNext[kIlChains], hold[kIlChains], myKe[kIlChains];
        long long keSum[kIlChains];
#pragma unroll
        for (int c = 0; c < kIlChains; ++c) {
            live[c] = c < nc && ctl[c].R >= 0;
            dead[c] = false;
            sPrev[c] = sLast[c] = myS[c] = myPrev[c] = 1.0;
            keNow[c] = keNext[c] = hold[c] = myKe[c] = 0;
            keSum[c] = 0;
        }
        for (long long t = 0; t < T; ++t) {
#pragma unroll
            for (int c = 0; c < kIlChains; ++c) {
                if (!live[c]) continue;
                IlCtl &q = ctl[c];
                const long long b = q.combo;
                double *const local = a.local ? a.local + b * a.row_stride : nullptr;
                // local evidence of the parked steps [t0, t0 + count): norm_t = s_t / (k_t s_{t-1})   core.py:385, :404
                auto flush = [&](long long t0, int count) {
                    if (local && lane < count) local[t0 + lane] = fast_div(myS[c], times_pow2(myPrev[c], myKe[c])) * pb.lc_prod;
                };
                il_wait_ge(&q.done, kIlCW * ((int)t + 1));
                const double st_sum = il_reduce(PPall + c * kIlPP + (int)(t & 1) * 3 * kIlPL, lane);
                int keAfter;
                const double kAfter = ondemand_scale(st_sum, hold[c], keAfter);  // k_{t+2}
                if (lane == 0) *reinterpret_cast<volatile double *>(&q.kappa[t & 1]) = kAfter;
                if (!(st_sum > 0.0) || isinf(st_sum)) {  // core.py:388-400
                    dead[c] = true;
                    live[c] = false;
                    flush(t & ~31LL, (int)(t & 31));
                    __syncwarp();
                    if (lane == 0) {
                        *reinterpret_cast<volatile int *>(&q.dead) = (int)t;
                        __threadfence_block();
                        *reinterpret_cast<volatile int *>(&q.svc_done) = 0x7fffffff;
                    }
                    continue;
                }
                keSum[c] += keNow[c];
                if (lane == (int)(t & 31)) {
                    myS[c] = st_sum;
                    myPrev[c] = sPrev[c];
                    myKe[c] = keNow[c];
                }
                if ((t & 31) == 31 || t == T - 1) flush(t & ~31LL, (int)(t & 31) + 1);
                const double *st = sm + (long long)c * a.il_stride + ((t + 1) & 1) * a.mma_pitch + halo;  // just filled
                if (store && !rawRows) {  // core.py:389, :408 -- normalised filtering distribution
                    const double inv = fast_rcp(st_sum);
                    double *row = a.alpha_seq + b * a.seq_stride + t * (long long)n;
                    for (int j = 2 * lane; j < n; j += 64) {
                        double2 x = *reinterpret_cast<const double2 *>(st + j);
                        x.x *= inv;
                        x.y *= inv;
                        __stcs(reinterpret_cast<double2 *>(row + j), x);
                    }
                }
                if ((a.flags & BLG_F_SAVE_STATE) && a.final_state && t == T - 1) {
                    const double inv = fast_rcp(st_sum);
                    double *fs = a.final_state + b * (long long)n;
                    for (int j = lane; j < n; j += 32) fs[j] = st[j] * inv;
                }
                sPrev[c] = st_sum;
                sLast[c] = st_sum;
                keNow[c] = keNext[c];
                keNext[c] = keAfter;
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    *reinterpret_cast<volatile int *>(&q.svc_done) = (int)t + 1;
                }
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < kIlChains; ++c) {
                if (c >= nc || ctl[c].R < 0) continue;
                const long long b = ctl[c].combo;
                // sum_t log(norm_t) = log(s_{T-1}) - sum_t log(k_t)   (core.py:403)
                double logE = log(sLast[c]) - (double)keSum[c] * 0.693147180559945309417232121458;
                if (dead[c])
                    logE = -INFINITY;
                else if (!(a.flags & BLG_F_INIT_STATE))
                    logE += log(pb.lc_prod);  // core.py:417
                a.logE[b] = logE;
                if (a.alive) a.alive[b] = dead[c] ? 0 : 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K2i backward
// Steps are counted from the end of the series: step s works on row i = T-1-s.
template <int TPW>
__global__ void __launch_bounds__(kIlNT, 1) bwd_fast1d_il_kernel(const PassArgs a) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const int n = pb.G, halo = a.halo, Gp = a.Gp;
    const long long T = a.T;
    const int nc = il_setup(a, sm, true);
    if (nc == 0) return;
    IlCtl *ctl = reinterpret_cast<IlCtl *>(sm + a.il_ctl);
    double *PPall = sm + a.ws_part;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + a.ws_ctl);  // [chain][ring slot]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool rawRows = a.row_scale != nullptr;  // BLG_F_RAW_POSTERIOR: rows leave unnormalised + their factor
    const uint32_t rowBytes = (uint32_t)(n * sizeof(double));
    uint64_t *ready = bars + 2 * kIlChains;  // one mbarrier per chain: 16 arrivals per step
    if (threadIdx.x == 0) {
        for (int j = 0; j < 2 * kIlChains; ++j) mbar_init(&bars[j], 1);
        for (int c = 0; c < kIlChains; ++c) mbar_init(&ready[c], kIlCW);
        fence_proxy_async();
    }
    __syncthreads();

    if (warp < kIlCW) {
        // ------------------------------------------------------------------ compute warps
        const int g8 = lane >> 2, u = lane & 3;
        int base[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) base[k] = (warp + k * kIlCW) * 64 + 8 * g8 + 2 * u;
        const int ntiles = (n + 63) >> 6;
        int ntw = (ntiles - warp + kIlCW - 1) / kIlCW;
        ntw = ntw < 0 ? 0 : (ntw > TPW ? TPW : ntw);
        bool live[kIlChains];
        uint32_t phases[kIlChains];  // bit s = parity of the next completion of ring slot s of the chain
        int est = 0;
#pragma unroll
        for (int c = 0; c < kIlChains; ++c) {
            live[c] = c < nc && ctl[c].R >= 0;
            phases[c] = 0u;
            if (live[c]) est += ((2 * ctl[c].R + 15 + (ctl[c].R & 1)) >> 3) * 128 * TPW + 700;
        }
        {   // the four warps of a sub-partition (warp mod 4) start a quarter of a round apart
            const long long until = clock64() + (long long)(warp >> 2) * (est >> 2);
            while (clock64() < until) {
            }
        }
        const double *likp = a.lik_table;
        const long long pitch = a.lik_pitch;
        double2 lk[TPW], lkn[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            lk[k] = base[k] < n ? __ldg(reinterpret_cast<const double2 *>(likp + (T - 1) * pitch + base[k])) : make_double2(1.0, 1.0);
            lkn[k] = lk[k];
        }
        for (long long s = 0; s < T; ++s) {
            const long long i = T - 1 - s;
            const int sb = (int)(i & 1);
            if (i > 0) {
#pragma unroll
                for (int k = 0; k < TPW; ++k)
                    if (base[k] < n) lkn[k] = __ldg(reinterpret_cast<const double2 *>(likp + (i - 1) * pitch + base[k]));
            }
#pragma unroll
            for (int c = 0; c < kIlChains; ++c) {
                if (!live[c]) continue;
                IlCtl &q = ctl[c];
                if (s >= 1) mbar_wait(&ready[c], (uint32_t)((s - 1) & 1));
                if (s >= 2) il_wait_ge(&q.svc_done, (int)s - 1);
                {
                    const int ds = ld_volatile_s32(&q.dead);
                    if (ds >= 0 && ds <= s - 2) {
                        live[c] = false;
                        continue;
                    }
                }
                double *slot = sm + (long long)c * a.il_stride;
                const double *cur = slot + (s & 1) * a.mma_pitch + halo;
                double *nxt = slot + ((s + 1) & 1) * a.mma_pitch + halo;
                const int R = q.R;
                double2 beta[TPW];
                if (s == 0) {
#pragma unroll
                    for (int k = 0; k < TPW; ++k) {
                        const double v = base[k] < n ? 1.0 / (double)n : 0.0;  // core.py:424-425
                        beta[k] = make_double2(v, v);
                    }
                } else {
                    const bool trans = (i + 1 >= q.lo) && (i + 1 < q.hi);
                    if (trans && R > 0) {
                        il_conv<TPW>(cur, base, ntw, R, slot + a.off_w + (2 * u - g8 + R + kMmaWPad), beta);  // transitionModels.py:117-118
                    } else {
#pragma unroll
                        for (int k = 0; k < TPW; ++k)
                            beta[k] = base[k] < n ? *reinterpret_cast<const double2 *>(cur + base[k]) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int k = 0; k < TPW; ++k)
                        if (base[k] >= n) beta[k] = make_double2(0.0, 0.0);
                }
                const double kb = s >= 2 ? *reinterpret_cast<volatile double *>(&q.kappa[s & 1]) : 1.0;
                const bool unit = __double2hiint(kb) == 0x3ff00000;
                mbar_wait(&bars[2 * c + sb], (phases[c] >> sb) & 1u);  // alpha[i] has arrived in the ring slot
                phases[c] ^= 1u << sb;
                double *A = slot + a.off_stage + sb * Gp;
                double spu = 0.0, sst = 0.0, sql = 0.0;
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    if (base[k] < n) {
                        const double2 al = *reinterpret_cast<const double2 *>(A + base[k]);
                        double2 pu, st;
                        pu.x = al.x * beta[k].x;  // posterior ~ alpha*beta   core.py:436
                        pu.y = al.y * beta[k].y;
                        sql += fast_div_pos1(pu.x, lk[k].x) + fast_div_pos1(pu.y, lk[k].y);  // core.py:463
                        st.x = beta[k].x * lk[k].x;  // beta*likelihood          core.py:467
                        st.y = beta[k].y * lk[k].y;
                        if (!unit) {
                            st.x *= kb;
                            st.y *= kb;
                        }
                        *reinterpret_cast<double2 *>(A + base[k]) = pu;
                        mma_store_pair(nxt, base[k], n, halo, st);
                        spu += pu.x + pu.y;
                        sst += st.x + st.y;
                    }
                }
                {
                    double *pp = PPall + c * kIlPP + (int)(s & 1) * 3 * kIlPL;
                    il_partial(pp, spu, warp, lane);
                    il_partial(pp + kIlPL, sst, warp, lane);  // sum of the new state (magnitude control only)
                    il_partial(pp + 2 * kIlPL, sql, warp, lane);
                }
                fence_proxy_async();  // alpha * beta in the ring slot is read by the bulk-async row store
                il_signal(&q.done, &ready[c], lane);
            }
#pragma unroll
            for (int k = 0; k < TPW; ++k) lk[k] = lkn[k];
        }
    } else {
        // ------------------------------------------------------------------ service warp
        bool live[kIlChains], dead[kIlChains];
        double mySpu[kIlChains], mySql[kIlChains];
        int hold[kIlChains];
#pragma unroll
        for (int c = 0; c < kIlChains; ++c) {
            live[c] = c < nc && ctl[c].R >= 0;
            dead[c] = false;
            mySpu[c] = mySql[c] = 1.0;
            hold[c] = 0;
            if (live[c] && lane == 0) {
                const long long b = ctl[c].combo;
                const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : a.alpha_seq + b * a.seq_stride;
                double *S0 = sm + (long long)c * a.il_stride + a.off_stage;
                bulk_load(S0 + ((T - 1) & 1) * Gp, src + (T - 1) * (long long)n, rowBytes, &bars[2 * c + ((T - 1) & 1)]);
                if (T >= 2) bulk_load(S0 + ((T - 2) & 1) * Gp, src + (T - 2) * (long long)n, rowBytes, &bars[2 * c + ((T - 2) & 1)]);
            }
        }
        for (long long s = 0; s < T; ++s) {
            const long long i = T - 1 - s;
            const int sb = (int)(i & 1);
#pragma unroll
            for (int c = 0; c < kIlChains; ++c) {
                if (!live[c]) continue;
                IlCtl &q = ctl[c];
                const long long b = q.combo;
                double *const local = a.local ? a.local + b * a.row_stride : nullptr;
                double *const rscale = rawRows ? a.row_scale + b * a.row_stride : nullptr;
                // rows [iLow, iLow + count): lane j holds row iLow + count - 1 - j
                auto flush = [&](long long iLow, int count) {
                    if (lane < count) {
                        const long long row = iLow + count - 1 - lane;
                        if (rscale) rscale[row] = fast_rcp(mySpu[c]);  // posterior = alpha*beta / sum(alpha*beta)   core.py:439-441
                        if (local) local[row] = fast_div(mySpu[c], mySql[c] * pb.lc_prod);  // 1/(sum(post/lik)*lc)  core.py:463
                    }
                };
                il_wait_ge(&q.done, kIlCW * ((int)s + 1));
                const double *pp = PPall + c * kIlPP + (int)(s & 1) * 3 * kIlPL;
                const double spu = il_reduce(pp, lane);
                const double sstate = il_reduce(pp + kIlPL, lane);
                const double sql = il_reduce(pp + 2 * kIlPL, lane);
                int ke;
                const double kAfter = ondemand_scale(sstate, hold[c], ke);
                if (lane == 0) *reinterpret_cast<volatile double *>(&q.kappa[s & 1]) = kAfter;  // used by step s+2
                if (!(spu > 0.0) || !(sstate > 0.0) || isinf(sstate)) {  // core.py:440-452
                    dead[c] = true;
                    live[c] = false;
                    flush(i + 1, (int)(s & 31));
                    __syncwarp();
                    if (lane == 0) {
                        *reinterpret_cast<volatile int *>(&q.dead) = (int)s;
                        __threadfence_block();
                        *reinterpret_cast<volatile int *>(&q.svc_done) = 0x7fffffff;
                        a.logE[b] = -INFINITY;
                        if (a.alive) a.alive[b] = -1;
                    }
                    continue;
                }
                if (lane == (int)(s & 31)) {
                    mySpu[c] = spu;
                    mySql[c] = sql;
                }
                if ((s & 31) == 31 || i == 0) flush(i, (int)(s & 31) + 1);
                double *row = a.alpha_seq + b * a.seq_stride + i * (long long)n;
                double *P = sm + (long long)c * a.il_stride + a.off_stage + sb * Gp;
                if (rawRows) {
                    if (lane == 0) {
                        fence_proxy_async();
                        bulk_store(row, P, rowBytes);
                    }
                } else {
                    const double inv = fast_rcp(spu);
                    for (int j = 2 * lane; j < n; j += 64) {
                        double2 x = *reinterpret_cast<const double2 *>(P + j);
                        x.x *= inv;
                        x.y *= inv;
                        __stcs(reinterpret_cast<double2 *>(row + j), x);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (rawRows) bulk_wait_read<0>();
                    if (i >= 2) {  // the slot is free again: prefetch alpha[i-2] into it
                        const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : a.alpha_seq + b * a.seq_stride;
                        fence_proxy_async();
                        bulk_load(P, src + (i - 2) * (long long)n, rowBytes, &bars[2 * c + sb]);
                    }
                    __threadfence_block();
                    *reinterpret_cast<volatile int *>(&q.svc_done) = (int)s + 1;
                }
            }
        }
        if (rawRows && lane == 0) bulk_wait_all();
    }
}

}  // namespace blg
