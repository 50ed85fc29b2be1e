// fast1d_ws.cuh -- K1w/K2w: warp-specialised version of the fast 1-D kernels (1-D grid, program = ONE
// GaussianRandomWalk: BASELINE.json configs[0], configs[1]).
//
// Why (measured on B200 with ncu, profiles/r1b_*): in the single-role kernels of fast1d.cuh the convolution is 57 % of
// the executed instructions but only 37 % of the stall samples -- the time goes into the short serial section around
// it (strided 8-byte global loads/stores at ~28 sectors per request, the 5-level shuffle reduction, the reciprocal,
// the row store), which a sweep with only ~3.5 resident chains per SM cannot hide.  Here a CTA is
//     NCW = NT/32 - 1 COMPUTE warps: per step  lik loads -> convolution (registers) -> x kappa x lik -> state + one
//                                    partial sum per thread to shared memory -> CTA barrier;
//     1 SERVICE warp:                after the barrier reduces the partial sums, computes the normaliser (Newton
//                                    reciprocal), hands it back through shared memory + a named barrier, stores the
//                                    normalised row with coalesced 16-byte accesses (forward) / scales and stores the
//                                    smoothed row and refills the alpha ring with bulk-async copies (backward), keeps
//                                    the log-evidence product -- all of it overlapped with the next convolution.
// The service warp's index rotates with the CTA's slot on its SM so that the four sub-partitions get the same mix.
// The likelihood table of the call is generated in "owner order" ([t][m][thread], lik_table_perm_kernel) so that
// every compute thread fetches its M cells with fully coalesced loads and cells beyond the grid read 0.
//
// Semantics: identical to fast1d.cuh (core.py:372-417, :424-470; transitionModels.py:96-118).
#pragma once

#include "fast1d.cuh"

namespace blg {

__device__ __forceinline__ void named_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// combo of this CTA (see combo_of_sm in common.cuh) plus the CTA's arrival index on its SM
__device__ __forceinline__ long long combo_and_slot(const PassArgs &a, int *sh, int &slot, int &sm) {
    if (threadIdx.x == 0) {
        int b = -1, k = 0;
        unsigned smid = 0xffffffffu;
        if (a.sm_assign) {
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            int *counters = a.sm_state, *claimed = a.sm_state + a.sm_count;
            b = -2;
            if ((int)smid < a.sm_count) {
                k = atomicAdd(&counters[smid], 1);
                if (k < a.sm_slots) {
                    b = a.sm_assign[smid * a.sm_slots + k];
                    if (b >= 0 && atomicCAS(&claimed[b], 0, 1) != 0) b = -1;
                }
            }
            if (b == -2) {  // late arrival: adopt a combo nobody has claimed yet
                b = -1;
                for (long long j = 0; j < a.B; ++j) {
                    const int cand = a.order ? a.order[j] : (int)j;
                    if (atomicCAS(&claimed[cand], 0, 1) == 0) {
                        b = cand;
                        break;
                    }
                }
            }
        } else {
            const long long j = combo_of_block(a);
            b = (int)(a.order ? a.order[j] : j);
            k = (int)(blockIdx.x / (a.num_sms > 0 ? a.num_sms : 1));
        }
        sh[0] = b;
        sh[1] = k;
        sh[2] = (a.sm_assign && (int)smid < a.sm_count) ? (int)smid : -1;
    }
    __syncthreads();
    slot = sh[1];
    sm = sh[2];
    return sh[0];
}

// PACING (round 2).  The per-CTA trace of a C2 sweep (profiles/r2s_sm_timeline.txt) showed that the chains sharing an SM
// do NOT advance together: the warp schedulers favour some warps, one chain of an SM finishes after 45 % of the kernel,
// the next after 60 %, and the last one runs alone for a third of the time -- with a single chain resident the FP64
// pipe idles during every epilogue and barrier (16 % of all SM-time had one active chain, 17 % two).  Every chain now
// publishes its step count (one global store by the service warp every pace_every steps) and holds back -- the service
// warp delays its arrival at the step barrier -- while it is more than pace_skew steps ahead of a chain of the same
// SM's list that has started.  The chain with the least progress never waits, so there is no deadlock; a chain that
// has not started (-1) or has finished (INT_MAX) never blocks anybody; a bounded number of polls is a safety net
// (pacing only shapes the schedule, results do not depend on it).
struct WsPace {
    int *progress;  // NULL: off
    int peer;       // this lane's peer combo or -1
    int every, skew;
    bool tired;     // gave up waiting once: stop pacing
};

__device__ __forceinline__ WsPace ws_pace_init(const PassArgs &a, long long b, int sm, int lane) {
    WsPace p;
    p.progress = (a.ws_progress && a.sm_assign && sm >= 0 && a.pace_every > 0) ? a.ws_progress : nullptr;
    p.peer = -1;
    p.every = a.pace_every;
    p.skew = a.pace_skew;
    p.tired = false;
    if (p.progress && lane < a.sm_slots) {
        const int q = a.sm_assign[sm * a.sm_slots + lane];
        if (q >= 0 && q != (int)b) p.peer = q;
    }
    return p;
}

__device__ __forceinline__ void ws_pace_done(const PassArgs &a, long long b) {
    if (a.ws_progress) *reinterpret_cast<volatile int *>(a.ws_progress + b) = 0x7fffffff;
}

// called by the whole service warp in front of the barrier of step `step` (steps done so far)
__device__ __forceinline__ void ws_pace(WsPace &p, long long b, long long step, int lane) {
    if (!p.progress || p.tired || ((int)step & (p.every - 1)) != 0) return;  // every = power of two
    if (lane == 0) *reinterpret_cast<volatile int *>(p.progress + b) = (int)step;
    bool gaveUp = false;
    if (p.peer >= 0) {
        const int need = (int)step - p.skew;
        const volatile int *q = p.progress + p.peer;
        for (int polls = 0;; ++polls) {
            const int v = *q;
            if (v < 0 || v >= need) break;
            if (polls > 20000) {  // ~10 ms: something is wrong with the peer; never hang on a scheduling hint
                gaveUp = true;
                break;
            }
            __nanosleep(200);
        }
    }
    if (__any_sync(0xffffffffu, gaveUp)) p.tired = true;
}

// v[0..M) -> cells i0.. of a haloed line, plus their mirror images inside the halo (halo <= n)
template <int M>
__device__ __forceinline__ void store_cells_mirrored(double *line, int i0, int n, int halo, const double (&v)[M]) {
    if (i0 + M <= n) {
#pragma unroll
        for (int m = 0; m < M; ++m) line[i0 + m] = v[m];
    } else {
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (i0 + m < n) line[i0 + m] = v[m];
    }
    if (i0 < halo) {
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (i0 + m < halo && i0 + m < n) line[-1 - (i0 + m)] = v[m];
    }
    if (i0 + M > n - halo) {
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (i0 + m >= n - halo && i0 + m < n) line[2 * n - 1 - (i0 + m)] = v[m];
    }
}

__device__ __forceinline__ void trace_end_lane0(const PassArgs &a, int lane) {
    if (a.trace && lane == 0) a.trace[4 * (long long)blockIdx.x + 3] = global_ns();
}

struct WsRoles {
    int lane, warp, rot, ct, i0;
    bool service, owner, shortw;
};

// UNEVEN SPLIT (round 2).  Every compute warp issues its DFMAs for 32 lanes x (cells per thread) whether the lanes own
// grid cells or not: at G = 1000 four warps x 9 cells cover 1152 slots, 13 % of the FP64 work is spent on cells beyond
// the grid.  With ML < M the LAST compute warp owns ML cells per thread (its own chunk layout of the weights, its own
// instantiation of the per-step body): 3 x 32 x 9 + 32 x 5 = 1024 slots.  The CTAs of an SM start on different
// sub-partitions (5 warps per CTA), so the short warps of the resident chains spread over the four sub-partitions.
template <int M, int ML, int NT>
__device__ __forceinline__ WsRoles ws_roles(int slot, int n) {
    constexpr int NW = NT / 32, NCW = NW - 1;
    WsRoles r;
    r.lane = threadIdx.x & 31;
    r.warp = threadIdx.x >> 5;
    // the service role rotates with the CTA's slot so that the four sub-partitions (warp index mod 4) see the same mix;
    // with a number of warps that is not a multiple of four the hardware's slot allocation already rotates the CTAs,
    // and the LAST warp serves: the compute warps of a 5-warp CTA then sit on four different sub-partitions
    r.rot = (NW % 4 == 0) ? slot % NW : NW - 1;
    r.service = r.warp == r.rot;
    const int cw = r.warp - (r.warp > r.rot ? 1 : 0);
    r.ct = cw * 32 + r.lane;
    r.shortw = (ML != M) && cw == NCW - 1;
    r.i0 = r.shortw ? (NCW - 1) * 32 * M + r.lane * ML : r.ct * M;
    r.owner = !r.service && r.i0 < n;
    return r;
}

// Control block of the ws kernels (doubles at a.ws_ctl): [0],[1] scale factor by step parity, [2] (as int) the step at
// which the service warp found a zero norm (-1: alive).  The partial sums PP are double buffered by step parity: the
// compute warps of step t+1 write while the service warp may still be reading those of step t.  All CTA-wide barriers
// inside the role branches are NAMED barriers with an explicit thread count (bar.sync 1, NT): the two roles reach them
// from different code paths, which __syncthreads() does not allow; both roles execute the same number of them.
//
// LAGGED SCALE (round 2).  The state is kept unnormalised; the only reason to scale it at all is to keep its magnitude
// in range.  Round 1 applied the exact normaliser of the PREVIOUS step, which the service warp produces ~400 cycles
// after the step barrier: the compute warps of short steps waited for it at a second (named) barrier -- 13 % of
// their stall samples in profiles/r1j_regions_c2_c3.txt.  Now step t multiplies by
//         k_t = s_{t-2}^(-3/4),     s_j = sum of the unnormalised state after step j,
// a value that has been in shared memory since the service warp passed the barrier of step t-1, so ONE CTA barrier
// per step remains and nobody waits for the service warp.  With u_t = conv(u_{t-1}) * k_t * lik_t the log-magnitude
// obeys x_t = x_{t-1} - 3/4 x_{t-2} + log(norm_t): a damped recursion (|roots| = 0.87, fixed point 4/3 log(norm));
// the exponent 1 would be marginally stable (|roots| = 1) and random-walk out of range over 10^4 steps.  The true
// evidence increment is recovered exactly by the service warp, norm_t = s_t / (k_t * s_{t-1}) (core.py:385), rows are
// normalised with the exact 1/s_t, or leave unnormalised (BLG_F_RAW_ALPHA / BLG_F_RAW_POSTERIOR) by one bulk store.
__device__ __forceinline__ double lagged_scale(double s) {  // s^(-3/4) to a few ulp (any positive factor would do)
    const double r = rsqrt(s);
    return r * sqrt(r);
}

// ------------------------------------------------------------------------------------------------ K1w forward
// The per-step loop of one compute warp with MM cells per thread (MM = M, or ML for the short last warp); NCOMP =
// compute threads of the CTA = pitch of one plane of the owner-order likelihood table.
template <int MM, int NT>
__device__ __forceinline__ void ws_fwd_compute(const PassArgs &a, const Fast1dSetup &s, const WsRoles &r, const double *W,
                                               double *PP, volatile double *ctl, volatile int *deadFlag, bool rawRows,
                                               bool first) {
    constexpr int NCOMP = (NT / 32 - 1) * 32;
    const int n = a.pb.G, halo = a.halo;
    const long long T = a.T;
    double *cur = s.buf0, *nxt = s.buf1;
    const double *likp = a.lik_table + r.ct;
    const long long pitch = a.lik_pitch;
    double lk[MM];  // likelihood of this thread's cells, fetched one step ahead (right after the previous use)
    if (r.owner) {
#pragma unroll
        for (int m = 0; m < MM; ++m) lk[m] = __ldg(likp + m * NCOMP);
    }
    // debugging aid (BLG_TRACE): cycles this warp spends in the convolution / the epilogue / waiting at the barrier
    const bool prof = a.trace != nullptr;
    long long cConv = 0, cEpi = 0, cBar = 0;
    for (long long t = 0; t < T; ++t) {
        double v[MM];
        const bool trans = (t > 0 || first) && (t - 1 >= s.f_lo) && (t - 1 < s.f_hi);
        const long long p0 = prof ? clock64() : 0;
        if (r.owner) {
            if (trans && s.R > 0) {
                conv_item<MM>(cur, r.i0, s.R, W, v);  // transitionModels.py:111
            } else {
#pragma unroll
                for (int m = 0; m < MM; ++m) v[m] = cur[r.i0 + m];
            }
        }
        const long long p1 = prof ? clock64() : 0;
        // lagged scale k_t: written by the service warp before it arrived at the barrier of step t-1
        const double kappa = t >= 2 ? ctl[t & 1] : 1.0;
        if (r.owner) {
            // alpha <- prior * likelihood (core.py:375-382); cells beyond the grid carry lik = 0
#pragma unroll
            for (int m = 0; m < MM; ++m) v[m] *= kappa * lk[m];
            if (t + 1 < T) {
#pragma unroll
                for (int m = 0; m < MM; ++m) lk[m] = __ldg(likp + (t + 1) * pitch + m * NCOMP);
            }
            store_cells_mirrored<MM>(nxt, r.i0, n, halo, v);
            PP[(t & 1) * NCOMP + r.ct] = tree_sum<MM>(v);
        }
        if (rawRows) fence_proxy_async();  // the new state is read by the service warp's bulk-async row store
        const long long p2 = prof ? clock64() : 0;
        named_sync(1, NT);  // new state and its partial sums are visible to everybody
        if (prof) {
            cConv += p1 - p0;
            cEpi += p2 - p1;
            cBar += clock64() - p2;
        }
        {   // a zero norm found by the service warp behind the barrier of an EARLIER step ends the chain here; the
            // flag carries the step so that both roles leave after the same number of barriers
            const int ds = *deadFlag;
            if (ds >= 0 && ds < t) break;
        }
        double *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    if (prof && r.lane == 0) {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        long long *w = a.trace + 4 * (long long)gridDim.x + ((long long)blockIdx.x * 8 + r.warp) * 4;
        w[0] = cConv;
        w[1] = cEpi;
        w[2] = cBar;
        w[3] = wid;
    }
}

template <int M, int ML, int NT>
__global__ void __launch_bounds__(NT, NT > 192 ? 2 : 4) fwd_fast1d_ws_kernel(const PassArgs a) {
    static_assert(ML <= M, "the short warp owns at most M cells per thread");
    constexpr int NW = NT / 32, NCOMP = (NW - 1) * 32;
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    int slot, smid;
    const long long b = combo_and_slot(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6), slot, smid);
    if (b < 0) return;
    trace_begin(a, b);
    const int n = pb.G;
    const long long T = a.T;
    Fast1dSetup s;
    // weights, windows (fast1d_setup without the per-cell likelihood tables: the table is always shared here)
    s.buf0 = sm + a.halo;
    s.buf1 = sm + (a.Gp + 2 * a.halo) + a.halo;
    s.rs.buf = sm + a.off_misc;
    s.rs.phase = 0;
    s.W = sm + a.off_w;
    s.sigma = a.pg.param[b];
    s.R = a.pg.radius[b];
    {
        const int *win = a.pg.window + b * 4;
        s.f_lo = win[0];
        s.f_hi = win[1];
    }
    if (!(s.sigma > 0.0) || s.R <= 0) s.R = 0;
    if ((2 * s.R + M) / M * (M + 1) > a.pg.w_len[0] || (ML != M && (2 * s.R + ML) / ML * (ML + 1) > a.ws_w2_len)) {
        if (threadIdx.x == 0) {  // radius beyond blg_program.max_radius
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
            ws_pace_done(a, b);
        }
        return;
    }
    const double wInv = build_weights_chunked<M>(s.W, a.pg.w_len[0], s.sigma, s.R, s.rs);
    double *const WL = sm + a.ws_w2;  // the same weights in the chunk layout of the short warp
    if (ML != M) copy_weights_chunked<ML>(WL, a.ws_w2_len, s.sigma, s.R, wInv);
    const WsRoles r = ws_roles<M, ML, NT>(slot, n);
    double *PP = sm + a.ws_part;
    volatile double *ctl = sm + a.ws_ctl;
    volatile int *deadFlag = reinterpret_cast<volatile int *>(sm + a.ws_ctl + 2);
    {
        const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)n : a.prior;
        for (int g = threadIdx.x; g < n; g += NT) store_mirrored(s.buf0, g, n, a.halo, init[g]);
        for (int j = threadIdx.x; j < 2 * NCOMP; j += NT) PP[j] = 0.0;  // threads without cells never write their slots
        if (threadIdx.x == 0) *deadFlag = -1;
    }
    __syncthreads();
    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    const bool first = (a.flags & BLG_F_TRANSITION_FIRST) != 0;
    // a backward pass follows (it is scale-free per row): rows leave unnormalised by one bulk-async copy per step
    const bool rawRows = store && a.use_bulk != 0 && (a.flags & BLG_F_RAW_ALPHA);

    if (!r.service) {
        // ------------------------------------------------------------------ compute warps
        if (ML != M && r.shortw)
            ws_fwd_compute<ML, NT>(a, s, r, WL, PP, ctl, deadFlag, rawRows, first);
        else
            ws_fwd_compute<M, NT>(a, s, r, s.W, PP, ctl, deadFlag, rawRows, first);
    } else {
        // ------------------------------------------------------------------ service warp
        double *seq = store ? a.alpha_seq + b * a.seq_stride : nullptr;
        const bool vec = a.use_bulk != 0;  // rows are 16-byte aligned
        const bool raw = rawRows;
        const uint32_t rowBytes = (uint32_t)(n * sizeof(double));
        LogProduct lp;
        lp.init();
        bool dead = false;
        double sPrev = 1.0;            // s_{t-1}; the initial state enters as it is (core.py:363, :382)
        double kNow = 1.0, kNext = 1.0;  // k_t, k_{t+1}
        WsPace pace = ws_pace_init(a, b, smid, r.lane);
        for (long long t = 0; t < T; ++t) {
            ws_pace(pace, b, t, r.lane);
            named_sync(1, NT);
            if (dead) break;  // the compute warps see the flag behind this barrier
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < NCOMP / 32; ++j) part += PP[(t & 1) * NCOMP + j * 32 + r.lane];
            const double st_sum = warp_sum(part);
            const double kAfter = lagged_scale(st_sum);  // k_{t+2}
            if (r.lane == 0) ctl[t & 1] = kAfter;
            const double norm = fast_div(st_sum, kNow * sPrev);  // core.py:385: evidence increment of step t
            if (!(st_sum > 0.0) || !(norm > 0.0) || isinf(st_sum)) {  // core.py:388-400
                dead = true;
                if (r.lane == 0) *deadFlag = (int)t;
                continue;  // one more barrier: the compute warps read the flag behind it
            }
            const double *st = (t & 1) ? s.buf0 : s.buf1;  // the buffer the compute warps just filled
            if (raw) {
                // the row leaves unnormalised, straight out of the state buffer; the buffer is rewritten by the compute
                // warps after the NEXT barrier, so the copy must have read it before this warp arrives there
                if (r.lane == 0) {
                    fence_proxy_async();
                    bulk_store(seq + t * (long long)n, st, rowBytes);
                }
            } else if (store) {  // core.py:389, :408 -- normalised filtering distribution
                const double inv = fast_rcp(st_sum);
                double *row = seq + t * (long long)n;
                if (vec) {
                    for (int j = 2 * r.lane; j < n; j += 64) {
                        double2 x = *reinterpret_cast<const double2 *>(st + j);
                        x.x *= inv;
                        x.y *= inv;
                        __stcs(reinterpret_cast<double2 *>(row + j), x);
                    }
                } else {
                    for (int j = r.lane; j < n; j += 32) __stcs(row + j, st[j] * inv);
                }
            }
            if ((a.flags & BLG_F_SAVE_STATE) && a.final_state && t == T - 1) {
                const double inv = fast_rcp(st_sum);
                double *fs = a.final_state + b * (long long)n;
                for (int j = r.lane; j < n; j += 32) fs[j] = st[j] * inv;
            }
            if (r.lane == 0) {
                lp.mul(norm);                                         // core.py:403
                if (a.local) a.local[b * a.row_stride + t] = norm * pb.lc_prod;  // core.py:404
                if (raw) bulk_wait_read<0>();
            }
            sPrev = st_sum;
            kNow = kNext;
            kNext = kAfter;
        }
        if (r.lane == 0) {
            if (raw) bulk_wait_all();
            double logE = lp.log_value();
            if (dead)
                logE = -INFINITY;
            else if (!(a.flags & BLG_F_INIT_STATE))
                logE += log(pb.lc_prod);  // core.py:417
            a.logE[b] = logE;
            if (a.alive) a.alive[b] = dead ? 0 : 1;
            ws_pace_done(a, b);
        }
        trace_end_lane0(a, r.lane);
    }
}

// ------------------------------------------------------------------------------------------------ K2w backward
template <int MM, int NT>
__device__ __forceinline__ void ws_bwd_compute(const PassArgs &a, const Fast1dSetup &s, const WsRoles &r, const double *W,
                                               double *PP, volatile double *ctl, volatile int *deadFlag, bool rawRows,
                                               double *S0, uint64_t *bars) {
    constexpr int NCOMP = (NT / 32 - 1) * 32;
    const int n = a.pb.G, halo = a.halo, Gp = a.Gp;
    const long long T = a.T;
    double *cur = s.buf0, *nxt = s.buf1;
    const double *likp = a.lik_table + r.ct;
    const long long pitch = a.lik_pitch;
    uint32_t phases = 0u;  // bit s = parity of the next completion of ring slot s
    double beta[MM];
#pragma unroll
    for (int m = 0; m < MM; ++m) beta[m] = (r.owner && r.i0 + m < n) ? 1.0 / (double)n : 0.0;  // core.py:424-425
    double lk[MM];  // likelihood of this thread's cells, fetched one step ahead (right after the previous use)
    if (r.owner) {
#pragma unroll
        for (int m = 0; m < MM; ++m) lk[m] = __ldg(likp + (T - 1) * pitch + m * NCOMP);
    }
    for (long long i = T - 1; i >= 0; --i) {
        const int sb = (int)(i & 1);
        if (i < T - 1) {
            const bool trans = (i + 1 >= s.b_lo) && (i + 1 < s.b_hi);
            if (r.owner) {
                if (trans && s.R > 0) {
                    conv_item<MM>(cur, r.i0, s.R, W, beta);  // transitionModels.py:117-118
                } else {
#pragma unroll
                    for (int m = 0; m < MM; ++m) beta[m] = cur[r.i0 + m];
                }
#pragma unroll
                for (int m = 0; m < MM; ++m)
                    if (r.i0 + m >= n) beta[m] = 0.0;
            }
        }
        // keeps the (scale-free) beta recursion in range: lagged power of sum(beta) two steps back (see lagged_scale)
        const double kb = i <= T - 3 ? ctl[i & 1] : 1.0;
        mbar_wait(&bars[sb], (phases >> sb) & 1u);
        phases ^= 1u << sb;
        double *A = S0 + sb * Gp;
        if (r.owner) {
            double st[MM];
            double spu0 = 0.0, spu1 = 0.0, sql0 = 0.0, sql1 = 0.0;  // two chains each: short dependency paths
#pragma unroll
            for (int m = 0; m < MM; ++m) {
                const int li = r.i0 + m;
                const double al = li < n ? A[li] : 0.0;
                const double pu = al * beta[m];                            // posterior ~ alpha*beta   core.py:436
                const double ql = li < n ? fast_div_pos(pu, lk[m]) : 0.0;  // core.py:463
                st[m] = beta[m] * kb * lk[m];                              // beta*likelihood          core.py:467
                if (li < n) A[li] = pu;
                if (m & 1) {
                    spu1 += pu;
                    sql1 += ql;
                } else {
                    spu0 += pu;
                    sql0 += ql;
                }
            }
            if (i > 0) {
#pragma unroll
                for (int m = 0; m < MM; ++m) lk[m] = __ldg(likp + (i - 1) * pitch + m * NCOMP);
            }
            store_cells_mirrored<MM>(nxt, r.i0, n, halo, st);
            double *pp = PP + sb * 3 * NCOMP;
            pp[r.ct] = spu0 + spu1;
            pp[NCOMP + r.ct] = tree_sum<MM>(st);  // sum of the new state (magnitude control only)
            pp[2 * NCOMP + r.ct] = sql0 + sql1;
        }
        if (rawRows) fence_proxy_async();  // alpha * beta in the ring slot is read by the bulk-async row store
        named_sync(1, NT);
        {
            const int ds = *deadFlag;  // step at which the service warp found a zero norm (rows run downwards)
            if (ds >= 0 && ds > i) break;
        }
        double *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
}

template <int M, int ML, int NT>
__global__ void __launch_bounds__(NT, NT > 192 ? 2 : 4) bwd_fast1d_ws_kernel(const PassArgs a) {
    static_assert(ML <= M, "the short warp owns at most M cells per thread");
    constexpr int NW = NT / 32, NCOMP = (NW - 1) * 32;
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    int slot, smid;
    const long long b = combo_and_slot(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6), slot, smid);
    if (b < 0) return;
    trace_begin(a, b);
    if (a.alive && a.alive[b] != 1) {  // the forward pass aborted (core.py:400)
        if (threadIdx.x == 0) ws_pace_done(a, b);
        return;
    }
    const int n = pb.G;
    const long long T = a.T;
    Fast1dSetup s;
    s.buf0 = sm + a.halo;
    s.buf1 = sm + (a.Gp + 2 * a.halo) + a.halo;
    s.rs.buf = sm + a.off_misc;
    s.rs.phase = 0;
    s.W = sm + a.off_w;
    s.sigma = a.pg.param[b];
    s.R = a.pg.radius[b];
    {
        const int *win = a.pg.window + b * 4;
        s.b_lo = win[2];
        s.b_hi = win[3];
    }
    if (!(s.sigma > 0.0) || s.R <= 0) s.R = 0;
    if ((2 * s.R + M) / M * (M + 1) > a.pg.w_len[0] || (ML != M && (2 * s.R + ML) / ML * (ML + 1) > a.ws_w2_len)) {
        if (threadIdx.x == 0) ws_pace_done(a, b);
        return;
    }
    const double wInv = build_weights_chunked<M>(s.W, a.pg.w_len[0], s.sigma, s.R, s.rs);
    double *const WL = sm + a.ws_w2;  // the same weights in the chunk layout of the short warp
    if (ML != M) copy_weights_chunked<ML>(WL, a.ws_w2_len, s.sigma, s.R, wInv);
    const WsRoles r = ws_roles<M, ML, NT>(slot, n);
    double *PP = sm + a.ws_part;  // [3][NCOMP]
    volatile double *ctl = sm + a.ws_ctl;
    volatile int *deadFlag = reinterpret_cast<volatile int *>(sm + a.ws_ctl + 2);
    double *seq = a.alpha_seq + b * a.seq_stride;
    const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : seq;  // filtering rows (out-of-place smoothing)
    double *const S0 = sm + a.off_stage;  // alpha[t] ring: 2 slots of Gp doubles; overwritten in place by alpha*beta
    const int Gp = a.Gp;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + a.off_misc + kMiscBarrierOffset);
    const uint32_t rowBytes = (uint32_t)(n * sizeof(double));
    for (int j = threadIdx.x; j < 6 * NCOMP; j += NT) PP[j] = 0.0;  // threads without cells never write their slots
    if (threadIdx.x == 0) {
        *deadFlag = -1;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_proxy_async();
    }
    __syncthreads();
    const bool rawRows = a.row_scale != nullptr;  // BLG_F_RAW_POSTERIOR: rows leave unnormalised + their factor

    if (!r.service) {
        // ------------------------------------------------------------------ compute warps
        if (ML != M && r.shortw)
            ws_bwd_compute<ML, NT>(a, s, r, WL, PP, ctl, deadFlag, rawRows, S0, bars);
        else
            ws_bwd_compute<M, NT>(a, s, r, s.W, PP, ctl, deadFlag, rawRows, S0, bars);
    } else {
        // ------------------------------------------------------------------ service warp
        const bool raw = rawRows;
        if (r.lane == 0) {
            bulk_load(S0 + ((T - 1) & 1) * Gp, src + (T - 1) * (long long)n, rowBytes, &bars[(T - 1) & 1]);
            if (T >= 2) bulk_load(S0 + ((T - 2) & 1) * Gp, src + (T - 2) * (long long)n, rowBytes, &bars[(T - 2) & 1]);
        }
        bool dead = false;
        long long i = T - 1;
        WsPace pace = ws_pace_init(a, b, smid, r.lane);
        for (; i >= 0; --i) {
            const int sb = (int)(i & 1);
            ws_pace(pace, b, T - 1 - i, r.lane);
            named_sync(1, NT);
            if (dead) break;
            double spu = 0.0, sstate = 0.0, sql = 0.0;
            const double *pp = PP + sb * 3 * NCOMP;
#pragma unroll
            for (int j = 0; j < NCOMP / 32; ++j) {
                spu += pp[j * 32 + r.lane];
                sstate += pp[NCOMP + j * 32 + r.lane];
                sql += pp[2 * NCOMP + j * 32 + r.lane];
            }
            spu = warp_sum(spu);
            sstate = warp_sum(sstate);
            sql = warp_sum(sql);
            if (r.lane == 0) ctl[i & 1] = lagged_scale(sstate);  // used by step i-2
            if (!(spu > 0.0) || !(sstate > 0.0) || isinf(sstate)) {  // core.py:440-452
                dead = true;
                if (r.lane == 0) *deadFlag = (int)i;
                continue;
            }
            const double inv = fast_rcp(spu);  // posterior = alpha*beta / sum(alpha*beta)   core.py:439-441
            double *row = seq + i * (long long)n;
            double *P = S0 + sb * Gp;
            if (raw) {
                if (r.lane == 0) {
                    fence_proxy_async();
                    bulk_store(row, P, rowBytes);
                    a.row_scale[b * a.row_stride + i] = inv;
                }
            } else {
                for (int j = 2 * r.lane; j < n; j += 64) {
                    double2 x = *reinterpret_cast<const double2 *>(P + j);
                    x.x *= inv;
                    x.y *= inv;
                    __stcs(reinterpret_cast<double2 *>(row + j), x);
                }
            }
            __syncwarp();
            if (r.lane == 0) {
                if (a.local) a.local[b * a.row_stride + i] = fast_div(spu, sql * pb.lc_prod);  // 1/(sum(post/lik)*lc)  core.py:463
                if (raw) bulk_wait_read<0>();
                if (i >= 2) {  // the slot is free again: prefetch alpha[i-2] into it
                    fence_proxy_async();
                    bulk_load(P, src + (i - 2) * (long long)n, rowBytes, &bars[sb]);
                }
            }
        }
        if (dead) {
            // drain the prefetch that is still in flight before the CTA's shared memory is released: the step that died
            // (index i + 1 after the loop's decrement and the extra barrier) consumed its slot; the other slot holds the
            // load issued one step earlier
            const long long died = i + 1;
            if (died >= 1) {
                const long long rr = died - 1;
                mbar_wait(&bars[rr & 1], (uint32_t)(((T - 1 - rr) >> 1) & 1));
            }
            if (r.lane == 0) {
                a.logE[b] = -INFINITY;
                if (a.alive) a.alive[b] = -1;
            }
        }
        if (raw && r.lane == 0) bulk_wait_all();
        if (r.lane == 0) ws_pace_done(a, b);
        trace_end_lane0(a, r.lane);
    }
}

}  // namespace blg
