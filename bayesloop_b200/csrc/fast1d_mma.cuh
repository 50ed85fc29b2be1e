// fast1d_mma.cuh -- K1m/K2m: the warp-specialised fast 1-D kernels (fast1d_ws.cuh) with the GaussianRandomWalk
// convolution issued as FP64 matrix instructions (mma.sync.m8n8k4.f64, "DMMA").
//
// Why (B200, profiles/r2s_dmma_mix.txt, r2v_regions_c2_ws.txt): the convolution is FP64-FMA work on either path -- DMMA and DFMA
// share the FP64 units (a mix of the two never exceeds the faster one) -- but
//   * a warp of DFMAs reaches 54-57 MAC/clk/SM in isolation, DMMA 63.8 (the full 64 lanes), and
//   * in the real kernel the DFMA loop keeps the pipe only ~62 % busy: with 3-4 compute warps per SM sub-partition the
//     schedulers spend their cycles on the 2-cycle issue cadence, operand-collector waits and the LDS -> DFMA
//     dependencies of 81-instruction chunks (stall_wait 29 %, selected 20 %, math 13 %, short scoreboard 8 % of the
//     samples inside the loop).  One DMMA is 256 MACs = 16 cycles of pipe time per instruction: two resident warps
//     saturate the pipe and the instruction count of the loop drops 8x.
//
// The convolution as a matrix product (no reshaping of the problem: the band of a Toeplitz matrix):
//     y[c] = sum_k w[k] x[c - R + k]            (transitionModels.py:111; reflect boundary = mirrored halo cells)
//   one 8 x 8 tile of outputs  Y[a][r] = y[tile + 8a + r]  accumulates, per group of 8 input offsets s = s0 + 8j .. + 7
//   (s0 = -R rounded down to an even number), TWO instructions (even and odd offsets):
//     A[a][u] = x[tile + 8a + s0 + 8j + 2u (+1)]   (8 x 4: ONE 16-byte LDS per lane feeds both -- 512 contiguous bytes per
//                                                   warp, conflict free, address = running pointer + constant),
//     B[u][r] = w[s0 + 8j + 2u (+1) - r + R]       (4 x 8 Toeplitz block, 0 outside 0..2R; the same for every tile),
//   over ceil((2R + 8) / 8) groups (the 8 outputs of a row need the offsets -R .. R+7): 2R + 11.5 taps on average
//   instead of 2R + 1 -- 7 of the extra taps are the price of sharing one input window between 8 neighbouring outputs,
//   the rest is the granularity of 8.  A first version used k-steps of 4 offsets on an XOR-swizzled state (2R + 11
//   taps): 12 instructions per DMMA, most of them address arithmetic, issue slots 47 % busy and the matrix pipe 63 %
//   (profiles/r2_dmma_variants.txt); here it is ~2 per DMMA and the pipe 75 % busy in the forward pass.
//   A compute warp owns TPW tiles (64 cells each, a compile-time distance apart), i.e. a lane owns the cell PAIRS
//   (tile + 8*(lane/4) + 2*(lane%4), +1) of its tiles -- the accumulator fragment of the instruction -- so likelihood
//   loads, state stores and row stores are 16-byte accesses, 512 contiguous bytes per warp: the forward rows go to HBM
//   straight from the registers of the compute warps (no service-warp copy, no bulk store).
//
// Roles, one named barrier per step, alpha ring of the backward pass and zero-norm protocol are those of fast1d_ws.cuh;
// the service warp is lean (see below), the chains of an SM are paced (ws_pace).  Semantics: core.py:372-417, :424-470; transitionModels.py:96-118.
#pragma once

#include <type_traits>

#include "fast1d_ws.cuh"

namespace blg {

constexpr int kMmaWPad = 16;  // zero weights in front of / behind the 2R+1 taps (the Toeplitz blocks reach 14 beyond)

__device__ __forceinline__ int swz(int i) { return i; }  // natural layout (an XOR swizzle lived here, see the header)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    // volatile: keeps the issue order chosen in mma_conv_body (independent accumulators between dependent instructions)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// halo cells on each side of the state: the groups reach R + 7 below the first and R + 14 beyond the last valid cell;
// only [-halo, n + halo) holds mirrored values, the rest of the buffer stays zero (read with zero weights only)
__host__ __device__ __forceinline__ int mma_halo(int R) { return (R + 7 + 7) & ~7; }

// zero-padded weights Wz[k + kMmaWPad] = w[k] (build_weights in common.cuh), normalised; len >= 2R + 1 + 2*kMmaWPad
__device__ __forceinline__ void mma_build_weights(double *Wz, int len, double sigma, int R, RedScratch &rs) {
    if (!(sigma > 0.0) || R <= 0) {  // transitionModels.py:110-113: identity (never read: R = 0 skips the convolution)
        for (int q = threadIdx.x; q < len; q += blockDim.x) Wz[q] = q == kMmaWPad ? 1.0 : 0.0;
        __syncthreads();
        return;
    }
    const double h = -0.5 / (sigma * sigma);
    double part = 0.0;
    for (int q = threadIdx.x; q < len; q += blockDim.x) {
        const int k = q - kMmaWPad;
        double v = 0.0;
        if (k >= 0 && k <= 2 * R) {
            const double x = (double)(k - R);
            v = exp(h * x * x);
        }
        Wz[q] = v;
        part += v;
    }
    const double inv = 1.0 / block_sum(part, rs);
    for (int q = threadIdx.x; q < len; q += blockDim.x) Wz[q] *= inv;
    __syncthreads();
}

// LEAN SERVICE WARP.  The step barrier counts the service warp, so a chain cannot step faster than the service warp's
// loop -- and that loop was a chain of ~50 DEPENDENT FP64 instructions (rsqrt + sqrt for the lagged scale, the division
// for the evidence increment, two frexp for the running product), each of which queues for the FP64 pipe behind the
// matrix instructions of the compute warps of every chain on the sub-partition.  The per-warp trace showed it: every
// chain took 5 300 - 6 600 cycles per step whatever its radius (profiles/r2d_ws_trace.txt), and an SM's throughput did
// not depend on how many chains were running (profiles/r2u_ws_pacing.txt).  Here:
//   * the lagged scale is a POWER OF TWO applied on demand (ondemand_scale): integer instructions only, multiplying by it
//     is exact, and on most steps it is 1 and skipped;
//   * the log-evidence telescopes: sum_t log norm_t = log s_{T-1} - sum_t log k_t = log s_{T-1} - ln 2 * (sum of the
//     integer exponents) -- one log() at the end, no per-step division, and more accurate than a sum of T logs;
//   * the per-step outputs that do need a division (local evidence, row scale) are parked in the lane t mod 32 and
//     written every 32 steps by the whole warp at once (32 independent divisions instead of 32 dependent ones).
// Rescaling ON DEMAND: the factor for step t+2 is 1 unless the binary exponent e of the sum after step t has left
// [-kScaleBits, kScaleBits]; then it is 2^-e, and the sum of the next step (still uncorrected) is not looked at.  The
// compute warps skip the multiplication on the steps whose factor is 1 -- all but one in ~30 for a typical series.
constexpr int kScaleBits = 100;
__device__ __forceinline__ double ondemand_scale(double s, int &hold, int &ke) {
    ke = 0;
    if (hold > 0) {
        --hold;
        return 1.0;
    }
    const int e = ((__double2hiint(s) >> 20) & 0x7ff) - 1023;  // floor(log2 s) for a normal positive s
    if (e >= -kScaleBits && e <= kScaleBits) return 1.0;
    ke = -e;
    hold = 1;
    return __hiloint2double((1023 + ke) << 20, 0);
}

__device__ __forceinline__ double times_pow2(double x, int ke) {  // exact unless the result leaves the normal range
    return x * __hiloint2double((1023 + ke) << 20, 0);
}

struct MmaRoles {
    int lane, warp, ct;
    bool service;
    int ntw;  // tiles of this compute warp that hold grid cells (warp-uniform)
};

template <int TPW, int NT>
__device__ __forceinline__ MmaRoles mma_roles(int n) {
    constexpr int NW = NT / 32, NCW = NW - 1;
    MmaRoles r;
    r.lane = threadIdx.x & 31;
    r.warp = threadIdx.x >> 5;
    r.service = r.warp == NW - 1;  // 5 or 9 warps per CTA: the hardware's slot allocation rotates the sub-partitions
    r.ct = r.warp * 32 + r.lane;
    const int ntiles = (n + 63) >> 6;
    // warp w owns tiles w, w + NCW, w + 2 NCW ... (the partial and the absent tiles spread over the warps)
    r.ntw = r.service ? 0 : (ntiles - r.warp + NCW - 1) / NCW;
    if (r.ntw > TPW) r.ntw = TPW;
    if (r.ntw < 0) r.ntw = 0;
    return r;
}

// acc[k] = convolution outputs of the lane's cell pair in tile k (see the header); line = interior pointer of a state
// buffer, base[k] = tile_k + 8*(lane/4) + 2*(lane%4), wz = Wz + 2*(lane%4) - (lane/4) + R + kMmaWPad
template <int TPW, bool FULL>
__device__ __forceinline__ void mma_conv_body(const double *__restrict__ line, const int (&base)[TPW], int ntw, int R,
                                              const double *__restrict__ wz, double2 (&acc)[TPW]) {
#pragma unroll
    for (int k = 0; k < TPW; ++k) acc[k] = make_double2(0.0, 0.0);
    // the 8 outputs of a fragment row need the input offsets -R .. R+7: groups of 8 starting at the (even) offset
    // s0 = -R or -R-1, ceil((2R + 8 + (R & 1)) / 8) of them -- half a group fewer on average than groups aligned at
    // multiples of 8 (2 ceil(R/8) + 1)
    const int s0 = -(R + (R & 1)), groups = (2 * R + 15 + (R & 1)) >> 3;
    const double *x0 = line + base[0] + s0;  // the tiles of a warp are a compile-time distance apart
    const double *w = wz + s0;
#pragma unroll 2
    for (int j = 0; j < groups; ++j) {
        const double b0 = w[0], b1 = w[1];
        double2 av[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k)
            if (FULL || k < ntw) av[k] = *reinterpret_cast<const double2 *>(x0 + (base[k] - base[0]));
#pragma unroll
        for (int k = 0; k < TPW; ++k)  // even offsets of every tile, then the odd ones: TPW instructions between
            if (FULL || k < ntw) dmma884(acc[k].x, acc[k].y, av[k].x, b0);  // two that use the same accumulator
#pragma unroll
        for (int k = 0; k < TPW; ++k)
            if (FULL || k < ntw) dmma884(acc[k].x, acc[k].y, av[k].y, b1);
        x0 += 8;
        w += 8;
    }
}

template <int TPW>
__device__ __forceinline__ void mma_conv(const double *__restrict__ line, const int (&base)[TPW], int ntw, int R,
                                         const double *__restrict__ wz, double2 (&acc)[TPW]) {
    if (ntw == TPW)  // warp-uniform: no predicates around the matrix instructions of a warp whose tiles all hold cells
        mma_conv_body<TPW, true>(line, base, ntw, R, wz, acc);
    else
        mma_conv_body<TPW, false>(line, base, ntw, R, wz, acc);
}

// the lane's pair (v.x, v.y) = cells (c, c+1), c even, into a swizzled haloed line plus the mirror images inside the
// halo (halo <= n, n even): the mirrored pair is (c+1, c) at -2-c / 2n-2-c, again an aligned pair
__device__ __forceinline__ void mma_store_pair(double *line, int c, int n, int halo, double2 v) {
    *reinterpret_cast<double2 *>(line + swz(c)) = v;
    if (c < halo) *reinterpret_cast<double2 *>(line + swz(-2 - c)) = make_double2(v.y, v.x);
    if (c + 2 > n - halo) *reinterpret_cast<double2 *>(line + swz(2 * n - 2 - c)) = make_double2(v.y, v.x);
}

// ------------------------------------------------------------------------------------------------ K1m forward
template <int TPW, int NT, bool PROF>
__global__ void __launch_bounds__(NT, (NT > 192 && TPW > 2) ? 2 : 4) fwd_fast1d_mma_kernel(const PassArgs a) {
    constexpr int NW = NT / 32, NCW = NW - 1, NCOMP = NCW * 32;
    extern __shared__ __align__(16) double sm[];  // the base of the dynamic window is 1 KB aligned in practice
    const DevProblem &pb = a.pb;
    int slot, smid;
    const long long b = combo_and_slot(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6), slot, smid);
    if (b < 0) return;
    trace_begin(a, b);
    const int n = pb.G, halo = a.halo;
    const long long T = a.T;
    double *const buf0 = sm + halo, *const buf1 = sm + a.mma_pitch + halo;  // interior pointers, 128-byte aligned
    RedScratch rs;
    rs.buf = sm + a.off_misc;
    rs.phase = 0;
    double *const Wz = sm + a.off_w;
    const double sigma = a.pg.param[b];
    int R = a.pg.radius[b];
    const int f_lo = a.pg.window[b * 4 + 0], f_hi = a.pg.window[b * 4 + 1];
    if (!(sigma > 0.0) || R <= 0) R = 0;
    if (2 * R + 1 + 2 * kMmaWPad > a.pg.w_len[0] || mma_halo(R) > halo) {  // radius beyond blg_program.max_radius
        if (threadIdx.x == 0) {
            a.logE[b] = NAN;
            if (a.alive) a.alive[b] = -2;
            ws_pace_done(a, b);
        }
        return;
    }
    // both buffers start as zeros: the cells beyond the mirrored halo are read (with zero weights) and never written
    for (int j = threadIdx.x; j < 2 * a.mma_pitch; j += NT) sm[j] = 0.0;
    mma_build_weights(Wz, a.pg.w_len[0], sigma, R, rs);  // ends with a CTA barrier
    const MmaRoles r = mma_roles<TPW, NT>(n);
    double *PP = sm + a.ws_part;
    volatile double *ctl = sm + a.ws_ctl;
    volatile int *deadFlag = reinterpret_cast<volatile int *>(sm + a.ws_ctl + 2);
    {
        const double *init = (a.flags & BLG_F_INIT_STATE) ? a.init_state + b * (long long)n : a.prior;
        for (int g = threadIdx.x; g < n; g += NT) {
            const double v = init[g];
            buf0[swz(g)] = v;
            if (g < halo) buf0[swz(-1 - g)] = v;
            if (g >= n - halo) buf0[swz(2 * n - 1 - g)] = v;
        }
        for (int j = threadIdx.x; j < 2 * NCOMP; j += NT) PP[j] = 0.0;  // lanes without cells never write their slots
        if (threadIdx.x == 0) *deadFlag = -1;
    }
    __syncthreads();
    const bool store = !(a.flags & BLG_F_EVIDENCE_ONLY);
    const bool first = (a.flags & BLG_F_TRANSITION_FIRST) != 0;
    // a backward pass follows (it is scale-free per row): rows leave unnormalised, straight from the registers
    const bool rawRows = store && (a.flags & BLG_F_RAW_ALPHA);
    double *seq = store ? a.alpha_seq + b * a.seq_stride : nullptr;

    if (!r.service) {
        // ------------------------------------------------------------------ compute warps
        // The chains of a sub-partition run their convolutions in a common window and their epilogues in the gap behind it
        // (DESIGN.md section 6), all at once, on ONE issue port: the gap is as long as the instructions the four warps
        // execute between their last and their next matrix instruction.  Hence the shape of this loop: the tiles of a
        // warp sit a compile-time distance apart (one running pointer per array, constant offsets), guards and mirror
        // stores exist only in the warps / tiles that need them (warp-uniform flags), counters are 32-bit, the trace is a
        // template parameter.
        const int g8 = r.lane >> 2, u = r.lane & 3;
        constexpr int TS = NCW * 64;  // cells between two tiles of a warp
        const int c00 = r.warp * 64 + 8 * g8 + 2 * u;  // the lane's cell pair in the warp's first tile
        int base[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) base[k] = c00 + k * TS;
        // warp-uniform: every tile of the warp lies inside the grid, and only its first tile can touch the left halo, only
        // its last one the right halo (halo <= the distance between two tiles of a warp) -> the short store path
        const int tileFirst = r.warp * 64, tileLast = (r.warp + (TPW - 1) * NCW) * 64, cLast = c00 + (TPW - 1) * TS;
        const bool fast = tileLast + 64 <= n && halo <= TS && (TPW == 1 || tileLast - TS + 64 <= n - halo);
        const bool mirFirst = tileFirst < halo, mirLast = tileLast + 64 > n - halo;
        const double *wz = Wz + (2 * u - g8 + R + kMmaWPad);
        double *cur = buf0 + c00, *nxt = buf1 + c00;  // the lane's pair in the first tile of either buffer
        const long long pitch = a.lik_pitch;
        const double *likq = a.lik_table + c00;      // likelihood row of the NEXT step
        double *rowp = rawRows ? seq + c00 : nullptr;  // alpha row of this step
        double2 lk[TPW];  // likelihood of the lane's cell pairs, fetched one step ahead
#pragma unroll
        for (int k = 0; k < TPW; ++k)
            lk[k] = base[k] < n ? __ldg(reinterpret_cast<const double2 *>(likq + k * TS)) : make_double2(0.0, 0.0);
        likq += pitch;
        const int Ti = (int)T;
        long long cConv = 0, cEpi = 0, cBar = 0;
        double *pp = PP + r.ct;
        for (int t = 0; t < Ti; ++t) {
            double2 v[TPW];
            const bool trans = (t > 0 || first) && t - 1 >= f_lo && t - 1 < f_hi;
            // lagged scale k_t: written by the service warp before it arrived at the barrier of step t-1.  A power of two,
            // 1 on most steps (integer test of the high word: an FP64 compare would queue for the FP64 pipe); read
            // in front of the convolution, off the gap between two windows
            const double kappa = t >= 2 ? ctl[t & 1] : 1.0;
            const bool unit = __double2hiint(kappa) == 0x3ff00000;
            const long long p0 = PROF ? clock64() : 0;
            if (trans && R > 0) {
                mma_conv<TPW>(cur - c00, base, r.ntw, R, wz, v);  // transitionModels.py:111
            } else {
#pragma unroll
                for (int k = 0; k < TPW; ++k)
                    v[k] = base[k] < n ? *reinterpret_cast<const double2 *>(cur + k * TS) : make_double2(0.0, 0.0);
            }
            const long long p1 = PROF ? clock64() : 0;
            double ps[TPW];
            // alpha <- prior * likelihood (core.py:375-382); cells beyond the grid carry lik = 0
            if (unit) {
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    const double ax = v[k].x, ay = v[k].y;
                    v[k].x = ax * lk[k].x;
                    v[k].y = ay * lk[k].y;
                    ps[k] = fma(ay, lk[k].y, v[k].x);  // v.x + v.y one dependent level earlier
                }
            } else {
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    v[k].x *= kappa * lk[k].x;
                    v[k].y *= kappa * lk[k].y;
                    ps[k] = v[k].x + v[k].y;
                }
            }
            if (fast) {  // warp-uniform: every tile inside the grid, mirrors at most from the first / the last tile
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    *reinterpret_cast<double2 *>(nxt + k * TS) = v[k];
                    if (rawRows) __stcs(reinterpret_cast<double2 *>(rowp + k * TS), v[k]);
                }
                if (t + 1 < Ti) {
#pragma unroll
                    for (int k = 0; k < TPW; ++k) lk[k] = __ldg(reinterpret_cast<const double2 *>(likq + k * TS));
                }
                if (mirFirst && c00 < halo) *reinterpret_cast<double2 *>(nxt - 2 * c00 - 2) = make_double2(v[0].y, v[0].x);
                if (mirLast && cLast + 2 > n - halo)
                    *reinterpret_cast<double2 *>(nxt - c00 + 2 * n - 2 - cLast) = make_double2(v[TPW - 1].y, v[TPW - 1].x);
            } else {
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    if (base[k] < n) {
                        mma_store_pair(nxt - c00, base[k], n, halo, v[k]);
                        if (rawRows) __stcs(reinterpret_cast<double2 *>(rowp + k * TS), v[k]);
                        if (t + 1 < Ti) lk[k] = __ldg(reinterpret_cast<const double2 *>(likq + k * TS));
                    } else {
                        ps[k] = 0.0;
                    }
                }
            }
            pp[(t & 1) * NCOMP] = tree_sum<TPW>(ps);
            likq += pitch;
            if (rawRows) rowp += n;
            const long long p2 = PROF ? clock64() : 0;
            named_sync(1, NT);  // new state and its partial sums are visible to everybody
            if (PROF) {
                const long long p3 = clock64();
                cConv += p1 - p0;
                cEpi += p2 - p1;
                cBar += p3 - p2;
                const int q = t - (Ti >> 1);  // 32 steps in the middle of the series, per warp: the SM clock
                if (q >= 0 && q < 32 && r.lane == 0 && r.warp < 4) {
                    long long *e = a.trace + 36LL * gridDim.x + (((long long)blockIdx.x * 4 + r.warp) * 32 + q) * 4;
                    e[0] = p0;
                    e[1] = p1;
                    e[2] = p2;
                    e[3] = p3;
                }
            }
            {   // a zero norm found by the service warp behind the barrier of an EARLIER step ends the chain here
                const int ds = *deadFlag;
                if (ds >= 0 && ds < t) break;
            }
            double *tmp = cur;
            cur = nxt;
            nxt = tmp;
        }
        if (PROF && r.lane == 0 && r.warp < 8) {
            unsigned wid;
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
            long long *w = a.trace + 4 * (long long)gridDim.x + ((long long)blockIdx.x * 8 + r.warp) * 4;
            w[0] = cConv;
            w[1] = cEpi;
            w[2] = cBar;
            w[3] = wid;
        }
    } else {
        // ------------------------------------------------------------------ service warp
        bool dead = false;
        double sPrev = 1.0, sLast = 1.0;  // s_{t-1}; the initial state enters as it is (core.py:363, :382)
        int keNow = 0, keNext = 0;        // binary exponents of k_t, k_{t+1}
        int hold = 0;
        long long keSum = 0;              // sum of the exponents of the factors applied so far
        double myS = 1.0, myPrev = 1.0;   // lane t mod 32 parks step t for the deferred local evidence
        int myKe = 0;
        double *const local = a.local ? a.local + b * a.row_stride : nullptr;
        // local evidence of the parked steps [t0, t0 + count): norm_t = s_t / (k_t s_{t-1})   core.py:385, :404
        auto flush = [&](long long t0, int count) {
            if (local && r.lane < count) local[t0 + r.lane] = fast_div(myS, times_pow2(myPrev, myKe)) * pb.lc_prod;
        };
        WsPace pace = ws_pace_init(a, b, smid, r.lane);
        for (long long t = 0; t < T; ++t) {
            ws_pace(pace, b, t, r.lane);
            named_sync(1, NT);
            if (dead) break;  // the compute warps see the flag behind this barrier
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < NCOMP / 32; ++j) part += PP[(t & 1) * NCOMP + j * 32 + r.lane];
            const double st_sum = warp_sum(part);
            int keAfter;
            const double kAfter = ondemand_scale(st_sum, hold, keAfter);  // k_{t+2}
            if (r.lane == 0) ctl[t & 1] = kAfter;
            if (!(st_sum > 0.0) || isinf(st_sum)) {  // core.py:388-400
                dead = true;
                if (r.lane == 0) *deadFlag = (int)t;
                flush(t & ~31LL, (int)(t & 31));
                continue;  // one more barrier: the compute warps read the flag behind it
            }
            keSum += keNow;
            if (r.lane == (int)(t & 31)) {
                myS = st_sum;
                myPrev = sPrev;
                myKe = keNow;
            }
            if ((t & 31) == 31 || t == T - 1) flush(t & ~31LL, (int)(t & 31) + 1);
            const double *st = (t & 1) ? buf0 : buf1;  // the buffer the compute warps just filled
            if (store && !rawRows) {  // core.py:389, :408 -- normalised filtering distribution
                const double inv = fast_rcp(st_sum);
                double *row = seq + t * (long long)n;
                if (a.use_bulk) {  // rows are 16-byte aligned
                    for (int j = 2 * r.lane; j < n; j += 64) {
                        double2 x = *reinterpret_cast<const double2 *>(st + swz(j));
                        x.x *= inv;
                        x.y *= inv;
                        __stcs(reinterpret_cast<double2 *>(row + j), x);
                    }
                } else {
                    for (int j = r.lane; j < n; j += 32) __stcs(row + j, st[swz(j)] * inv);
                }
            }
            if ((a.flags & BLG_F_SAVE_STATE) && a.final_state && t == T - 1) {
                const double inv = fast_rcp(st_sum);
                double *fs = a.final_state + b * (long long)n;
                for (int j = r.lane; j < n; j += 32) fs[j] = st[swz(j)] * inv;
            }
            sPrev = st_sum;
            sLast = st_sum;
            keNow = keNext;
            keNext = keAfter;
        }
        if (r.lane == 0) {
            // sum_t log(norm_t) = log(s_{T-1}) - sum_t log(k_t)   (core.py:403)
            double logE = log(sLast) - (double)keSum * 0.693147180559945309417232121458;
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

// ------------------------------------------------------------------------------------------------ K2m backward
template <int TPW, int NT>
__global__ void __launch_bounds__(NT, (NT > 192 && TPW > 2) ? 2 : 4) bwd_fast1d_mma_kernel(const PassArgs a) {
    constexpr int NW = NT / 32, NCW = NW - 1, NCOMP = NCW * 32;
    extern __shared__ __align__(16) double sm[];  // the base of the dynamic window is 1 KB aligned in practice
    const DevProblem &pb = a.pb;
    int slot, smid;
    const long long b = combo_and_slot(a, reinterpret_cast<int *>(sm + a.off_misc + kMiscBarrierOffset + 6), slot, smid);
    if (b < 0) return;
    trace_begin(a, b);
    if (a.alive && a.alive[b] != 1) {  // the forward pass aborted (core.py:400)
        if (threadIdx.x == 0) ws_pace_done(a, b);
        return;
    }
    const int n = pb.G, halo = a.halo;
    const long long T = a.T;
    double *const buf0 = sm + halo, *const buf1 = sm + a.mma_pitch + halo;
    RedScratch rs;
    rs.buf = sm + a.off_misc;
    rs.phase = 0;
    double *const Wz = sm + a.off_w;
    const double sigma = a.pg.param[b];
    int R = a.pg.radius[b];
    const int b_lo = a.pg.window[b * 4 + 2], b_hi = a.pg.window[b * 4 + 3];
    if (!(sigma > 0.0) || R <= 0) R = 0;
    if (2 * R + 1 + 2 * kMmaWPad > a.pg.w_len[0] || mma_halo(R) > halo) {
        if (threadIdx.x == 0) ws_pace_done(a, b);
        return;
    }
    for (int j = threadIdx.x; j < 2 * a.mma_pitch; j += NT) sm[j] = 0.0;
    mma_build_weights(Wz, a.pg.w_len[0], sigma, R, rs);
    const MmaRoles r = mma_roles<TPW, NT>(n);
    double *PP = sm + a.ws_part;  // [2 parities][3][NCOMP]
    volatile double *ctl = sm + a.ws_ctl;
    volatile int *deadFlag = reinterpret_cast<volatile int *>(sm + a.ws_ctl + 2);
    double *seq = a.alpha_seq + b * a.seq_stride;
    const double *src = a.alpha_src ? a.alpha_src + b * a.src_stride : seq;  // filtering rows (out-of-place smoothing)
    double *const S0 = sm + a.off_stage;  // alpha[t] ring: 2 slots of Gp doubles (natural order); alpha*beta in place
    const int Gp = a.Gp;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + a.off_misc + kMiscBarrierOffset);
    const uint32_t rowBytes = (uint32_t)(n * sizeof(double));
    for (int j = threadIdx.x; j < 6 * NCOMP; j += NT) PP[j] = 0.0;
    if (threadIdx.x == 0) {
        *deadFlag = -1;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_proxy_async();
    }
    __syncthreads();
    const bool rawRows = a.row_scale != nullptr;  // BLG_F_RAW_POSTERIOR: rows leave unnormalised + their factor

    if (!r.service) {
        // ------------------------------------------------------------------ compute warps (lean loop: see the forward kernel)
        const int g8 = r.lane >> 2, u = r.lane & 3;
        constexpr int TS = NCW * 64;  // cells between two tiles of a warp
        const int c00 = r.warp * 64 + 8 * g8 + 2 * u;
        int base[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) base[k] = c00 + k * TS;
        const int tileFirst = r.warp * 64, tileLast = (r.warp + (TPW - 1) * NCW) * 64, cLast = c00 + (TPW - 1) * TS;
        const bool full = tileLast + 64 <= n;  // warp-uniform: all tiles of the warp inside the grid
        const bool fast = full && halo <= TS && (TPW == 1 || tileLast - TS + 64 <= n - halo);
        const bool mirFirst = tileFirst < halo, mirLast = tileLast + 64 > n - halo;
        const double *wz = Wz + (2 * u - g8 + R + kMmaWPad);
        double *cur = buf0 + c00, *nxt = buf1 + c00;
        const long long pitch = a.lik_pitch;
        const int Ti = (int)T;
        const double *likq = a.lik_table + (long long)(Ti - 1) * pitch + c00;  // likelihood row fetched next
        uint32_t phases = 0u;  // bit s = parity of the next completion of ring slot s
        double2 beta[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            const double v = base[k] < n ? 1.0 / (double)n : 0.0;  // core.py:424-425
            beta[k] = make_double2(v, v);
        }
        double2 lk[TPW];
#pragma unroll
        for (int k = 0; k < TPW; ++k)
            lk[k] = base[k] < n ? __ldg(reinterpret_cast<const double2 *>(likq + k * TS)) : make_double2(1.0, 1.0);
        likq -= pitch;
        double *pp = PP + r.ct;
        double *const S0c = S0 + c00;
        for (int i = Ti - 1; i >= 0; --i) {
            const int sb = i & 1;
            if (i < Ti - 1) {
                const bool trans = (i + 1 >= b_lo) && (i + 1 < b_hi);
                if (trans && R > 0) {
                    mma_conv<TPW>(cur - c00, base, r.ntw, R, wz, beta);  // transitionModels.py:117-118
                } else {
#pragma unroll
                    for (int k = 0; k < TPW; ++k)
                        if (base[k] < n) beta[k] = *reinterpret_cast<const double2 *>(cur + k * TS);
                }
                if (!full) {
#pragma unroll
                    for (int k = 0; k < TPW; ++k)
                        if (base[k] >= n) beta[k] = make_double2(0.0, 0.0);
                }
            }
            // keeps the (scale-free) beta recursion in range: a power of two, 1 on most steps (ondemand_scale)
            const double kb = i <= Ti - 3 ? ctl[i & 1] : 1.0;
            const bool unit = __double2hiint(kb) == 0x3ff00000;  // integer compare: no FP64 instruction in front of the sweep
            mbar_wait(&bars[sb], (phases >> sb) & 1u);
            phases ^= 1u << sb;
            double *A = S0c + sb * Gp;  // alpha[i], the lane's pair of the first tile
            // tile by tile with three running sums (few live registers: the gap between two convolution windows is
            // issue bound, spills and register moves are instructions too)
            double spu = 0.0, sst = 0.0, sql = 0.0;
            // one tile: posterior product into the ring slot, new state (+ mirrors), the three sums, next likelihood
            auto tile = [&](int k, auto fastPath) {
                constexpr bool FAST = decltype(fastPath)::value;
                const double2 al = *reinterpret_cast<const double2 *>(A + k * TS);
                double2 pu, st;
                pu.x = al.x * beta[k].x;  // posterior ~ alpha*beta   core.py:436
                pu.y = al.y * beta[k].y;
                *reinterpret_cast<double2 *>(A + k * TS) = pu;
                st.x = beta[k].x * lk[k].x;  // beta*likelihood          core.py:467
                st.y = beta[k].y * lk[k].y;
                if (!unit) {
                    st.x *= kb;
                    st.y *= kb;
                }
                if (FAST) {  // mirrors at most from the first / the last tile of the warp
                    *reinterpret_cast<double2 *>(nxt + k * TS) = st;
                    if (k == 0 && mirFirst && c00 < halo) *reinterpret_cast<double2 *>(nxt - 2 * c00 - 2) = make_double2(st.y, st.x);
                    if (k == TPW - 1 && mirLast && cLast + 2 > n - halo)
                        *reinterpret_cast<double2 *>(nxt - c00 + 2 * n - 2 - cLast) = make_double2(st.y, st.x);
                } else {
                    mma_store_pair(nxt - c00, base[k], n, halo, st);
                }
                // sum(post / lik), core.py:463.  Likelihoods inside the range of the hardware reciprocal seed (all but
                // pathological rows; warp-uniform test on the high words) take two Newton-corrected reciprocals
                // (the vote needs the whole warp: short path only, where no lane is masked off)
                const int lo = min(__double2hiint(lk[k].x), __double2hiint(lk[k].y));
                if (FAST && !__any_sync(0xffffffffu, lo < 0x03d00000))
                    sql += fma(pu.y, fast_rcp_pos1(lk[k].y), pu.x * fast_rcp_pos1(lk[k].x));
                else
                    sql += fast_div_pos(pu.x, lk[k].x) + fast_div_pos(pu.y, lk[k].y);
                spu += pu.x + pu.y;
                sst += st.x + st.y;
                if (i > 0) lk[k] = __ldg(reinterpret_cast<const double2 *>(likq + k * TS));
            };
            if (fast) {  // warp-uniform
#pragma unroll
                for (int k = 0; k < TPW; ++k) tile(k, std::true_type{});
            } else {
#pragma unroll
                for (int k = 0; k < TPW; ++k)
                    if (base[k] < n) tile(k, std::false_type{});
            }
            likq -= pitch;
            double *q = pp + sb * 3 * NCOMP;
            q[0] = spu;
            q[NCOMP] = sst;  // sum of the new state (magnitude control only)
            q[2 * NCOMP] = sql;
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
    } else {
        // ------------------------------------------------------------------ service warp
        const bool raw = rawRows;
        if (r.lane == 0) {
            bulk_load(S0 + ((T - 1) & 1) * Gp, src + (T - 1) * (long long)n, rowBytes, &bars[(T - 1) & 1]);
            if (T >= 2) bulk_load(S0 + ((T - 2) & 1) * Gp, src + (T - 2) * (long long)n, rowBytes, &bars[(T - 2) & 1]);
        }
        bool dead = false;
        long long i = T - 1;
        double mySpu = 1.0, mySql = 1.0;  // lane (T-1-i) mod 32 parks row i for the deferred divisions
        int hold = 0;
        double *const local = a.local ? a.local + b * a.row_stride : nullptr;
        double *const rscale = raw ? a.row_scale + b * a.row_stride : nullptr;
        // rows [iLow, iLow + count): lane j holds row iLow + count - 1 - j
        auto flush = [&](long long iLow, int count) {
            if (r.lane < count) {
                const long long row = iLow + count - 1 - r.lane;
                if (rscale) rscale[row] = fast_rcp(mySpu);  // posterior = alpha*beta / sum(alpha*beta)   core.py:439-441
                if (local) local[row] = fast_div(mySpu, mySql * pb.lc_prod);  // 1/(sum(post/lik)*lc)     core.py:463
            }
        };
        WsPace pace = ws_pace_init(a, b, smid, r.lane);
        for (; i >= 0; --i) {
            const int sb = (int)(i & 1);
            const long long done = T - 1 - i;  // rows finished so far
            ws_pace(pace, b, done, r.lane);
            named_sync(1, NT);
            if (dead) break;
            double *row = seq + i * (long long)n;
            double *P = S0 + sb * Gp;
            if (raw && r.lane == 0) {  // the unnormalised row leaves first: the copy engine reads the slot while this warp adds
                fence_proxy_async();
                bulk_store(row, P, rowBytes);
            }
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
            int ke;
            const double kAfter = ondemand_scale(sstate, hold, ke);
            if (r.lane == 0) ctl[i & 1] = kAfter;  // used by step i-2
            if (!(spu > 0.0) || !(sstate > 0.0) || isinf(sstate)) {  // core.py:440-452
                dead = true;
                if (r.lane == 0) *deadFlag = (int)i;
                flush(i + 1, (int)(done & 31));
                continue;
            }
            if (r.lane == (int)(done & 31)) {
                mySpu = spu;
                mySql = sql;
            }
            if ((done & 31) == 31 || i == 0) flush(i, (int)(done & 31) + 1);
            if (!raw) {
                const double inv = fast_rcp(spu);
                for (int j = 2 * r.lane; j < n; j += 64) {
                    double2 x = *reinterpret_cast<const double2 *>(P + j);
                    x.x *= inv;
                    x.y *= inv;
                    __stcs(reinterpret_cast<double2 *>(row + j), x);
                }
            }
            __syncwarp();
            if (r.lane == 0) {
                if (raw) bulk_wait_read<0>();
                if (i >= 2) {  // the slot is free again: prefetch alpha[i-2] into it
                    fence_proxy_async();
                    bulk_load(P, src + (i - 2) * (long long)n, rowBytes, &bars[sb]);
                }
            }
        }
        if (dead) {
            // drain the prefetch that is still in flight before the CTA's shared memory is released (see fast1d_ws.cuh)
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
