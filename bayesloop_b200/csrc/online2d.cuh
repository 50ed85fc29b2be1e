// online2d.cuh -- K7/K8: one OnlineStudy step (core.py:2157-2175) for 2-D grids that do not fit in shared memory,
// tiled over the whole GPU instead of one persistent CTA per hypothesis.
//
// The stream kernels (resident.cuh, STREAM) walk one hypothesis per CTA through global scratch buffers: correct for
// any operator program, but a step of 256 hypotheses on a 512 x 512 grid leaves them at ~6 % of the HBM roofline
// (2.9 ms for 1.07 GB of compulsory traffic).  A step with T = 1 has no time recursion to keep on chip, so here every
// hypothesis is cut into 64 x 64 tiles and the grid is (tiles x hypotheses):
//
//   K7 online2d_tile_kernel    load tile + halo (reflected at the grid edges) -> axis-0 convolution of all haloed
//                              columns -> axis-1 convolution -> (RegimeSwitch clamp) -> x likelihood -> UNNORMALISED
//                              cells to a global scratch [H][G], two partial sums per tile (sum v, sum v*lik)
//   K8 online2d_finish_kernel  per hypothesis: partial sums in fixed order -> norm, log-evidence increment, alive;
//                              final_state = scratch / sum (the scratch also decouples the halo reads of K7 from
//                              the write-back: init_state and final_state are the same buffer in OnlineStudy.step)
//
// Normalisation is lazy: p <- max(T(p), limit) / S1, alpha = p * lik, norm = sum(alpha) = S2 / S1 and the stored
// posterior is alpha / norm = (v * lik) / S2 -- one pass over the cells, no grid-wide barrier.
// Rows must satisfy the caller's promise BLG_F_SEPARABLE_ROWS (include/blgrid.h): active operators of a row are
// GaussianRandomWalks on distinct axes, optionally followed by one RegimeSwitch, or a single reset.
// Algorithmic HBM bytes per cell: K7 8 read + 8 written, K8 8 read + 8 written.
#pragma once

#include "common.cuh"
#include "online2d_phases.h"

namespace blg {

// per-hypothesis descriptor of one step, resolved once per launch by online2d_desc_kernel (the tile CTAs used to
// walk the operator program in global memory themselves: a chain of dependent loads at the start of every tile)
struct O2Hyp {
    int R0, R1;   // radii of the active random walks (0: none along that axis)
    int mode;     // 0 convolution, 1 pointwise (state as it is), 2 reset
    int clamp;    // RegimeSwitch lower bound active
    double sig0, sig1, limit, scale;
};

struct O2Geom {
    int tilesY, tilesX;  // tiles per hypothesis
    int P;               // pitch of the shared-memory buffers (doubles, odd)
    int inRowsMax;       // kTH + 2 * max R0 of the launch
    int w0len, w1len;    // padded weight table lengths (doubles)
    double *scratch;     // [H][G] unnormalised posterior cells
    double *partial;     // [H][tiles][2]: sum(v * lik), sum(v)
    O2Hyp *hyp;          // [H]
    int *counter;        // work queue: next unit (zeroed before the launch)
    int chunk;           // tiles per unit: kO2Chunk, halved while the launch has fewer than ~6 units per CTA (few hypotheses
                         // per GPU under sharding: the tail of the queue would leave SMs idle)
    int pipelined;       // 1: the output tile has a shared-memory buffer of its own, so the next tile's loads overlap
                         //    the axis-1 pass and the epilogue; 0 (radii too wide for that): it reuses `in`
};

constexpr int kO2Chunk = 8;  // most tiles of one hypothesis handed out per queue access (weights are built once per unit)

struct O2Lik {
    const PassArgs &a;
    LikTables tb;
    const StepC *sc;
    __device__ __forceinline__ double operator()(int gi, int gj, long long g) const {
        if (a.pb.om_kind == BLG_OM_TABLE) return __ldg(a.lik_table + g);
        return lik_cell(a.pb, tb, sc, gi, gj);
    }
};

// 8-byte asynchronous global->shared copy (LDGSTS): the loads of a thread are all in flight at once instead of one
// load -> store round trip per row (the r1k capture: half of K7's stall samples sat on the STS behind the tile loads)
struct AsyncCopy {
    __device__ __forceinline__ void operator()(double *dst, const double *src) const {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
};

// active operators of every hypothesis at the step handed to the models (index -1: TRANSITION_FIRST, T = 1)
__global__ void online2d_desc_kernel(const PassArgs a, O2Hyp *out) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= a.B) return;
    const int K = a.pg.n_ops;
    O2Hyp d;
    d.R0 = d.R1 = 0;
    d.sig0 = d.sig1 = d.limit = 0.0;
    d.scale = 1.0;
    d.clamp = 0;
    bool reset = false;
    for (int k = 0; k < K; ++k) {
        const int lo = a.pg.window[(h * K + k) * 4 + 0], hi = a.pg.window[(h * K + k) * 4 + 1];
        if (-1 < lo || -1 >= hi) continue;
        const int kind = a.pg.kind[k];
        const double par = a.pg.param[h * K + k];
        if (kind == BLG_OP_GRW) {
            const int R = a.pg.radius[h * K + k];
            if (!(par > 0.0) || R <= 0) continue;  // transitionModels.py:110-113
            if (a.pg.axis[k] == 0) {
                d.R0 = R;
                d.sig0 = par;
            } else {
                d.R1 = R;
                d.sig1 = par;
            }
        } else if (kind == BLG_OP_REGIME) {
            d.clamp = 1;
            d.limit = par;
        } else if (kind == BLG_OP_RESET) {
            reset = true;
            d.scale = par;
        }
    }
    if (reset) d.clamp = 0;
    d.mode = reset ? 2 : ((d.R0 == 0 && d.R1 == 0) ? 1 : 0);
    out[h] = d;
}

// shared memory (doubles): in[inRowsMax][P] | mid[kTH][P] | out[kTH][kOutP] | W0[w0len] | W1[w1len] | reduction
// scratch [4 * kMaxWarps]
//
// PERSISTENT, PIPELINED (round 2): one CTA per SM takes units of kO2Chunk consecutive tiles of one hypothesis from an
// atomic queue (cost per tile varies with the radii).  Per unit the weights are built once; per tile
//     wait for the haloed input tile (cp.async) -> axis-0 convolution in -> mid -> barrier
//     -> issue the NEXT tile's loads into `in` -> axis-1 convolution mid -> out -> barrier -> epilogue (clamp,
//        likelihood multiply, coalesced global store, partial sums)
// so the tile loads overlap the second convolution and the epilogue.  (A variant that fused the epilogue into the
// axis-1 write-back from registers was measured and dropped: its per-row scattered 8-byte global accesses cost what
// the saved shared-memory pass gained, 1.33 ms per C5 step either way.)
// ASYNC = false keeps the plain LDG -> STS loads (no overlap; the reference point of profiles/r2a_online_ab.txt).
template <bool ASYNC>
__global__ void __launch_bounds__(o2::kThreads, 1) online2d_tile_kernel(const PassArgs a, const O2Geom geo) {
    extern __shared__ __align__(16) double sm[];
    constexpr int TH = o2::kTH;
    __shared__ int unitSh;
    const DevProblem &pb = a.pb;
    const int tiles = geo.tilesY * geo.tilesX;
    const int chunk = geo.chunk;
    const int chunksPerHyp = (tiles + chunk - 1) / chunk;
    const long long units = a.B * (long long)chunksPerHyp;
    double *in = sm;
    double *mid = in + (size_t)geo.inRowsMax * geo.P;
    double *own = mid + (size_t)TH * geo.P;
    double *out = geo.pipelined ? own : in;
    double *W0 = own + (geo.pipelined ? (size_t)TH * o2::kOutP : 0);
    double *W1 = W0 + geo.w0len;
    RedScratch rs;
    rs.buf = W1 + geo.w1len;
    rs.phase = 0;
    O2Lik lik{a, {pb.tabA[0], pb.tabA[1], pb.tabA[2], pb.tabB[0], pb.tabB[1]}, a.steps};
    const int tid = threadIdx.x, nt = blockDim.x;

    for (;;) {
        __syncthreads();  // the previous unit is done with shared memory (and with unitSh)
        if (tid == 0) unitSh = atomicAdd(geo.counter, 1);
        __syncthreads();
        const long long u = unitSh;
        if (u >= units) break;
        const long long h = u / chunksPerHyp;
        const int first = (int)(u - h * chunksPerHyp) * chunk;
        const int count = min(chunk, tiles - first);
        const O2Hyp hp = geo.hyp[h];
        const bool clamp = hp.clamp != 0;
        const double *src = a.init_state + h * (long long)pb.G;
        double *dst = geo.scratch + h * (long long)pb.G;
        double *part = geo.partial + ((size_t)h * tiles + first) * 2;

        o2::Tile<TH> t;
        t.n0 = pb.n0;
        t.n1 = pb.n1;
        t.R0 = hp.R0;
        t.R1 = hp.R1;
        t.P = geo.P;
        auto place = [&](int k) {
            const int tile = first + k, ty = tile / geo.tilesX;
            t.r0 = ty * TH;
            t.c0 = (tile - ty * geo.tilesX) * o2::kTW;
        };
        auto publish = [&](int k, double s1, double s2) {  // per-tile partial sums; every thread calls (block reduction)
            block_sum2(s2, s1, rs);
            if (tid == 0) {
                part[2 * k] = s2;
                // no clamp: the transitioned prior is used as it is (S1 = 1, carried by the hypothesis' first tile)
                part[2 * k + 1] = clamp ? s1 : (first + k == 0 ? 1.0 : 0.0);
            }
        };

        if (hp.mode != 0) {
            for (int k = 0; k < count; ++k) {
                place(k);
                double s1 = 0.0, s2 = 0.0;
                o2::pointwise_phase(t, src, hp.mode == 2 ? a.reset_base : nullptr, hp.scale, dst, clamp, hp.limit, lik, tid,
                                    nt, s1, s2);
                publish(k, s1, s2);
            }
            continue;
        }
        // a radius beyond the tables the host sized (blg_program.max_radius is a promise): poison this hypothesis
        const bool fits = o2::padded_taps(hp.R0, o2::kM0) <= geo.w0len && o2::padded_taps(hp.R1, o2::kM1) <= geo.w1len &&
                          t.inRows() <= geo.inRowsMax && t.inCols() <= geo.P;
        if (!fits) {
            if (tid == 0)
                for (int k = 0; k < count; ++k) {
                    part[2 * k] = NAN;
                    part[2 * k + 1] = 1.0;
                }
            continue;
        }
        place(0);
        if (ASYNC) {
            o2::load_phase(t, src, in, tid, nt, AsyncCopy());
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            o2::load_phase(t, src, in, tid, nt);
        }
        if (hp.R0 > 0) {
            build_weights(W0, o2::padded_taps(hp.R0, o2::kM0), hp.sig0, hp.R0, rs);
        } else {
            for (int j = tid; j < o2::kM0; j += nt) W0[j] = j == 0 ? 1.0 : 0.0;
        }
        if (hp.R1 > 0) {
            build_weights(W1, o2::padded_taps(hp.R1, o2::kM1), hp.sig1, hp.R1, rs);
        } else {
            for (int j = tid; j < o2::kM1; j += nt) W1[j] = j == 0 ? 1.0 : 0.0;
        }
        for (int k = 0; k < count; ++k) {
            if (ASYNC) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();  // input tile (and, for k = 0, the weights) visible to everybody
            o2::conv0_phase(t, in, mid, W0, tid, nt);
            __syncthreads();  // `in` is dead: the next tile may land in it while this one is finished out of `mid`
            const o2::Tile<TH> cur = t;
            auto loadNext = [&]() {
                if (k + 1 < count) {
                    place(k + 1);
                    if (ASYNC) {
                        o2::load_phase(t, src, in, tid, nt, AsyncCopy());
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    } else {
                        o2::load_phase(t, src, in, tid, nt);
                    }
                }
            };
            if (geo.pipelined) loadNext();
            o2::conv1_phase(cur, mid, out, W1, tid, nt);
            __syncthreads();
            double s1 = 0.0, s2 = 0.0;
            o2::epilogue_phase(cur, out, dst, clamp, hp.limit, lik, tid, nt, s1, s2);
            publish(k, s1, s2);  // its barrier also orders this tile's reads of `out` before the next axis-1 pass
            if (!geo.pipelined) loadNext();  // `out` aliases `in`: the next tile can only land now
        }
    }
}

// K8a: one warp per hypothesis adds the per-tile partial sums in a fixed order (lane-strided, then a shuffle tree:
// deterministic), writes the evidence increment / alive flag and leaves 1 / sum(u) for K8b (0: dead, state untouched).
// (Round 1 let every block of the streaming kernel re-derive the sums with a serial loop over the tiles: ~10 us of
// dependent L2 loads in front of every block made K8 latency bound, 335 us for 1.07 GB.)
__global__ void __launch_bounds__(128) online2d_sums_kernel(const PassArgs a, const O2Geom geo, double *__restrict__ inv) {
    const long long h = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (h >= a.B) return;
    const int tiles = geo.tilesY * geo.tilesX;
    const double *p = geo.partial + (size_t)h * tiles * 2;
    double s2 = 0.0, s1 = 0.0;
    for (int k = lane; k < tiles; k += 32) {
        s2 += p[2 * k];
        s1 += p[2 * k + 1];
    }
    s2 = warp_sum(s2);
    s1 = warp_sum(s1);
    if (lane != 0) return;
    const double norm = s2 / s1;
    const bool dead = !(norm > 0.0) || !(s2 > 0.0) || isinf(norm);  // core.py:2171 has no guard; the engine reports it
    double logE = dead ? -INFINITY : log(norm);
    if (!dead && !(a.flags & BLG_F_INIT_STATE)) logE += log(a.pb.lc_prod);
    a.logE[h] = logE;
    if (a.local && !dead) a.local[h * a.row_stride] = norm * a.pb.lc_prod;
    if (a.alive) a.alive[h] = dead ? 0 : 1;
    inv[h] = dead ? 0.0 : 1.0 / s2;
}

// K8b, grid (chunks, H): final_state = scratch / sum(u), streaming; like the persistent kernels the state of a dead
// hypothesis is left untouched.
__global__ void __launch_bounds__(256) online2d_finish_kernel(const PassArgs a, const O2Geom geo, const double *__restrict__ inv) {
    const long long h = blockIdx.y;
    const double f = __ldg(inv + h);
    if (!(f > 0.0)) return;
    const long long G = a.pb.G;
    const double *src = geo.scratch + h * G;
    double *dst = a.final_state + h * G;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (long long)gridDim.x * blockDim.x)
        dst[g] = __ldcs(src + g) * f;
}

}  // namespace blg
