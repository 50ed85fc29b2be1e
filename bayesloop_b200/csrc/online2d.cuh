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

struct O2Geom {
    int tilesY, tilesX;  // tiles per hypothesis
    int P;               // pitch of the shared-memory buffers (doubles, odd)
    int inRowsMax;       // kTH + 2 * max R0 of the launch
    int w0len, w1len;    // padded weight table lengths (doubles)
    double *scratch;     // [H][G] unnormalised posterior cells
    double *partial;     // [H][tiles][2]: sum(v * lik), sum(v)
};

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
// load -> store round trip per row (the r1k capture: half of K7's stall samples sit on the STS behind the tile loads)
struct AsyncCopy {
    __device__ __forceinline__ void operator()(double *dst, const double *src) const {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
};

// shared memory (doubles): in[inRowsMax][P] | mid[kTH][P] | W0[w0len] | W1[w1len] | reduction scratch [4 * kMaxWarps]
// ASYNC: tile loads through cp.async (opt-in, BLG_ONLINE2D_ASYNC=1).  TH / NT: rows of a tile and threads per CTA --
// 64 / 512 (one CTA per SM) is the configuration that went through the B200 parity run; 32 / 256 (opt-in,
// BLG_ONLINE2D_TH=32) halves shared memory and registers per CTA so that two CTAs share an SM and one tile loads while
// the other convolves.  Both opt-in variants were written after the last GPU run of round 1.
template <bool ASYNC, int TH, int NT>
__global__ void __launch_bounds__(NT, TH == 64 ? 1 : 2) online2d_tile_kernel(const PassArgs a, const O2Geom geo) {
    extern __shared__ __align__(16) double sm[];
    const DevProblem &pb = a.pb;
    const int tiles = geo.tilesY * geo.tilesX;
    const long long h = blockIdx.x / tiles;
    const int tile = blockIdx.x - (int)h * tiles;
    const int ty = tile / geo.tilesX, tx = tile - ty * geo.tilesX;
    double *in = sm;
    double *mid = in + (size_t)geo.inRowsMax * geo.P;
    double *W0 = mid + (size_t)TH * geo.P;
    double *W1 = W0 + geo.w0len;
    RedScratch rs;
    rs.buf = W1 + geo.w1len;
    rs.phase = 0;

    // active operators of this hypothesis at the step handed to the models (index -1: TRANSITION_FIRST, T = 1)
    const int K = a.pg.n_ops;
    int R0 = 0, R1 = 0;
    double sig0 = 0.0, sig1 = 0.0, limit = 0.0, scale = 1.0;
    bool clamp = false, reset = false;
    for (int k = 0; k < K; ++k) {
        const int lo = a.pg.window[(h * K + k) * 4 + 0], hi = a.pg.window[(h * K + k) * 4 + 1];
        if (-1 < lo || -1 >= hi) continue;
        const int kind = a.pg.kind[k];
        const double par = a.pg.param[h * K + k];
        if (kind == BLG_OP_GRW) {
            const int R = a.pg.radius[h * K + k];
            if (!(par > 0.0) || R <= 0) continue;  // transitionModels.py:110-113
            if (a.pg.axis[k] == 0) {
                R0 = R;
                sig0 = par;
            } else {
                R1 = R;
                sig1 = par;
            }
        } else if (kind == BLG_OP_REGIME) {
            clamp = true;
            limit = par;
        } else if (kind == BLG_OP_RESET) {
            reset = true;
            scale = par;
        }
    }

    o2::Tile<TH> t;
    t.n0 = pb.n0;
    t.n1 = pb.n1;
    t.r0 = ty * TH;
    t.c0 = tx * o2::kTW;
    t.R0 = R0;
    t.R1 = R1;
    t.P = geo.P;
    O2Lik lik{a, {pb.tabA[0], pb.tabA[1], pb.tabA[2], pb.tabB[0], pb.tabB[1]}, a.steps};
    const double *src = a.init_state + h * (long long)pb.G;
    double *dst = geo.scratch + h * (long long)pb.G;
    double s1 = 0.0, s2 = 0.0;
    const int tid = threadIdx.x, nt = blockDim.x;

    if (reset || (R0 == 0 && R1 == 0)) {
        if (reset) clamp = false;
        o2::pointwise_phase(t, src, reset ? a.reset_base : nullptr, scale, dst, clamp, limit, lik, tid, nt, s1, s2);
    } else {
        // a radius beyond the tables the host sized (blg_program.max_radius is a promise): poison this hypothesis
        const bool fits = o2::padded_taps(R0, o2::kM0) <= geo.w0len && o2::padded_taps(R1, o2::kM1) <= geo.w1len &&
                          t.inRows() <= geo.inRowsMax && t.inCols() <= geo.P;
        if (!fits) {
            s2 = NAN;
        } else {
            if (ASYNC) {
                o2::load_phase(t, src, in, tid, nt, AsyncCopy());
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else {
                o2::load_phase(t, src, in, tid, nt);
            }
            if (R0 > 0) {
                build_weights(W0, o2::padded_taps(R0, o2::kM0), sig0, R0, rs);
            } else {
                for (int j = tid; j < o2::kM0; j += nt) W0[j] = j == 0 ? 1.0 : 0.0;
            }
            if (R1 > 0) {
                build_weights(W1, o2::padded_taps(R1, o2::kM1), sig1, R1, rs);
            } else {
                for (int j = tid; j < o2::kM1; j += nt) W1[j] = j == 0 ? 1.0 : 0.0;
            }
            if (ASYNC) asm volatile("cp.async.wait_group 0;" ::: "memory");  // the weights were built behind the loads
            __syncthreads();
            o2::conv0_phase(t, in, mid, W0, tid, nt);
            __syncthreads();
            o2::conv1_phase(t, mid, in, W1, tid, nt);  // the haloed input is dead: its buffer takes the output tile
            __syncthreads();
            o2::epilogue_phase(t, in, dst, clamp, limit, lik, tid, nt, s1, s2);
        }
    }
    block_sum2(s2, s1, rs);
    if (threadIdx.x == 0) {
        double *p = geo.partial + ((size_t)h * tiles + tile) * 2;
        p[0] = s2;
        p[1] = clamp ? s1 : (tile == 0 ? 1.0 : 0.0);  // no clamp: the transitioned prior is used as it is (S1 = 1)
    }
}

// grid (chunks, H): every block re-derives the sums of its hypothesis from the per-tile partials in the same order.
__global__ void __launch_bounds__(256) online2d_finish_kernel(const PassArgs a, const O2Geom geo) {
    const long long h = blockIdx.y;
    const int tiles = geo.tilesY * geo.tilesX;
    const double *p = geo.partial + (size_t)h * tiles * 2;
    double s2 = 0.0, s1 = 0.0;
    for (int k = 0; k < tiles; ++k) {
        s2 += p[2 * k];
        s1 += p[2 * k + 1];
    }
    const double norm = s2 / s1;
    const bool dead = !(norm > 0.0) || !(s2 > 0.0) || isinf(norm);  // core.py:2171 has no guard; the engine reports it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double logE = dead ? -INFINITY : log(norm);
        if (!dead && !(a.flags & BLG_F_INIT_STATE)) logE += log(a.pb.lc_prod);
        a.logE[h] = logE;
        if (a.local && !dead) a.local[h * a.T] = norm * a.pb.lc_prod;
        if (a.alive) a.alive[h] = dead ? 0 : 1;
    }
    if (dead) return;  // like the persistent kernels: the state of a dead hypothesis is left untouched
    const double inv = 1.0 / s2;
    const long long G = a.pb.G;
    const double *src = geo.scratch + h * G;
    double *dst = a.final_state + h * G;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (long long)gridDim.x * blockDim.x)
        dst[g] = src[g] * inv;
}

}  // namespace blg
