// online2d_phases.h -- the per-thread phases of the tiled OnlineStudy step (online2d.cuh), written as plain
// host/device functions of (thread id, thread count) so that the SAME index arithmetic runs in the sm_100a kernel
// (tid = threadIdx.x, a __syncthreads() between phases) and in the CPU emulation tools/emu/online2d_emu.cpp
// (a loop over tid per phase), which checks it against a direct reflect convolution without a GPU.
//
// One tile = TH x kTW cells of one hypothesis (TH = 64).  Reference semantics: the transition of OnlineStudy.step
// (bayesloop/core.py:2166 -> transitionModels.py:96-115, gaussian_filter1d along each axis, mode='reflect') followed by
// posterior = prior * likelihood (core.py:2170).
#pragma once

#ifdef __CUDACC__
#define BLG_HD __host__ __device__ __forceinline__
#else
#define BLG_HD inline
#endif

namespace blg {
namespace o2 {

constexpr int kTH = 64;        // rows of a tile (default; the functions below take the rows as template parameter TH)
constexpr int kTW = 64;        // columns of a tile
constexpr int kM0 = 16;        // rows per work item of the axis-0 convolution
constexpr int kM1 = 8;         // cells per work item of the axis-1 convolution
constexpr int kThreads = 512;

BLG_HD int reflect(int i, int n) {  // NI_EXTEND_REFLECT for any index (d c b a | a b c d | d c b a)
    if ((unsigned)i < (unsigned)n) return i;
    const int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m >= n ? p - 1 - m : m;
}

BLG_HD int padded_taps(int R, int M) { return (2 * R + 1 + M - 1) / M * M; }  // weight table length (zero padded)

template <int TH = kTH>
struct Tile {
    int n0, n1;  // grid
    int r0, c0;  // first row / column of the tile
    int R0, R1;  // radii of this hypothesis; 0 = no convolution along that axis (single tap of weight 1)
    int P;       // pitch (doubles) of the shared-memory buffers: odd, >= kTW + 2 * max R1 of the launch
    BLG_HD int inRows() const { return TH + 2 * R0; }
    BLG_HD int inCols() const { return kTW + 2 * R1; }
};

// Phase L: the tile with its halo (R0 rows above / below, R1 columns left / right, reflected at the grid edges; tiles
// that stick out of the grid read reflected cells too -- finite values that the epilogue masks) -> in[inRows][P].
// `copy(dst, src)` moves one cell: a plain load + store, or an asynchronous global->shared copy in the kernel.
struct PlainCopy {
    BLG_HD void operator()(double *dst, const double *src) const { *dst = *src; }
};

template <int TH, class Copy = PlainCopy>
BLG_HD void load_phase(const Tile<TH> &t, const double *src, double *in, int tid, int nt, Copy copy = Copy()) {
    const int rows = t.inRows(), cols = t.inCols();
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;  // one row per warp: coalesced, no integer division
    for (int i = warp; i < rows; i += nw) {
        const double *line = src + (long long)reflect(t.r0 - t.R0 + i, t.n0) * t.n1;
        double *row = in + i * t.P;
        for (int j = lane; j < cols; j += 32) copy(row + j, line + reflect(t.c0 - t.R1 + j, t.n1));
    }
}

// "Valid" correlation of nLines lines: dst[l][i] = sum_{k < taps} W[k] * src[l][i + k] for i < nOut.  Work item = M
// consecutive outputs of one line; the inputs stream through a register window (one load + one weight per M FMAs).
// W holds padded_taps entries (zeros behind the taps); input indices are clamped to lineLen - 1, which only the zero
// weights ever reach.  Consecutive threads take consecutive LINES (odd pitch => conflict-free 64-bit accesses).
template <int M>
BLG_HD void conv_valid(const double *src, double *dst, const double *W, int taps, int nOut, int lineLen, int elemStride,
                       int nLines, int lineStride, int dElemStride, int dLineStride, int tid, int nt) {
    const int S = (nOut + M - 1) / M;
    const int nItems = S * nLines;
    const int last = lineLen - 1;
    for (int w = tid; w < nItems; w += nt) {
        const int s = w / nLines, l = w - s * nLines;
        const double *line = src + l * lineStride;
        const int i0 = s * M;
        int idx = i0;
        double win[M], acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int q = idx + m < last ? idx + m : last;
            win[m] = line[q * elemStride];
            acc[m] = 0.0;
        }
        idx += M;
        // whole chunks of M taps, then the remaining taps behind uniform guards (a zero-padded last chunk would spend
        // up to M - 1 useless FMAs per output: 25 % of the work at the radii of BASELINE.json configs[4])
        const int full = taps / M, rem = taps - full * M;
        int j0 = 0;
        for (int c = 0; c < full; ++c, j0 += M) {
#pragma unroll
            for (int u = 0; u < M; ++u) {
                const double wt = W[j0 + u];
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = fma(wt, win[(u + m) % M], acc[m]);
                const int q = idx < last ? idx : last;
                win[u] = line[q * elemStride];
                ++idx;
            }
        }
#pragma unroll
        for (int u = 0; u < M - 1; ++u) {
            if (u < rem) {
                const double wt = W[j0 + u];
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = fma(wt, win[(u + m) % M], acc[m]);
                const int q = idx < last ? idx : last;
                win[u] = line[q * elemStride];
                ++idx;
            }
        }
        double *out = dst + l * dLineStride;
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (i0 + m < nOut) out[(i0 + m) * dElemStride] = acc[m];
    }
}

// Phase A: axis-0 convolution of every column of the haloed tile: in[inRows][P] -> mid[TH][P] (inCols columns).
template <int TH>
BLG_HD void conv0_phase(const Tile<TH> &t, const double *in, double *mid, const double *W0, int tid, int nt) {
    conv_valid<kM0>(in, mid, W0, 2 * t.R0 + 1, TH, t.inRows(), t.P, t.inCols(), 1, t.P, 1, tid, nt);
}

// Phase B: axis-1 convolution of every row: mid[TH][P] -> out[TH][kOutP] (kTW columns).  `out` is a buffer of its own:
// the haloed input buffer is already receiving the NEXT tile while this phase and the epilogue run.
constexpr int kOutP = kTW + 1;  // odd pitch: consecutive threads take consecutive rows => conflict-free 64-bit stores

template <int TH>
BLG_HD void conv1_phase(const Tile<TH> &t, const double *mid, double *out, const double *W1, int tid, int nt) {
    conv_valid<kM1>(mid, out, W1, 2 * t.R1 + 1, kTW, t.inCols(), 1, TH, t.P, 1, kOutP, tid, nt);
}

// Phase E: v = transitioned prior of the cell (optionally clamped from below: RegimeSwitch, transitionModels.py:405),
// u = v * likelihood -> dst (global, unnormalised); per-thread partial sums s1 += v, s2 += u.  Consecutive threads
// take consecutive cells of a row: coalesced likelihood loads and stores.
// `lik(gi, gj, g)` returns the likelihood of grid cell (gi, gj), g = gi * n1 + gj.
template <int TH, class Lik>
BLG_HD void epilogue_phase(const Tile<TH> &t, const double *out, double *dst, bool clamp, double limit, Lik lik, int tid, int nt,
                           double &s1, double &s2) {
    // batches of 4 cells per thread: the likelihood requests of a batch are all in flight before the first product
    // (one request -> multiply -> store at a time left 12 % of the kernel's stall samples on that multiply)
    constexpr int B = 4;
    for (int e0 = tid; e0 < TH * kTW; e0 += B * nt) {
        double lk[B], v[B];
        long long g[B];
        bool ok[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int e = e0 + b * nt;
            const int i = e / kTW, j = e - i * kTW;
            const int gi = t.r0 + i, gj = t.c0 + j;
            ok[b] = e < TH * kTW && gi < t.n0 && gj < t.n1;
            g[b] = (long long)gi * t.n1 + gj;
            lk[b] = ok[b] ? lik(gi, gj, g[b]) : 0.0;
            v[b] = ok[b] ? out[i * kOutP + j] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (ok[b]) {
                double x = v[b];
                if (clamp) x = x < limit ? limit : x;
                const double u = x * lk[b];
                dst[g[b]] = u;
                s1 += x;
                s2 += u;
            }
    }
}

// Hypotheses without a convolution (Static, RegimeSwitch alone, Independent / reset): the same epilogue straight from
// global memory.  reset != nullptr: v = reset[g] * scale (transitionModels.py:350-359), else v = src[g].
template <int TH, class Lik>
BLG_HD void pointwise_phase(const Tile<TH> &t, const double *src, const double *reset, double scale, double *dst, bool clamp,
                            double limit, Lik lik, int tid, int nt, double &s1, double &s2) {
    constexpr int B = 4;  // batches as in epilogue_phase: all global requests of a batch in flight together
    for (int e0 = tid; e0 < TH * kTW; e0 += B * nt) {
        double lk[B], v[B];
        long long g[B];
        bool ok[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int e = e0 + b * nt;
            const int i = e / kTW, j = e - i * kTW;
            const int gi = t.r0 + i, gj = t.c0 + j;
            ok[b] = e < TH * kTW && gi < t.n0 && gj < t.n1;
            g[b] = (long long)gi * t.n1 + gj;
            lk[b] = ok[b] ? lik(gi, gj, g[b]) : 0.0;
            v[b] = ok[b] ? (reset ? reset[g[b]] * scale : src[g[b]]) : 0.0;
        }
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (ok[b]) {
                double x = v[b];
                if (clamp) x = x < limit ? limit : x;
                const double u = x * lk[b];
                dst[g[b]] = u;
                s1 += x;
                s2 += u;
            }
    }
}

}  // namespace o2
}  // namespace blg
