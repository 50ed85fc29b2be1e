// api.cu -- host side of libblgrid.so: the C ABI of include/blgrid.h on top of the sm_100a kernels.
//
// Boundary replaced (reference file:line): the per-time-step Python loops of Study.fit (bayesloop/core.py:372-411,
// :434-470), the per-combination loop and averaging of HyperStudy.fit (core.py:1349-1366, :1379-1382, :1410) and
// the per-hypothesis loop of OnlineStudy.step (core.py:2157-2175, :2195-2212).  No torch types cross this file; the
// caller hands raw device pointers and a cudaStream_t.
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "aux.cuh"
#include "kernels.h"

using namespace blg;

namespace {

thread_local char g_err[512];
std::atomic<long long> g_launches{0};
const char *g_last_kernel = "none";

int fail(const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return -1;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            snprintf(g_err, sizeof g_err, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

inline int even_up(int x) { return (x + 1) & ~1; }

// Dispatch / tuning options of a plan.  Resolved ONCE: blg_plan_create reads the BLG_* environment variables
// (debugging aids of tools/), blg_plan_set_option overrides them by name; no call on the per-step path looks at the
// process environment, so which kernel family runs is a property of the plan, not of process state.
struct Opts {
    int no_fast1d = 0;       // never take the fused 1-D kernels (fast1d.cuh / fast1d_ws.cuh)
    int no_ws = 0;           // fused 1-D kernels: single-role variant instead of the warp-specialised one
    int no_bulk = 0;         // no bulk-async (TMA) row copies
    int no_lik_table = 0;    // evaluate the likelihood in the passes instead of one shared [T][G] table per call
    int no_order = 0;        // ignore blg_program.order
    int no_sm_assign = 0;    // ignore blg_program.sm_assign
    int force_stream = 0;    // global-memory stream kernels even where the state fits on chip
    int cluster2d = 0;       // cluster-resident 2-D kernels even where one SM's shared memory would do
    int no_cluster2d = 0;    // never take the cluster-resident 2-D kernels
    int cluster2d_c = 0;     // force the cluster size (2, 4, 8); 0 = largest that fits
    int online2d = 1;        // tiled OnlineStudy step for 2-D grids beyond shared memory (0: stream kernels)
    int online2d_small = 0;  // tiled OnlineStudy step also on grids that would fit in shared memory
    int online2d_async = 1;  // tile loads through cp.async (0: LDG -> STS round trips)
    int serpentine = 1;      // alternate the block -> combo mapping per wave of SMs
    int verbose = 0;         // print the launch geometry to stderr
    int ws_m = 0, ws_nt = 0; // warp-specialised 1-D kernels: force (cells per thread, threads); 0 = by grid size
    int ws_ml = 0;           // ... cells per thread of the last compute warp (0 = ws_m)
    int ws_even = 0;         // ... same number of cells per thread in every compute warp (no uneven split)
    int mma_nt = 0;          // ... 288: 8 compute warps per chain where 4 would do (0 = by grid size)
    int il = 0;              // ... the chains of an SM interleaved in one CTA (fast1d_il.cuh; measured slower, opt-in)
    int no_mma = 0;          // ... convolution with DFMAs (fast1d_ws.cuh) instead of FP64 matrix instructions (fast1d_mma.cuh)
    int ws_pace_every = 4;   // ... chains of an SM publish their step count every N steps (power of two) ...
    int ws_pace_skew = 8;    // ... and hold back when more than this many steps ahead of a peer (0 = no pacing)
    char trace[256] = {0};   // per-CTA trace files <trace>.<kernel>.<n>.csv (debugging aid)
};

struct OptName {
    const char *name, *env;
    int Opts::*field;
};
const OptName kOptNames[] = {
    {"no_fast1d", "BLG_NO_FAST1D", &Opts::no_fast1d},
    {"no_ws", "BLG_NO_WS", &Opts::no_ws},
    {"no_bulk", "BLG_NO_BULK", &Opts::no_bulk},
    {"no_lik_table", "BLG_NO_LIK_TABLE", &Opts::no_lik_table},
    {"no_order", "BLG_NO_ORDER", &Opts::no_order},
    {"no_sm_assign", "BLG_NO_SM_ASSIGN", &Opts::no_sm_assign},
    {"force_stream", "BLG_FORCE_STREAM", &Opts::force_stream},
    {"cluster2d", "BLG_CLUSTER2D", &Opts::cluster2d},
    {"no_cluster2d", "BLG_NO_CLUSTER2D", &Opts::no_cluster2d},
    {"cluster2d_c", "BLG_CLUSTER2D_C", &Opts::cluster2d_c},
    {"online2d", "BLG_ONLINE2D", &Opts::online2d},
    {"online2d_small", "BLG_ONLINE2D_SMALL", &Opts::online2d_small},
    {"online2d_async", "BLG_ONLINE2D_ASYNC", &Opts::online2d_async},
    {"serpentine", "BLG_SERPENTINE", &Opts::serpentine},
    {"verbose", "BLG_VERBOSE", &Opts::verbose},
    {"ws_m", "BLG_WS_M", &Opts::ws_m},
    {"ws_nt", "BLG_WS_NT", &Opts::ws_nt},
    {"ws_ml", "BLG_WS_ML", &Opts::ws_ml},
    {"ws_even", "BLG_WS_EVEN", &Opts::ws_even},
    {"no_mma", "BLG_NO_MMA", &Opts::no_mma},
    {"il", "BLG_IL", &Opts::il},
    {"mma_nt", "BLG_MMA_NT", &Opts::mma_nt},
    {"ws_pace_every", "BLG_WS_PACE_EVERY", &Opts::ws_pace_every},
    {"ws_pace_skew", "BLG_WS_PACE_SKEW", &Opts::ws_pace_skew},
};

void opts_from_env(Opts &o) {
    for (const OptName &n : kOptNames)
        if (const char *e = getenv(n.env)) {
            char *end = nullptr;
            const long v = strtol(e, &end, 10);
            o.*(n.field) = (end == e) ? 1 : (int)v;  // "BLG_X=" or "BLG_X=yes" switch a flag on
        }
    if (const char *e = getenv("BLG_TRACE")) snprintf(o.trace, sizeof o.trace, "%s", e);
}

constexpr int kMiscDoubles = 384;  // reduction scratch (128) + params (16) + radius/window ints (40) + 2 mbarriers
                                   // + per-warp partial sums of the fast 1-D kernels (192)
constexpr size_t kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

}  // namespace

struct blg_plan {
    blg_problem pb;
    DevProblem dev;
    int device;
    int num_sms;
    double *d_tables;  // c0[n0] c1[n1] A0 A1 A2 [n0 each] B0 B1 [n1 each]
    StepC *d_steps;
    long long steps_cap;
    double *d_w;
    long long w_cap;
    double *d_lik;  // internal likelihood table [T][G] shared by the combos of a call
    long long lik_cap;
    int *d_sm_state;  // per-SM arrival counters + per-combo claim flags (fast 1-D kernels with sm_assign)
    long long sm_state_cap;
    double *d_factor;  // device scalar: re-base factor of blg_wave_weights / blg_rebase
    double *d_o2;  // tiled OnlineStudy step: unnormalised cells [B][G] + per-tile partial sums, kept between steps
    long long o2_cap;
    Opts opt;
};

extern "C" {

int blg_version(void) { return BLG_ABI_VERSION; }
const char *blg_last_error(void) { return g_err; }
const char *blg_backend(void) { return "cuda:sm_100a"; }
int64_t blg_launch_count(void) { return g_launches.load(); }
const char *blg_last_kernel(void) { return g_last_kernel; }

int blg_plan_create(const blg_problem *p, blg_plan **out) {
    if (!p || !out) return fail("null argument");
    if (p->ndim < 1 || p->ndim > 2) return fail("ndim must be 1 or 2");
    const int n0 = p->n[0], n1 = p->ndim == 2 ? p->n[1] : 1;
    if (n0 < 1 || n1 < 1) return fail("bad grid size");
    if (!p->coords[0] || (p->ndim == 2 && !p->coords[1])) return fail("coords missing");
    const int om = p->om_kind;
    const bool twoD = (om == BLG_OM_GAUSSIAN || om == BLG_OM_SCALED_AR1 || om == BLG_OM_AR1 || om == BLG_OM_LAPLACE);
    if (om != BLG_OM_TABLE && twoD != (p->ndim == 2)) return fail("observation model / grid dimension mismatch");
    if ((om == BLG_OM_AR1 || om == BLG_OM_SCALED_AR1) ? p->seg_len != 2 : (om != BLG_OM_TABLE && p->seg_len != 1))
        return fail("segment length does not match the observation model");
    if (om == BLG_OM_GAUSSIAN_MEAN && p->n_cols != 2) return fail("GAUSSIAN_MEAN needs 2 data columns");
    if (p->n_cols < 1) return fail("n_cols must be >= 1");

    blg_plan *pl = new blg_plan();
    pl->pb = *p;
    pl->pb.n[1] = n1;
    CUDA_TRY(cudaGetDevice(&pl->device));
    CUDA_TRY(cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, pl->device));
    {   // scratch of the stream kernels comes from the stream-ordered pool: keep it cached between calls
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, pl->device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    opts_from_env(pl->opt);

    // host tables (see lik_column in common.cuh)
    std::vector<double> h((size_t)(4 * n0 + 3 * n1), 0.0);
    double *c0 = h.data(), *c1 = c0 + n0, *A0 = c1 + n1, *A1 = A0 + n0, *A2 = A1 + n0, *B0 = A2 + n0, *B1 = B0 + n1;
    for (int i = 0; i < n0; ++i) c0[i] = p->coords[0][i];
    for (int j = 0; j < n1; ++j) c1[j] = p->ndim == 2 ? p->coords[1][j] : 0.0;
    bool useA[3] = {false, false, false}, useB[2] = {false, false};
    for (int i = 0; i < n0; ++i) {
        const double x = c0[i];
        switch (om) {
            case BLG_OM_POISSON:
                A0[i] = x;
                // lambda == 0: the reference evaluates 0**k * exp(-0) / k! = (k == 0) (observationModels.py:502); with
                // log(0) = -inf the device form k * log(lambda) would be 0 * -inf = NaN for a zero count
                A1[i] = x == 0.0 ? -DBL_MAX : log(x);
                useA[0] = useA[1] = true;
                break;
            case BLG_OM_GAUSSIAN:
            case BLG_OM_AR1:
            case BLG_OM_GAUSSIAN_MEAN:
            case BLG_OM_LAPLACE:
                A0[i] = x;
                useA[0] = true;
                break;
            case BLG_OM_SCALED_AR1:
                A0[i] = x;
                A1[i] = 1.0 / (1.0 - x * x);
                A2[i] = -0.5 * log(1.0 - x * x);
                useA[0] = useA[1] = useA[2] = true;
                break;
            case BLG_OM_WHITE_NOISE:
                A0[i] = 1.0 / (2.0 * x * x);
                A1[i] = -0.5 * log(2.0 * M_PI * x * x);
                useA[0] = useA[1] = true;
                break;
            case BLG_OM_BERNOULLI:
                A0[i] = (x > 1.0 || x < 0.0) ? 0.0 : x;
                useA[0] = true;
                break;
            default:
                break;
        }
    }
    for (int j = 0; j < n1; ++j) {
        const double y = c1[j];
        switch (om) {
            case BLG_OM_GAUSSIAN:
            case BLG_OM_AR1:
            case BLG_OM_SCALED_AR1:
                B0[j] = 1.0 / (2.0 * y * y);
                B1[j] = -0.5 * log(2.0 * M_PI * y * y);
                useB[0] = useB[1] = true;
                break;
            case BLG_OM_LAPLACE:
                B0[j] = 1.0 / y;
                B1[j] = -log(2.0 * y);
                useB[0] = useB[1] = true;
                break;
            default:
                break;
        }
    }
    if (cudaMalloc(&pl->d_tables, h.size() * sizeof(double)) != cudaSuccess) {
        delete pl;
        return fail("cudaMalloc of plan tables failed");
    }
    if (cudaMemcpy(pl->d_tables, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(pl->d_tables);
        delete pl;
        return fail("upload of plan tables failed");
    }
    DevProblem &d = pl->dev;
    d.ndim = p->ndim;
    d.n0 = n0;
    d.n1 = n1;
    d.G = n0 * n1;
    d.om_kind = om;
    d.seg = p->seg_len;
    d.ncols = p->n_cols;
    d.ncols_eff = om == BLG_OM_GAUSSIAN_MEAN ? 1 : p->n_cols;
    d.lc_prod = p->lattice[0] * (p->ndim == 2 ? p->lattice[1] : 1.0);
    double *base = pl->d_tables;
    d.c0 = base;
    d.c1 = base + n0;
    double *dA0 = base + n0 + n1;
    d.tabA[0] = useA[0] ? dA0 : nullptr;
    d.tabA[1] = useA[1] ? dA0 + n0 : nullptr;
    d.tabA[2] = useA[2] ? dA0 + 2 * n0 : nullptr;
    d.tabB[0] = useB[0] ? dA0 + 3 * n0 : nullptr;
    d.tabB[1] = useB[1] ? dA0 + 3 * n0 + n1 : nullptr;
    pl->d_steps = nullptr;
    pl->steps_cap = 0;
    pl->d_w = nullptr;
    pl->w_cap = 0;
    pl->d_lik = nullptr;
    pl->lik_cap = 0;
    pl->d_sm_state = nullptr;
    pl->sm_state_cap = 0;
    pl->d_o2 = nullptr;
    pl->o2_cap = 0;
    pl->d_factor = nullptr;
    if (cudaMalloc(&pl->d_factor, sizeof(double)) != cudaSuccess) {
        cudaFree(pl->d_tables);
        delete pl;
        return fail("cudaMalloc of plan scalars failed");
    }
    *out = pl;
    return 0;
}

int blg_plan_set_option(blg_plan *pl, const char *name, int64_t value) {
    if (!pl || !name) return fail("null argument");
    for (const OptName &n : kOptNames)
        if (strcmp(n.name, name) == 0) {
            pl->opt.*(n.field) = (int)value;
            return 0;
        }
    return fail("unknown plan option: %s", name);
}

void blg_plan_destroy(blg_plan *pl) {
    if (!pl) return;
    cudaFree(pl->d_tables);
    if (pl->d_steps) cudaFree(pl->d_steps);
    if (pl->d_w) cudaFree(pl->d_w);
    if (pl->d_lik) cudaFree(pl->d_lik);
    if (pl->d_sm_state) cudaFree(pl->d_sm_state);
    if (pl->d_o2) cudaFree(pl->d_o2);
    if (pl->d_factor) cudaFree(pl->d_factor);
    delete pl;
}

}  // extern "C"

namespace {

int ensure_steps(blg_plan *pl, long long count) {
    if (count <= pl->steps_cap) return 0;
    if (pl->d_steps) CUDA_TRY(cudaFree(pl->d_steps));
    pl->d_steps = nullptr;
    pl->steps_cap = 0;
    CUDA_TRY(cudaMalloc(&pl->d_steps, (size_t)count * sizeof(StepC)));
    pl->steps_cap = count;
    return 0;
}

int ensure_w(blg_plan *pl, long long count) {
    if (count <= pl->w_cap) return 0;
    if (pl->d_w) CUDA_TRY(cudaFree(pl->d_w));
    pl->d_w = nullptr;
    pl->w_cap = 0;
    CUDA_TRY(cudaMalloc(&pl->d_w, (size_t)count * sizeof(double)));
    pl->w_cap = count;
    return 0;
}

// Shared likelihood table: worth it as soon as a few combos share it; bounded so it never competes with alpha_seq.
// permM > 0: "owner order" of the warp-specialised 1-D kernels (rows of permM planes x permNC entries, see
// lik_table_perm_kernel); a caller-supplied table (BLG_OM_TABLE) is permuted into the same scratch.
int prep_lik_table(blg_plan *pl, const blg_inputs *in, PassArgs &a, cudaStream_t st, int permM = 0, int permML = 0,
                   int permNC = 0, bool force = false) {
    const DevProblem &d = pl->dev;
    a.lik_pitch = d.G;
    if (!permM && (d.om_kind == BLG_OM_TABLE || (pl->opt.no_lik_table && !force))) return 0;
    const long long pitch = permM ? (long long)permM * permNC : (long long)d.G;
    const long long count = in->T * pitch;
    if (!permM && !force && in->B < 4) return 0;
    if (count * 8 > (6LL << 30)) return (permM || force) ? 1 : 0;
    if (count > pl->lik_cap) {
        if (pl->d_lik) CUDA_TRY(cudaFree(pl->d_lik));
        pl->d_lik = nullptr;
        pl->lik_cap = 0;
        if (cudaMalloc(&pl->d_lik, (size_t)count * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return (permM || force) ? 1 : 0;  // no room: keep evaluating the likelihood in the passes
        }
        pl->lik_cap = count;
    }
    const int nt = 256;
    if (permM)
        lik_table_perm_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, st>>>(
            d, pl->d_steps, d.om_kind == BLG_OM_TABLE ? in->lik_table : nullptr, in->T, permM, permML, permNC, pl->d_lik);
    else
        lik_table_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, st>>>(d, pl->d_steps, in->T, pl->d_lik);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    a.pb.om_kind = BLG_OM_TABLE;
    a.lik_table = pl->d_lik;
    a.lik_pitch = pitch;
    return 0;
}

int prep_steps(blg_plan *pl, const blg_inputs *in, cudaStream_t st) {
    const DevProblem &d = pl->dev;
    if (d.om_kind == BLG_OM_TABLE) return ensure_steps(pl, 1);
    if (!in->data) return fail("data pointer missing");
    const long long count = in->T * d.ncols_eff;
    if (ensure_steps(pl, count > 0 ? count : 1)) return -1;
    if (count > 0) {
        const int nt = 256;
        prep_steps_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, st>>>(in->data, pl->d_steps, in->T, d.om_kind,
                                                                             d.ncols, d.ncols_eff);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

struct Layout {
    size_t bytes;
    int nt;
};

// Shared-memory layout of the resident kernels; returns false if the grid does not fit.
bool resident_layout(const blg_plan *pl, const blg_program &pg, bool backward, bool want_stage, PassArgs &a, Layout &lay) {
    const DevProblem &d = pl->dev;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = even_up(d.n1);
    int off = 2 * a.Gp;
    a.off_stage = -1;
    if (backward && want_stage) {
        a.off_stage = off;
        off += 2 * a.Gp;
    }
    a.off_tab = off;
    off += 3 * a.n0p + 2 * a.n1p;
    a.off_w = off;
    int woff = 0;
    for (int k = 0; k < pg.n_ops; ++k) {
        a.pg.w_off[k] = woff;
        a.pg.w_len[k] = 0;
        if (pg.kind[k] == BLG_OP_GRW) {
            const int taps = 2 * pg.max_radius[k] + 1;
            const int len = ((taps + kConvM - 1) / kConvM) * kConvM + kConvM;
            a.pg.w_len[k] = even_up(len);
            woff += a.pg.w_len[k];
        }
    }
    off += woff;
    a.off_misc = even_up(off);
    lay.bytes = (size_t)(a.off_misc + kMiscDoubles) * sizeof(double);
    int nt = ((d.G + 3) / 4 + 31) / 32 * 32;
    if (nt < 128) nt = 128;
    if (nt > 256 && nt <= 512) nt = 512;
    if (nt > 512) nt = 1024;
    lay.nt = nt;
    return lay.bytes <= kSmemLimit;
}

// Fast path (fast1d.cuh): 1-D grid, program = one GaussianRandomWalk, halo <= n, one work item per thread.
bool fast1d_layout(const blg_plan *pl, const blg_program &pg, bool backward, int M, PassArgs &a, Layout &lay) {
    const DevProblem &d = a.pb;  // om_kind is TABLE when the shared likelihood table is in use
    if (pl->opt.no_fast1d || pl->opt.force_stream) return false;
    if (d.ndim != 1 || pg.n_ops != 1 || pg.kind[0] != BLG_OP_GRW) return false;
    const int halo = even_up(pg.max_radius[0] + 2 * M);
    const int items = (d.G + M - 1) / M;
    if (halo > d.G || items > 1024) return false;
    a.halo = halo;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = 2;
    const int pitch = a.Gp + 2 * halo;
    int off = 2 * pitch;
    a.off_stage = -1;
    if (backward) {
        a.off_stage = off;
        off += 2 * a.Gp;
    }
    a.off_tab = -1;
    if (d.om_kind != BLG_OM_TABLE) {
        a.off_tab = off;
        off += 3 * a.n0p;
    }
    a.off_w = off;
    const int taps = 2 * pg.max_radius[0] + 1;
    a.pg.w_off[0] = 0;
    a.pg.w_len[0] = ((taps + M - 1) / M + 1) * (M + 1);  // chunk-padded layout of conv_item (fast1d.cuh)
    off += a.pg.w_len[0];
    a.off_misc = even_up(off);
    lay.bytes = (size_t)(a.off_misc + kMiscDoubles) * sizeof(double);
    int nt = (items + 31) / 32 * 32;
    lay.nt = nt;
    return lay.bytes <= kSmemLimit;
}

// Warp-specialised fast path (fast1d_ws.cuh): same shapes as fast1d_layout; chooses (M, threads) so that the compute
// warps (threads/32 - 1) cover the grid with M cells per thread.
bool fast1d_ws_layout(const blg_plan *pl, const blg_program &pg, bool backward, PassArgs &a, Layout &lay, int &M, int &ML) {
    const DevProblem &d = pl->dev;
    if (pl->opt.no_fast1d || pl->opt.no_ws || pl->opt.force_stream) return false;
    if (d.ndim != 1 || pg.n_ops != 1 || pg.kind[0] != BLG_OP_GRW) return false;
    // 4 compute warps (one per SM sub-partition) + the service warp = 160 threads; beyond 128 * 11 cells 8 compute warps
    // (two per sub-partition) = 288 threads.  Round 2, per-warp trace profiles/r2d_ws_trace.txt: with 3 compute warps
    // per chain the busiest sub-partition carried 1.6x the mean convolution load (the chain's pace is set by it); with 4
    // the load is level (0.93).  Among the compiled geometries {M cells per thread, ML in the last compute warp} the
    // one with the fewest cell slots that cover the grid wins (every slot costs its DFMAs, owned or not): G = 1000 ->
    // 3 x 32 x 9 + 32 x 5 = 1024 slots instead of 4 x 32 x 9 = 1152.
    int nt = 0;
    M = ML = 0;
    for (int pass = 0; pass < 2 && !M; ++pass) {
        const int want = pass ? 288 : 160;
        long long best = 0;
        for (const int *g = fast1d_ws_geometries(); g[0]; g += 3) {
            if (g[2] != want || (pl->opt.ws_even && g[0] != g[1])) continue;
            const long long slots = (long long)(want - 64) * g[0] + 32LL * g[1];
            if (slots >= d.G && (!M || slots < best)) {
                M = g[0];
                ML = g[1];
                nt = want;
                best = slots;
            }
        }
    }
    if (!M) return false;
    if (pl->opt.ws_m > 0 && pl->opt.ws_nt > 0) {  // tuning override: (cells per thread [, in the last warp], threads)
        const int ml = pl->opt.ws_ml > 0 ? pl->opt.ws_ml : pl->opt.ws_m;
        if ((long long)(pl->opt.ws_nt - 64) * pl->opt.ws_m + 32LL * ml >= d.G &&
            fwd_fast1d_ws_entry(pl->opt.ws_m, ml, pl->opt.ws_nt)) {
            M = pl->opt.ws_m;
            ML = ml;
            nt = pl->opt.ws_nt;
        }
    }
    const int halo = even_up(pg.max_radius[0] + 2 * M);
    if (halo > d.G) return false;
    a.halo = halo;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = 2;
    const int pitch = a.Gp + 2 * halo;
    int off = 2 * pitch;
    a.off_stage = -1;
    if (backward) {
        a.off_stage = off;
        off += 2 * a.Gp;
    }
    a.off_tab = -1;
    a.off_w = off;
    const int taps = 2 * pg.max_radius[0] + 1;
    a.pg.w_off[0] = 0;
    a.pg.w_len[0] = ((taps + M - 1) / M + 1) * (M + 1);  // chunk-padded layout of conv_item (fast1d.cuh)
    off += a.pg.w_len[0];
    a.ws_w2 = off;
    a.ws_w2_len = ML != M ? ((taps + ML - 1) / ML + 1) * (ML + 1) : 0;  // the same weights in chunks of ML
    off += a.ws_w2_len;
    a.off_misc = even_up(off);
    off = a.off_misc + kMiscDoubles;
    a.ws_part = off;
    off += 6 * (nt - 32);  // partial sums: [2 parities][3 sums][compute threads]
    a.ws_ctl = off;
    off += 4;
    lay.bytes = (size_t)off * sizeof(double);
    lay.nt = nt;
    return lay.bytes <= kSmemLimit;
}

// DMMA variant of the warp-specialised 1-D kernels (fast1d_mma.cuh): tiles of 64 cells, swizzled state buffers.
bool fast1d_mma_layout(const blg_plan *pl, const blg_program &pg, bool backward, PassArgs &a, Layout &lay, int &tpw) {
    const DevProblem &d = pl->dev;
    if (pl->opt.no_fast1d || pl->opt.no_ws || pl->opt.no_mma || pl->opt.force_stream) return false;
    if (d.ndim != 1 || pg.n_ops != 1 || pg.kind[0] != BLG_OP_GRW || d.G % 2) return false;
    const int ntiles = (d.G + 63) / 64;
    int nt = 160;
    tpw = (ntiles + 3) / 4;
    if (tpw > 6 || (pl->opt.mma_nt == 288 && ntiles > 4)) {  // 8 compute warps (two per sub-partition)
        nt = 288;
        tpw = (ntiles + 7) / 8;
        if (tpw == 3) tpw = 4;  // compiled: 1, 2 (56 registers, 4 chains per SM), 4, 5, 6 (2 chains per SM)
    }
    if (tpw > 6 || !fwd_fast1d_mma_entry(tpw, nt)) return false;
    const int halo = (pg.max_radius[0] + 7 + 7) & ~7;  // mma_halo(): multiple of 8
    if (halo > d.G) return false;
    a.halo = halo;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = 2;
    a.mma_pitch = halo + 64 * ntiles + halo + 8;
    int off = 2 * a.mma_pitch;
    a.off_stage = -1;
    if (backward) {
        a.off_stage = off;
        off += 2 * a.Gp;
    }
    a.off_tab = -1;
    a.off_w = off;
    a.pg.w_off[0] = 0;
    a.pg.w_len[0] = even_up(2 * pg.max_radius[0] + 1 + 2 * 16);  // kMmaWPad zeros on both sides
    off += a.pg.w_len[0];
    a.ws_w2 = off;
    a.ws_w2_len = 0;
    a.off_misc = even_up(off);
    off = a.off_misc + kMiscDoubles;
    a.ws_part = off;
    off += 6 * (nt - 32);  // partial sums: [2 parities][3 sums][compute threads]
    a.ws_ctl = off;
    off += 4;
    lay.bytes = (size_t)off * sizeof(double);
    lay.nt = nt;
    return lay.bytes <= kSmemLimit;
}

// Interleaved variant (fast1d_il.cuh): one CTA per SM works through the SM's list of up to 4 chains.
bool fast1d_il_layout(const blg_plan *pl, const blg_inputs *in, bool backward, PassArgs &a, Layout &lay, int &tpw) {
    const DevProblem &d = pl->dev;
    const blg_program &pg = in->prog;
    if (pl->opt.no_fast1d || pl->opt.no_ws || pl->opt.no_mma || !pl->opt.il || pl->opt.force_stream) return false;
    if (d.ndim != 1 || pg.n_ops != 1 || pg.kind[0] != BLG_OP_GRW || d.G % 2) return false;
    if (!pg.sm_assign || pg.sm_count != pl->num_sms || pg.sm_slots < 1 || pg.sm_slots > 4 || pl->opt.no_sm_assign) return false;
    if ((long long)pg.sm_count * pg.sm_slots < in->B) return false;
    const int ntiles = (d.G + 63) / 64;
    tpw = (ntiles + 15) / 16;
    if (tpw > 2 || !(backward ? bwd_fast1d_il_entry(tpw) : fwd_fast1d_il_entry(tpw))) return false;
    const int halo = (pg.max_radius[0] + 7 + 7) & ~7;  // mma_halo()
    if (halo > d.G) return false;
    a.halo = halo;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = 2;
    a.mma_pitch = halo + 64 * ntiles + halo + 8;
    int off = 2 * a.mma_pitch;  // offsets inside a chain slot
    a.off_stage = -1;
    if (backward) {
        a.off_stage = off;
        off += 2 * a.Gp;
    }
    a.off_tab = -1;
    a.off_w = off;
    a.pg.w_off[0] = 0;
    a.pg.w_len[0] = even_up(2 * pg.max_radius[0] + 1 + 2 * 16);
    off += a.pg.w_len[0];
    a.il_stride = even_up(off);
    off = 4 * a.il_stride;
    a.off_misc = off;
    off += kMiscDoubles;
    a.ws_part = off;
    off += 4 * 2 * 3 * 16 * 16;  // kIlChains * kIlPP
    a.il_ctl = off;
    off += 4 * 8;           // kIlChains * kIlCtlDoubles
    a.ws_ctl = off;         // mbarriers: alpha rings (backward) [4 chains][2 slots], then one per chain (step done)
    off += 12;
    a.ws_w2 = 0;
    a.ws_w2_len = 0;
    lay.bytes = (size_t)off * sizeof(double);
    lay.nt = 17 * 32;
    return lay.bytes <= kSmemLimit;
}

// Stream kernels (grids that do not fit in shared memory): state in global scratch, convolution tiles in smem.
bool stream_layout(const blg_plan *pl, const blg_program &pg, PassArgs &a, Layout &lay, int chunkM = 0) {
    const DevProblem &d = pl->dev;
    a.halo = 0;
    a.Gp = even_up(d.G);
    a.n0p = even_up(d.n0);
    a.n1p = even_up(d.n1);
    a.off_stage = -1;
    int off = 0;
    a.off_tab = off;
    off += 3 * a.n0p + 2 * a.n1p;
    a.off_w = off;
    int woff = 0;
    for (int k = 0; k < pg.n_ops; ++k) {
        a.pg.w_off[k] = woff;
        a.pg.w_len[k] = 0;
        if (pg.kind[k] == BLG_OP_GRW) {
            const int taps = 2 * pg.max_radius[k] + 1;
            a.pg.w_len[k] = chunkM ? ((taps + chunkM - 1) / chunkM + 1) * (chunkM + 1)
                                   : even_up(((taps + kConvM - 1) / kConvM) * kConvM + kConvM);
            woff += a.pg.w_len[k];
        }
    }
    off += woff;
    a.off_misc = even_up(off);
    off = a.off_misc + kMiscDoubles;
    a.off_tile = off;
    const long long room = (long long)(kSmemLimit / sizeof(double)) - off;
    long long tile = room < 24576 ? room : 24576;  // <= 192 KB; the rest of the SM's memory stays L1 for the streamed state
    const int longest = d.n0 > d.n1 ? d.n0 : d.n1;
    int rmax = 0;
    for (int k = 0; k < pg.n_ops; ++k)
        if (pg.kind[k] == BLG_OP_GRW && pg.max_radius[k] > rmax) rmax = pg.max_radius[k];
    // generic stream kernel: one complete line; 2-D kernels: one line with halo + its output line
    if (tile < (chunkM ? 2LL * longest + 2 * (rmax + 2 * chunkM) : (long long)longest)) return false;
    a.tile_doubles = (int)tile;
    lay.bytes = (size_t)(off + tile) * sizeof(double);
    lay.nt = 1024;
    return true;
}

int launch_stream(PassKernel kernel, blg_plan *pl, PassArgs &a, Layout lay, long long B, cudaStream_t st, const char *name,
                  int threads = 1024) {
    lay.nt = threads;
    g_last_kernel = name;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    long long grid = B < pl->num_sms ? B : pl->num_sms;  // one persistent CTA per SM, looping over combos
    double *scratch = nullptr;
    CUDA_TRY(cudaMallocAsync(&scratch, (size_t)grid * 2 * a.Gp * sizeof(double), st));
    a.scratch = scratch;
    a.use_bulk = 0;
    if (pl->opt.verbose)
        fprintf(stderr, "[blgrid] %s: grid %lld x %d threads, %zu B smem/CTA, tile %d doubles, scratch %.1f MB\n", name,
                grid, lay.nt, lay.bytes, a.tile_doubles, grid * 2.0 * a.Gp * 8 / 1e6);
    kernel<<<(unsigned)grid, lay.nt, lay.bytes, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(scratch, st));
    return 0;
}

// Cluster-resident 2-D kernels (cluster2d.cuh): picks the cluster size C (bands of rows per CTA) such that one band
// is one work item per thread in both convolutions, the axis-0 radius fits in the smallest band, and the state
// buffer (+ the alpha staging band of the backward pass) fits in shared memory.
bool cluster2d_layout(const blg_plan *pl, const blg_program &pg, uint32_t flags, bool backward, PassArgs &a, Layout &lay,
                      int &C, int &M0out) {
    const DevProblem &d = pl->dev;
    if (d.ndim != 2 || pl->opt.no_cluster2d) return false;
    if (flags & (BLG_F_INIT_STATE | BLG_F_SAVE_STATE | BLG_F_TRANSITION_FIRST | BLG_F_ACCUMULATE)) return false;
    if (!cluster2d_supports(pg.n_ops, pg.kind, pg.axis)) return false;
    if (d.n1 % 2) return false;  // 16-byte aligned bands for the bulk-async copies
    int NT, M0, M1, cells, wpad;
    cluster2d_params(&NT, &M0, &M1, &cells, &wpad);
    int r0max = 0, r1max = 0;
    for (int k = 0; k < pg.n_ops; ++k)
        if (pg.kind[k] == BLG_OP_GRW) {
            int &r = pg.axis[k] == 0 ? r0max : r1max;
            if (pg.max_radius[k] > r) r = pg.max_radius[k];
        }
    if (r1max + M1 > d.n1) return false;
    const int force = pl->opt.cluster2d_c;
    for (int c = 8; c >= 2; c /= 2) {
        if (force && force != c) continue;
        const int nb = (d.n0 + c - 1) / c;
        const int last = d.n0 - (c - 1) * nb;
        if (last < 1 || r0max > last) continue;
        // rows per axis-0 work item: 16, or 13 where that wastes fewer rows of the band (25-row bands of a 200^2 grid)
        int m0 = M0;
        if ((nb + 12) / 13 * 13 < (nb + M0 - 1) / M0 * M0 && d.n1 * ((nb + 12) / 13) <= NT) m0 = 13;
        if (nb * ((d.n1 + M1 - 1) / M1) > NT || d.n1 * ((nb + m0 - 1) / m0) > NT || nb * d.n1 > cells * NT) continue;
        const int nbp = (nb + m0 - 1) / m0 * m0;
        int off = 0;
        a.c2_nb = nb;
        a.c2_h0 = r0max;
        a.c2_off_x = off;
        a.c2_rows = 2 * r0max + nbp;
        a.c2_x_doubles = a.c2_rows * d.n1;
        off += a.c2_x_doubles;
        a.c2_off_s = off;
        if (backward || a.pb.om_kind == BLG_OM_TABLE) off += nb * d.n1;  // alpha band / staged likelihood band
        a.off_w = off;
        int woff = 0;
        for (int k = 0; k < pg.n_ops; ++k) {
            a.pg.w_off[k] = woff;
            a.pg.w_len[k] = 0;
            if (pg.kind[k] == BLG_OP_GRW) {
                a.pg.w_len[k] = even_up(2 * pg.max_radius[k] + 1 + wpad);  // zero taps: unguarded tap loops
                woff += a.pg.w_len[k];
            }
        }
        off += woff;
        a.off_misc = even_up(off);
        lay.bytes = (size_t)(a.off_misc + kMiscDoubles) * sizeof(double);
        lay.nt = NT;
        if (lay.bytes > kSmemLimit) continue;
        a.halo = 0;
        a.Gp = even_up(d.G);
        C = c;
        M0out = m0;
        return true;
    }
    return false;
}

int launch_cluster(const blg_plan *pl, PassKernel kernel, const PassArgs &a, const Layout &lay, long long B, int C,
                   cudaStream_t st, const char *name) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(B * C), 1, 1);
    cfg.blockDim = dim3((unsigned)lay.nt, 1, 1);
    cfg.dynamicSmemBytes = lay.bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {   // a device (or partition) that cannot co-schedule one such cluster: tell the caller to use the stream kernels
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
            cudaGetLastError();
            return 1;
        }
        if (pl->opt.verbose)
            fprintf(stderr, "[blgrid] %s: %lld clusters x %d CTAs x %d threads, %zu B smem/CTA, band %d rows, halo %d rows, "
                            "%d clusters resident\n", name, B, C, lay.nt, lay.bytes, a.c2_nb, a.c2_h0, clusters);
    }
    g_last_kernel = name;
    long long *trace = nullptr;
    PassArgs a2 = a;
    const long long nblk = B * C;
    if (pl->opt.trace[0]) {  // per-CTA phase cycle counters (forward kernel built with PROF)
        CUDA_TRY(cudaMalloc(&trace, (size_t)nblk * 8 * sizeof(long long)));
        CUDA_TRY(cudaMemset(trace, 0, (size_t)nblk * 8 * sizeof(long long)));
        a2.trace = trace;
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, a2));
    ++g_launches;
    if (trace) {
        std::vector<long long> h((size_t)nblk * 8);
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        static int seq = 0;
        char path[512];
        snprintf(path, sizeof path, "%s.%s.%d.csv", pl->opt.trace, name, seq++);
        if (FILE *f = fopen(path, "w")) {
            fprintf(f, "block,c0,c1,c2,c3,c4,c5,c6,c7\n");  // phase counters: see PROF in cluster2d.cuh
            for (long long i = 0; i < nblk; ++i) {
                fprintf(f, "%lld", i);
                for (int k = 0; k < 8; ++k) fprintf(f, ",%lld", h[8 * i + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return 0;
}

// outputs per thread of the fast 1-D kernels: 9 in the forward pass (4 full warps per 1000-cell combo, one per SM
// sub-partition), 7 in the backward pass (more live registers per cell); measured on B200, see profiles/
// Per-SM assignment (blg_program.sm_assign): zeroed counters / claim flags and a grid of sm_count*sm_slots CTAs.
int prep_sm_assign(blg_plan *pl, const blg_inputs *in, PassArgs &a, long long &grid, cudaStream_t st) {
    const blg_program &pg = in->prog;
    a.sm_assign = nullptr;
    a.sm_state = nullptr;
    a.sm_count = 0;
    a.sm_slots = 0;
    grid = in->B;
    if (!pg.sm_assign || pg.sm_count != pl->num_sms || pg.sm_slots < 1 || pl->opt.no_sm_assign) return 0;
    if ((long long)pg.sm_count * pg.sm_slots < in->B) return 0;
    const long long need = pg.sm_count + 2 * in->B;  // arrival counters, claim flags, progress (ws_pace)
    if (need > pl->sm_state_cap) {
        if (pl->d_sm_state) CUDA_TRY(cudaFree(pl->d_sm_state));
        pl->d_sm_state = nullptr;
        pl->sm_state_cap = 0;
        CUDA_TRY(cudaMalloc(&pl->d_sm_state, (size_t)need * sizeof(int)));
        pl->sm_state_cap = need;
    }
    CUDA_TRY(cudaMemsetAsync(pl->d_sm_state, 0, (size_t)(pg.sm_count + in->B) * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(pl->d_sm_state + pg.sm_count + in->B, 0xff, (size_t)in->B * sizeof(int), st));  // -1: not started
    a.ws_progress = pl->d_sm_state + pg.sm_count + in->B;
    {   // pacing of the chains of an SM (fast1d_ws.cuh): publish / check every 2^k steps, hold back beyond `skew` steps
        int every = pl->opt.ws_pace_every;
        while (every & (every - 1)) every &= every - 1;  // power of two (0 = off)
        a.pace_every = pl->opt.ws_pace_skew > 0 ? every : 0;
        a.pace_skew = pl->opt.ws_pace_skew;
    }
    a.sm_assign = pg.sm_assign;
    a.sm_state = pl->d_sm_state;
    a.sm_count = pg.sm_count;
    a.sm_slots = pg.sm_slots;
    grid = (long long)pg.sm_count * pg.sm_slots;
    return 0;
}

constexpr int kFastM = 9;  // cells per thread of the single-role fused 1-D kernels (the one compiled fallback)

int launch_resident(const blg_plan *pl, PassKernel kernel, const PassArgs &a, const Layout &lay, long long B, cudaStream_t st,
                    const char *name) {
    g_last_kernel = name;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    // several CTAs (combos) must share an SM: ask for the full shared-memory carveout
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (pl->opt.verbose) {
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, lay.nt, lay.bytes);
        fprintf(stderr, "[blgrid] %s: grid %lld x %d threads, %zu B smem/CTA, %d CTA/SM, halo %d, bulk %d\n", name, B,
                lay.nt, lay.bytes, occ, a.halo, a.use_bulk);
    }
    long long *trace = nullptr;
    PassArgs a2 = a;
    if (pl->opt.trace[0]) {  // debugging aid: per-CTA {smid, combo, start, end} (globaltimer ns), dumped as CSV
        // + per warp {convolution, epilogue, barrier cycles, hardware warp id} (warp-specialised forward kernel)
        // + per warp and step (32 steps in the middle of the series) the SM clock at {convolution start, convolution
        //   issued, epilogue done, barrier passed} (DMMA forward kernel): [B][4 warps][32 steps][4]
        CUDA_TRY(cudaMalloc(&trace, (size_t)B * (36 + 512) * sizeof(long long)));
        CUDA_TRY(cudaMemset(trace, 0, (size_t)B * (36 + 512) * sizeof(long long)));
        a2.trace = trace;
    }
    kernel<<<(unsigned)B, lay.nt, lay.bytes, st>>>(a2);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    if (trace) {
        std::vector<long long> h((size_t)B * (36 + 512));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        static int seq = 0;
        char path[512];
        snprintf(path, sizeof path, "%s.%s.%d.csv", pl->opt.trace, name, seq++);
        if (FILE *f = fopen(path, "w")) {
            fprintf(f, "block,smid,combo,start_ns,end_ns");
            for (int w = 0; w < 8; ++w) fprintf(f, ",w%d_conv,w%d_epi,w%d_bar,w%d_hwid", w, w, w, w);
            fprintf(f, "\n");
            for (long long i = 0; i < B; ++i) {
                fprintf(f, "%lld,%lld,%lld,%lld,%lld", i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                for (int k = 0; k < 32; ++k) fprintf(f, ",%lld", h[4 * B + 32 * i + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        bool any = false;
        for (size_t k = (size_t)36 * B; k < h.size() && !any; ++k) any = h[k] != 0;
        snprintf(path, sizeof path, "%s.%s.%d.events.csv", pl->opt.trace, name, seq - 1);
        if (FILE *f = any ? fopen(path, "w") : nullptr) {
            fprintf(f, "block,warp,step,conv_start,conv_issued,epi_done,bar_passed\n");
            for (long long i = 0; i < B; ++i)
                for (int w = 0; w < 4; ++w)
                    for (int q = 0; q < 32; ++q) {
                        const long long *e = &h[(size_t)36 * B + ((size_t)(i * 4 + w) * 32 + q) * 4];
                        if (e[0]) fprintf(f, "%lld,%d,%d,%lld,%lld,%lld,%lld\n", i, w, q, e[0], e[1], e[2], e[3]);
                    }
            fclose(f);
        }
    }
    return 0;
}

int fill_args(blg_plan *pl, const blg_inputs *in, const blg_outputs *out, uint32_t flags, PassArgs &a) {
    const blg_program &pg = in->prog;
    if (pg.n_ops < 0 || pg.n_ops > BLG_MAX_OPS) return fail("n_ops out of range");
    if (pg.n_ops > 0 && (!pg.param || !pg.radius || !pg.window)) return fail("program arrays missing");
    a.pb = pl->dev;
    a.pg.n_ops = pg.n_ops;
    bool hasReset = false;
    for (int k = 0; k < pg.n_ops; ++k) {
        a.pg.kind[k] = pg.kind[k];
        a.pg.axis[k] = pg.axis[k];
        if (pg.kind[k] == BLG_OP_GRW && (pg.axis[k] < 0 || pg.axis[k] >= pl->dev.ndim)) return fail("GRW axis out of range");
        if (pg.kind[k] == BLG_OP_RESET) hasReset = true;
        if (pg.kind[k] < BLG_OP_GRW || pg.kind[k] > BLG_OP_NOTEQUAL) return fail("unknown operator kind");
    }
    if (hasReset && !in->reset_base) return fail("reset_base required by a RESET operator");
    a.pg.param = pg.param;
    a.pg.radius = pg.radius;
    a.pg.window = pg.window;
    a.order = pl->opt.no_order ? nullptr : pg.order;
    a.T = in->T;
    a.B = in->B;
    a.prior = in->prior;
    a.reset_base = in->reset_base;
    a.lik_table = in->lik_table;
    a.log_weight = in->log_weight;
    a.init_state = in->init_state;
    a.logE = out->log_evidence;
    a.local = out->local_evidence;
    a.alive = out->alive;
    a.alpha_seq = out->alpha_seq;
    a.avg = out->avg;
    a.final_state = out->final_state;
    a.row_scale = (flags & BLG_F_RAW_POSTERIOR) ? out->row_scale : nullptr;
    a.seq_stride = out->seq_stride > 0 ? out->seq_stride : in->T * (long long)pl->dev.G;
    a.row_stride = out->row_stride > 0 ? out->row_stride : in->T;
    a.alpha_src = in->alpha_src;
    a.src_stride = in->src_stride > 0 ? in->src_stride : in->T * (long long)pl->dev.G;
    if (a.seq_stride < in->T * (long long)pl->dev.G || a.row_stride < in->T) return fail("seq_stride / row_stride smaller than one sequence");
    a.steps = pl->d_steps;
    a.flags = flags;
    a.num_sms = pl->num_sms;
    a.serpentine = (pl->opt.serpentine && in->B > pl->num_sms) ? 1 : 0;
    if (pl->dev.om_kind == BLG_OM_TABLE && !in->lik_table) return fail("lik_table required for BLG_OM_TABLE");
    return 0;
}

}  // namespace

extern "C" {

int blg_forward(blg_plan *pl, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream) {
    if (!pl || !in || !out) return fail("null argument");
    if (!out->log_evidence) return fail("log_evidence output missing");
    if (in->B <= 0 || in->T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool store = !(flags & BLG_F_EVIDENCE_ONLY);
    if (store && !out->alpha_seq) return fail("alpha_seq required unless EVIDENCE_ONLY");
    if ((flags & BLG_F_INIT_STATE) ? !in->init_state : !in->prior) return fail("initial state missing");
    if ((flags & BLG_F_SAVE_STATE) && !out->final_state) return fail("final_state missing");
    PassArgs a;
    memset(&a, 0, sizeof a);
    if (prep_steps(pl, in, st)) return -1;
    if (fill_args(pl, in, out, flags, a)) return -1;
    Layout lay;
    const bool bulkOk = store && (pl->dev.G % 2 == 0) && ((uintptr_t)out->alpha_seq % 16 == 0) && a.seq_stride % 2 == 0 &&
                        !pl->opt.no_bulk;
    {   // DMMA kernels: rows leave by 16-byte stores from registers, the likelihood arrives by 16-byte loads
        int tpw = 0;
        const bool rowsOk = !store || bulkOk;
        const bool tabOk = pl->dev.om_kind != BLG_OM_TABLE || (uintptr_t)in->lik_table % 16 == 0;
        if (rowsOk && tabOk && fast1d_il_layout(pl, in, false, a, lay, tpw)) {  // the chains of an SM in one CTA
            const int rc = prep_lik_table(pl, in, a, st, 0, 0, 0, true);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = bulkOk ? 1 : 0;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                return launch_resident(pl, fwd_fast1d_il_entry(tpw), a, lay, a.sm_count, st, "fwd_fast1d_il");
            }
            a.pb.om_kind = pl->dev.om_kind;
            a.lik_table = in->lik_table;
        }
        if (rowsOk && tabOk && fast1d_mma_layout(pl, in->prog, false, a, lay, tpw)) {
            const int rc = prep_lik_table(pl, in, a, st, 0, 0, 0, true);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = bulkOk ? 1 : 0;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                if (PassKernel k = fwd_fast1d_mma_entry(tpw, lay.nt, pl->opt.trace[0] != 0)) return launch_resident(pl, k, a, lay, grid, st, "fwd_fast1d_mma");
                return fail("DMMA forward kernel missing");
            }
        }
        a.pb.om_kind = pl->dev.om_kind;
        a.lik_table = in->lik_table;
    }
    {
        int wsM = 0, wsML = 0;
        if (fast1d_ws_layout(pl, in->prog, false, a, lay, wsM, wsML)) {
            const int rc = prep_lik_table(pl, in, a, st, wsM, wsML, lay.nt - 32);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = bulkOk ? 1 : 0;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                if (PassKernel k = fwd_fast1d_ws_entry(wsM, wsML, lay.nt)) return launch_resident(pl, k, a, lay, grid, st, "fwd_fast1d_ws");
                return fail("warp-specialised forward kernel missing");
            }
        }
    }
    if (prep_lik_table(pl, in, a, st)) return -1;
    const int M = kFastM;
    if (fast1d_layout(pl, in->prog, false, M, a, lay)) {
        a.use_bulk = bulkOk ? 1 : 0;
        long long grid = in->B;
        if (prep_sm_assign(pl, in, a, grid, st)) return -1;
        if (PassKernel k = fwd_fast1d_entry(M, lay.nt)) return launch_resident(pl, k, a, lay, grid, st, "fwd_fast1d");
    }
    a.halo = 0;
    {   // one OnlineStudy step on a large 2-D grid: tiles x hypotheses over the whole GPU (online2d.cuh)
        const uint32_t need = BLG_F_SEPARABLE_ROWS | BLG_F_EVIDENCE_ONLY | BLG_F_INIT_STATE | BLG_F_TRANSITION_FIRST |
                              BLG_F_SAVE_STATE;
        if (in->T == 1 && pl->dev.ndim == 2 && (flags & need) == need && pl->opt.online2d && in->B <= 65535 &&
            (pl->opt.online2d_small || !resident_layout(pl, in->prog, false, false, a, lay))) {
            int r0 = 0, r1 = 0;
            bool ok = true;
            for (int k = 0; k < in->prog.n_ops; ++k) {
                if (in->prog.kind[k] == BLG_OP_NOTEQUAL) ok = false;
                if (in->prog.kind[k] == BLG_OP_GRW) {
                    int &r = in->prog.axis[k] == 0 ? r0 : r1;
                    if (in->prog.max_radius[k] > r) r = in->prog.max_radius[k];
                }
            }
            O2Launch L;
            if (ok && online2d_plan(pl->dev.n0, pl->dev.n1, r0, r1, pl->opt.online2d_async != 0, &L)) {
                const size_t doubles = online2d_scratch_doubles(in->B, pl->dev.G, L);
                if ((long long)doubles > pl->o2_cap) {  // grows once per study: no allocation on the per-step path
                    if (pl->d_o2) CUDA_TRY(cudaFree(pl->d_o2));
                    pl->d_o2 = nullptr;
                    pl->o2_cap = 0;
                    CUDA_TRY(cudaMalloc(&pl->d_o2, doubles * sizeof(double)));
                    pl->o2_cap = (long long)doubles;
                }
                if (pl->opt.verbose)
                    fprintf(stderr, "[blgrid] online2d: %lld hypotheses x %d x %d tiles, %zu B smem/CTA, radii <= %d / %d\n",
                            (long long)in->B, L.tilesY, L.tilesX, L.smemBytes, r0, r1);
                const int rc = online2d_run(a, L, pl->d_o2, st);
                g_launches += 4;
                g_last_kernel = "online2d";
                if (rc != 0) return fail("online2d launch failed: %s", cudaGetErrorString((cudaError_t)rc));
                return 0;
            }
        }
    }
    {
        int C = 0, m0 = 16;
        const bool want = pl->opt.cluster2d || !resident_layout(pl, in->prog, false, false, a, lay);
        if (want && !pl->opt.force_stream && (!store || (uintptr_t)out->alpha_seq % 16 == 0) &&
            cluster2d_layout(pl, in->prog, flags, false, a, lay, C, m0)) {
            const int rc = launch_cluster(pl, fwd_cluster2d_entry(pl->opt.trace[0] != 0, m0), a, lay, in->B, C, st, "fwd_cluster2d");
            if (rc <= 0) return rc;  // 1: clusters cannot be scheduled here -> stream kernels below
        }
    }
    if (!resident_layout(pl, in->prog, false, false, a, lay) || pl->opt.force_stream) {
        if (!stream_layout(pl, in->prog, a, lay)) return fail("grid / kernel radius too large for the stream forward kernel");
        return launch_stream(fwd_resident_entry(1024, true), pl, a, lay, in->B, st, "fwd_stream");
    }
    a.use_bulk = bulkOk ? 1 : 0;
    return launch_resident(pl, fwd_resident_entry(lay.nt, false), a, lay, in->B, st, "fwd_resident");
}

int blg_backward(blg_plan *pl, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream) {
    if (!pl || !in || !out) return fail("null argument");
    if (!out->alpha_seq || !out->log_evidence) return fail("alpha_seq / log_evidence missing");
    if (in->B <= 0 || in->T <= 0) return 0;
    const bool acc = (flags & BLG_F_ACCUMULATE) != 0;
    if (acc && (!out->avg || !in->log_weight)) return fail("avg and log_weight required with ACCUMULATE");
    cudaStream_t st = (cudaStream_t)stream;
    PassArgs a;
    memset(&a, 0, sizeof a);
    if (prep_steps(pl, in, st)) return -1;
    if (fill_args(pl, in, out, flags, a)) return -1;
    if (flags & BLG_F_RAW_POSTERIOR) {  // kernels that normalise their rows leave the factor at 1
        if (!out->row_scale) return fail("row_scale required with RAW_POSTERIOR");
        if (acc) return fail("RAW_POSTERIOR and ACCUMULATE exclude each other");
        const long long n = in->B * in->T;
        fill_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out->row_scale, in->B, in->T, a.row_stride, 1.0);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    Layout lay;
    const bool alignedRows = (pl->dev.G % 2 == 0) && ((uintptr_t)out->alpha_seq % 16 == 0) && a.seq_stride % 2 == 0 &&
                             (!in->alpha_src || ((uintptr_t)in->alpha_src % 16 == 0 && a.src_stride % 2 == 0)) && !pl->opt.no_bulk;
    {
        int tpw = 0;
        const bool tabOk = pl->dev.om_kind != BLG_OM_TABLE || (uintptr_t)in->lik_table % 16 == 0;
        if (alignedRows && !acc && tabOk && fast1d_il_layout(pl, in, true, a, lay, tpw)) {  // the chains of an SM in one CTA
            const int rc = prep_lik_table(pl, in, a, st, 0, 0, 0, true);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = 1;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                return launch_resident(pl, bwd_fast1d_il_entry(tpw), a, lay, a.sm_count, st, "bwd_fast1d_il");
            }
            a.pb.om_kind = pl->dev.om_kind;
            a.lik_table = in->lik_table;
        }
        if (alignedRows && !acc && tabOk && fast1d_mma_layout(pl, in->prog, true, a, lay, tpw)) {
            const int rc = prep_lik_table(pl, in, a, st, 0, 0, 0, true);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = 1;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                if (PassKernel k = bwd_fast1d_mma_entry(tpw, lay.nt)) return launch_resident(pl, k, a, lay, grid, st, "bwd_fast1d_mma");
                return fail("DMMA backward kernel missing");
            }
            a.pb.om_kind = pl->dev.om_kind;
            a.lik_table = in->lik_table;
        }
        int wsM = 0, wsML = 0;
        if (alignedRows && !acc && fast1d_ws_layout(pl, in->prog, true, a, lay, wsM, wsML)) {
            const int rc = prep_lik_table(pl, in, a, st, wsM, wsML, lay.nt - 32);
            if (rc < 0) return -1;
            if (rc == 0) {
                a.use_bulk = 1;
                long long grid = in->B;
                if (prep_sm_assign(pl, in, a, grid, st)) return -1;
                if (PassKernel k = bwd_fast1d_ws_entry(wsM, wsML, lay.nt)) return launch_resident(pl, k, a, lay, grid, st, "bwd_fast1d_ws");
                return fail("warp-specialised backward kernel missing");
            }
        }
    }
    if (prep_lik_table(pl, in, a, st)) return -1;
    const int M = kFastM;
    if (fast1d_layout(pl, in->prog, true, M, a, lay)) {
        a.use_bulk = alignedRows ? 1 : 0;
        long long grid = in->B;
        if (prep_sm_assign(pl, in, a, grid, st)) return -1;
        if (PassKernel k = bwd_fast1d_entry(M, lay.nt)) return launch_resident(pl, k, a, lay, grid, st, "bwd_fast1d");
    }
    a.halo = 0;
    {
        int C = 0, m0 = 16;
        const bool want = pl->opt.cluster2d || !resident_layout(pl, in->prog, true, false, a, lay);
        if (want && !pl->opt.force_stream && (uintptr_t)out->alpha_seq % 16 == 0 &&
            cluster2d_layout(pl, in->prog, flags, true, a, lay, C, m0)) {
            const int rc = launch_cluster(pl, bwd_cluster2d_entry(pl->opt.trace[0] != 0, m0), a, lay, in->B, C, st, "bwd_cluster2d");
            if (rc <= 0) return rc;
        }
    }
    bool fits = alignedRows && resident_layout(pl, in->prog, true, true, a, lay);
    if (!fits) fits = false;
    if (pl->opt.force_stream || (!fits && !resident_layout(pl, in->prog, true, false, a, lay))) {
        if (!stream_layout(pl, in->prog, a, lay)) return fail("grid / kernel radius too large for the stream backward kernel");
        return launch_stream(bwd_resident_entry(1024, true), pl, a, lay, in->B, st, "bwd_stream");
    }
    a.use_bulk = (fits && a.off_stage >= 0) ? 1 : 0;
    return launch_resident(pl, bwd_resident_entry(lay.nt, false), a, lay, in->B, st, "bwd_resident");
}

int blg_accumulate(blg_plan *pl, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream) {
    (void)flags;
    if (!pl || !in || !out || !out->alpha_seq || !out->avg || !in->log_weight) return fail("null argument");
    if (in->B <= 0 || in->T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_w(pl, in->B)) return -1;
    const int nt = 256;
    weights_kernel<<<(unsigned)((in->B + nt - 1) / nt), nt, 0, st>>>(in->log_weight, out->alive, in->B, pl->d_w);
    const long long count = in->T * (long long)pl->dev.G;
    const long long seqStride = out->seq_stride > 0 ? out->seq_stride : count;
    const long long rowStride = out->row_stride > 0 ? out->row_stride : in->T;
    const unsigned blocks = (unsigned)((count + nt - 1) / nt);
    if (out->row_scale)  // stateless: the factors are applied whenever the caller hands them (kernels pre-fill 1.0)
        accumulate_kernel<true><<<blocks, nt, 0, st>>>(out->alpha_seq, pl->d_w, in->B, count, out->avg, out->row_scale, in->T,
                                                       pl->dev.G, seqStride, rowStride);
    else
        accumulate_kernel<false><<<blocks, nt, 0, st>>>(out->alpha_seq, pl->d_w, in->B, count, out->avg, nullptr, in->T,
                                                        pl->dev.G, seqStride, rowStride);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_scale(blg_plan *pl, double *x, int64_t count, double factor, void *stream) {
    if (!pl || !x) return fail("null argument");
    if (count <= 0) return 0;
    const int nt = 256;
    scale_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(x, count, factor);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_wave_weights(blg_plan *pl, const double *log_evidence, const double *log_prior, int64_t B, double *shift,
                     double *avg, int64_t count, double *log_weight, void *stream) {
    if (!pl || !log_evidence || !log_prior || !shift || !log_weight) return fail("null argument");
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    wave_weights_kernel<<<1, 1024, 0, st>>>(log_evidence, log_prior, B, shift, pl->d_factor, log_weight);
    ++g_launches;
    if (avg && count > 0) {
        const int nt = 256;
        scale_dev_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, st>>>(avg, count, pl->d_factor);
        ++g_launches;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_rebase(blg_plan *pl, double *x, int64_t count, const double *from, const double *to, void *stream) {
    if (!pl || !x || !from || !to) return fail("null argument");
    if (count <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    rebase_factor_kernel<<<1, 1, 0, st>>>(from, to, pl->d_factor);
    const int nt = 256;
    scale_dev_kernel<<<(unsigned)((count + nt - 1) / nt), nt, 0, st>>>(x, count, pl->d_factor);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_finalize(blg_plan *pl, double *seq, int64_t T, double *means, uint32_t flags, void *stream) {
    if (!pl || !seq) return fail("null argument");
    if (T <= 0) return 0;
    const DevProblem &d = pl->dev;
    long long blocks = T < (long long)pl->num_sms * 8 ? T : (long long)pl->num_sms * 8;
    finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seq, T, d.G, d.n1, d.ndim, d.c0, d.c1, means,
                                                                         (flags & BLG_F_NORMALIZE_ROWS) ? 1 : 0);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_marginal(blg_plan *pl, const double *seq, int64_t T, int32_t axis, double *out, void *stream) {
    if (!pl || !seq || !out) return fail("null argument");
    const DevProblem &d = pl->dev;
    if (axis < 0 || axis >= d.ndim) return fail("axis out of range");
    if (T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d.ndim == 1) {
        CUDA_TRY(cudaMemcpyAsync(out, seq, (size_t)T * d.G * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const int nt = 256;
    const long long threads = axis == 0 ? T * d.n0 * 32 : T * d.n1;
    marginal_kernel<<<(unsigned)((threads + nt - 1) / nt), nt, 0, st>>>(seq, T, d.n0, d.n1, axis, out);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_time_average(blg_plan *pl, const double *seq, int64_t T, double *out, void *stream) {
    if (!pl || !seq || !out) return fail("null argument");
    if (T <= 0) return fail("empty sequence");
    const int nt = 256;
    time_average_kernel<<<(unsigned)((pl->dev.G + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(seq, T, pl->dev.G, out);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_share_apply(blg_plan *pl, const blg_inputs *in, const blg_outputs *out, const double *ratio, int64_t ratio_stride,
                    int64_t n_groups, const int32_t *cp_step, void *stream) {
    if (!pl || !in || !out || !ratio || !cp_step || !out->alpha_seq || !out->row_scale) return fail("null argument");
    if (in->B <= 0 || in->T <= 0) return 0;
    if (n_groups <= 0 || in->B % n_groups != 0) return fail("B must be n_cp * n_groups");
    cudaStream_t st = (cudaStream_t)stream;
    const DevProblem &d = pl->dev;
    const long long row = (long long)d.G;
    const long long seqStride = out->seq_stride > 0 ? out->seq_stride : in->T * row;
    const long long rowStride = out->row_stride > 0 ? out->row_stride : in->T;
    const long long ratioStride = ratio_stride > 0 ? ratio_stride : in->T * row;
    const double *lik = in->lik_table;
    if (d.om_kind != BLG_OM_TABLE) {  // likelihood rows of the series: the plan's shared table
        if (prep_steps(pl, in, st)) return -1;
        const long long count = in->T * row;
        if (count > pl->lik_cap) {
            if (pl->d_lik) CUDA_TRY(cudaFree(pl->d_lik));
            pl->d_lik = nullptr;
            pl->lik_cap = 0;
            CUDA_TRY(cudaMalloc(&pl->d_lik, (size_t)count * sizeof(double)));
            pl->lik_cap = count;
        }
        lik_table_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d, pl->d_steps, in->T, pl->d_lik);
        ++g_launches;
        lik = pl->d_lik;
    } else if (!lik) {
        return fail("lik_table required for BLG_OM_TABLE");
    }
    if (n_groups * in->T > 2147483647LL) return fail("too many rows for one blg_share_apply call");
    const int nCp = (int)(in->B / n_groups);
    const bool vec = d.G % 2 == 0 && seqStride % 2 == 0 && ratioStride % 2 == 0 && (uintptr_t)out->alpha_seq % 16 == 0 &&
                     (uintptr_t)ratio % 16 == 0 && (uintptr_t)lik % 16 == 0;
    for (int k0 = 0; k0 < nCp; k0 += kShareK) {
        const int nK = nCp - k0 < kShareK ? nCp - k0 : kShareK;
        if (vec)
            share_apply_kernel<true><<<(unsigned)(n_groups * in->T), 256, 0, st>>>(
                out->alpha_seq, seqStride, ratio, ratioStride, lik, in->T, d.G, d.lc_prod, out->row_scale, out->local_evidence,
                rowStride, out->alive, n_groups, cp_step, k0, nK);
        else
            share_apply_kernel<false><<<(unsigned)(n_groups * in->T), 256, 0, st>>>(
                out->alpha_seq, seqStride, ratio, ratioStride, lik, in->T, d.G, d.lc_prod, out->row_scale, out->local_evidence,
                rowStride, out->alive, n_groups, cp_step, k0, nK);
        ++g_launches;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int blg_mix(blg_plan *pl, const double *state, const double *weight, int64_t K, int64_t n, double *out, void *stream) {
    if (!pl || !state || !weight || !out) return fail("null argument");
    if (n <= 0) return 0;
    const int nt = 256;
    mix_kernel<<<(unsigned)((n + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(state, weight, K, n, out);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
