/*
 * blgrid.h -- C ABI of the B200-native grid forward-backward engine (libblgrid.so).
 *
 * The reference (christophmark/bayesloop) has NO foreign-function interface: its hot path is the
 * Python loop Study.fit (bayesloop/core.py:330-486) calling the plugin methods
 *     ObservationModel.processedPdf(grid, segment)            bayesloop/observationModels.py:35-56
 *     TransitionModel.computeForwardPrior(posterior, t)       bayesloop/transitionModels.py:49-60 ...
 *     TransitionModel.computeBackwardPrior(posterior, t)      bayesloop/transitionModels.py:62-63 ...
 * at exactly five call sites (core.py:375, :411, :455, :467, :2146/:2166).  This header is the boundary a
 * maintainer would bind instead (ctypes stub in INTEGRATION.md): the per-combo "transition program" is what
 * the plugin tree lowers to, the observation op-code replaces processedPdf/pdf, and one blg_forward /
 * blg_backward call replaces the whole per-time-step loop for a batch of B hyper-parameter combinations.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns 0 on success, <0 on error (text via blg_last_error()).
 *   - every array marked "device" is a pointer into the memory space of the implementation: CUDA device
 *     memory for libblgrid.so, host memory for the CPU oracle oracle/libblgrid_oracle.so, which exports
 *     the same symbols and exists ONLY as the parity checker for the tests.
 *   - the caller owns every buffer; the library allocates only small per-plan tables.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - the calls are stateless apart from scratch the plan grows: nothing a call leaves in the plan changes the
 *     result of a later call (blg_accumulate applies row_scale whenever the caller passes it).
 *   - grids are C-contiguous, last parameter fastest (np.meshgrid(..., indexing='ij'), core.py:169):
 *     cell g = i0 * n[1] + i1.  All floating point data is IEEE binary64 (the reference computes in
 *     float64 throughout).
 */
#ifndef BLGRID_H_
#define BLGRID_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLG_ABI_VERSION 1
#define BLG_MAX_OPS 16

/* Observation models evaluated on the device (replaces ObservationModel.pdf, observationModels.py). */
enum blg_om_kind {
    BLG_OM_POISSON = 1,       /* observationModels.py:502   lambda^k exp(-lambda)/k!                    (1-D) */
    BLG_OM_GAUSSIAN = 2,      /* observationModels.py:566   N(d; mean, std)                             (2-D) */
    BLG_OM_SCALED_AR1 = 3,    /* observationModels.py:892   N(d_t; r d_{t-1}, s sqrt(1-r^2))   seg=2    (2-D) */
    BLG_OM_AR1 = 4,           /* observationModels.py:830   N(d_t; r d_{t-1}, s)               seg=2    (2-D) */
    BLG_OM_WHITE_NOISE = 5,   /* observationModels.py:767   N(d; 0, std)                                (1-D) */
    BLG_OM_GAUSSIAN_MEAN = 6, /* observationModels.py:705   N(d[0]; mean, d[1]); data has 2 columns     (1-D) */
    BLG_OM_LAPLACE = 7,       /* observationModels.py:635   exp(-|d-mean|/b)/(2b)                       (2-D) */
    BLG_OM_BERNOULLI = 8,     /* observationModels.py:430   p or 1-p                                    (1-D) */
    BLG_OM_TABLE = 100        /* user plugin: likelihood table lik[T][G] precomputed through processedPdf     */
};

/* Transition operators (replace TransitionModel.computeForwardPrior / computeBackwardPrior). */
enum blg_op_kind {
    BLG_OP_GRW = 1,      /* transitionModels.py:96-115  gaussian_filter1d(axis, sigma_n), reflect, truncate 4    */
    BLG_OP_REGIME = 2,   /* transitionModels.py:394-412 p = max(p, limit); p /= sum(p)                           */
    BLG_OP_RESET = 3,    /* transitionModels.py:289-314 (ChangePoint), :339-360 (Independent), :788-815 (Serial): */
                         /*                             p = reset_base * param   (reset_base sums to 1)          */
    BLG_OP_NOTEQUAL = 4  /* transitionModels.py:450-471 p = max(p)-p; p/=sum; p = max(p, limit); p/=sum         */
};

enum blg_flags {
    BLG_F_EVIDENCE_ONLY = 1u << 0,    /* core.py:340  do not store alpha[t] (alpha_seq may be NULL)               */
    BLG_F_INIT_STATE = 1u << 1,       /* start from init_state[b][G] instead of prior (OnlineStudy, core.py:2166) */
    BLG_F_TRANSITION_FIRST = 1u << 2, /* apply the transition program BEFORE the first likelihood (core.py:2166)  */
    BLG_F_SAVE_STATE = 1u << 3,       /* write the last normalised alpha to final_state[b][G]                     */
    BLG_F_ACCUMULATE = 1u << 4,       /* backward: avg[t][g] += exp(log_weight[b]) * max(post, 1e-300)            */
                                      /* (core.py:1358-1366) instead of overwriting alpha_seq with the posterior  */
    BLG_F_NORMALIZE_ROWS = 1u << 5,   /* finalize: divide each [G] row by its sum first (core.py:1379-1382)       */
    BLG_F_RAW_ALPHA = 1u << 6,        /* forward: the rows of alpha_seq may be left UNNORMALISED (each row scaled */
                                      /* by a positive factor): valid only as the input of blg_backward, which is */
                                      /* scale-free per row (core.py:436-441 renormalises alpha*beta)             */
    BLG_F_RAW_POSTERIOR = 1u << 7,    /* backward: the smoothed rows may be left unnormalised; row_scale[b][t]    */
                                      /* (required) receives the factor that normalises row t of combo b (1.0 if  */
                                      /* the implementation normalised the row itself).  blg_accumulate multiplies */
                                      /* by row_scale whenever it is non-NULL; blg_finalize(NORMALIZE_ROWS) normalises B = 1 */
    BLG_F_SEPARABLE_ROWS = 1u << 8    /* forward, caller's promise about the program: in every row the operators    */
                                      /* that are active at the steps of this call are GRWs on distinct axes,      */
                                      /* optionally followed by ONE REGIME, or a single RESET.  Lets an              */
                                      /* OnlineStudy step (T = 1, core.py:2157-2175) run tiled over the whole GPU;  */
                                      /* a hint only -- results do not depend on it                                 */
};

/* Static description of the grid and the observation model (host pointers; copied by blg_plan_create). */
typedef struct blg_problem {
    int32_t ndim;            /* 1 or 2 (len(observationModel.parameterNames))                                   */
    int32_t n[2];            /* gridSize (core.py:157); n[1] = 1 when ndim == 1                                   */
    const double *coords[2]; /* HOST: marginalGrid (core.py:156): parameter values per axis                      */
    double lattice[2];       /* latticeConstant (core.py:161-166); 1.0 for the unused axis                       */
    int32_t om_kind;         /* enum blg_om_kind                                                                */
    int32_t seg_len;         /* observationModel.segmentLength (preprocessing.py:14-26)                         */
    int32_t n_cols;          /* columns of multi-dimensional data (observationModels.py:49-50); 1 for 1-D data   */
    int32_t reserved;
} blg_problem;

/* Transition program: n_ops operators applied in order at every time step, each with a per-combo parameter
 * and a per-combo active window (half-open ranges of the 0-based time-step index i whose time stamp is handed
 * to the transition model).  The forward pass applies op k after step i iff f_lo <= i < f_hi; the backward pass
 * applies it at step i iff b_lo <= i < b_hi (computeBackwardPrior(p, t) == computeForwardPrior(p, t-1),
 * transitionModels.py:62-63; Serial model selection transitionModels.py:767-786). */
typedef struct blg_program {
    int32_t n_ops;
    int32_t kind[BLG_MAX_OPS]; /* enum blg_op_kind                                                             */
    int32_t axis[BLG_MAX_OPS]; /* GRW: grid axis (parameterNames.index(target), transitionModels.py:107)         */
    int32_t max_radius[BLG_MAX_OPS]; /* HOST knowledge: upper bound of radius[.][k] over the B combos of the call;   */
                               /*   sizes the shared-memory weight tables without a device round trip          */
    const double *param;       /* device [B][n_ops]: GRW sigma_n = sigma/lattice[axis]; REGIME/NOTEQUAL limit =   */
                               /*   10^log10pMin * prod(lattice); RESET scale (prod(lattice) or 1)               */
    const int32_t *radius;     /* device [B][n_ops]: GRW radius int(4*sigma_n+0.5) (scipy _filters.py:747); else 0 */
    const int32_t *window;     /* device [B][n_ops][4]: f_lo, f_hi, b_lo, b_hi                                   */
    const int32_t *order;      /* device [B] or NULL: permutation of the combos in descending cost (sum of radii); */
                               /*   scheduling hint only -- results do not depend on it                           */
    const int32_t *sm_assign;  /* device [sm_count][sm_slots] or NULL: combos pre-assigned to each SM (-1 = empty),  */
    int32_t sm_count;          /*   balanced by the caller; persistent CTAs claim the slots of the SM they run on.   */
    int32_t sm_slots;          /*   Scheduling hint only; every combo must appear exactly once.                     */
} blg_program;

typedef struct blg_inputs {
    int64_t T;                /* number of data segments len(formattedData) (core.py:349)                         */
    int64_t B;                /* hyper-parameter combinations in this call (rows of hyperGridValues, core.py:1163) */
    const double *data;       /* device [T + seg_len - 1][n_cols] raw series; NaN = missing (observationModels:53) */
    const double *prior;      /* device [G] initial alpha = Study._computePrior() (core.py:184-265)               */
    const double *reset_base; /* device [G] OM prior on the grid normalised to sum 1 (transitionModels.py:300-310), */
                              /*   or NULL when the program holds no RESET op                                     */
    const double *lik_table;  /* device [T][G], only for BLG_OM_TABLE                                             */
    blg_program prog;
    const double *log_weight; /* device [B] backward+ACCUMULATE: logE_b + log hyperprior_b - shift                 */
    const double *init_state; /* device [B][G], only with BLG_F_INIT_STATE                                        */
    const double *alpha_src;  /* blg_backward, optional: read the filtering rows HERE ([B] sequences src_stride doubles */
    int64_t src_stride;       /*   apart, 0 = T * G) and write the smoothed rows to alpha_seq -- the out-of-place form,   */
                              /*   for filtering rows that several calls share (change-point prefix sharing)            */
} blg_inputs;

typedef struct blg_outputs {
    double *log_evidence;   /* device [B]     logEvidence incl. + log prod(lattice) (core.py:403,:417); -inf if dead */
    double *local_evidence; /* device [B][T]  forward: norm*prod(lattice) (core.py:404); backward overwrites with    */
                            /*                1/(sum(post/lik)*prod(lattice)) (core.py:463)                        */
    int32_t *alive;         /* device [B]     1 = finished; 0 = zero-norm abort in the forward pass (core.py:388-400); */
                            /*                -1 = zero-norm abort in the backward pass (core.py:440-452)            */
    double *alpha_seq;      /* device [B][T][G] forward: filtering distributions (core.py:408); backward without     */
                            /*                ACCUMULATE: overwritten by the smoothed posteriors (core.py:436-441)  */
    double *avg;            /* device [T][G]  backward with ACCUMULATE: running weighted sum (caller zero-fills)     */
    double *final_state;    /* device [B][G]  only with BLG_F_SAVE_STATE                                            */
    double *row_scale;      /* device [B][T]  backward with BLG_F_RAW_POSTERIOR: written; accumulate: read if non-NULL */
    int64_t seq_stride;     /* doubles between the sequences of consecutive combos in alpha_seq; 0 = T * G (packed).       */
    int64_t row_stride;     /* doubles between the rows of consecutive combos in local_evidence and row_scale; 0 = T.     */
                            /* Strides let a call work on a WINDOW of time steps of sequences that live in a larger       */
                            /* [B][T_full][G] buffer (pointer offset + T of the window): the sub-calls of the change-point */
                            /* prefix sharing run in place this way.                                */
} blg_outputs;

typedef struct blg_plan blg_plan;

int blg_version(void);
const char *blg_last_error(void);

/* "cuda:sm_100a" for libblgrid.so, "cpu-oracle" for oracle/libblgrid_oracle.so */
const char *blg_backend(void);

int blg_plan_create(const blg_problem *problem, blg_plan **plan);
void blg_plan_destroy(blg_plan *plan);

/* Dispatch / tuning option of a plan, by name (e.g. "force_stream", "cluster2d", "cluster2d_c", "no_ws", "online2d",
 * "online2d_async", "verbose"; the full list is the table kOptNames in bayesloop_b200/csrc/api.cu).  Options are
 * resolved here and at plan creation (which reads the BLG_* debugging variables of the environment ONCE): no
 * per-call path looks at process state.  They select WHICH kernel family runs, never what it computes: results
 * agree to rounding whatever the options.  Unknown names return -1.  The CPU oracle accepts and ignores them. */
int blg_plan_set_option(blg_plan *plan, const char *name, int64_t value);

/* Forward filter for B combos: replaces the loop core.py:372-411 (+ :417). */
int blg_forward(blg_plan *plan, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream);

/* Backward smoother for B combos: replaces the loop core.py:434-470, and with BLG_F_ACCUMULATE also the
 * evidence-weighted averaging of core.py:1358-1366 (linear-space form). */
int blg_backward(blg_plan *plan, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream);

/* Evidence-weighted sum of the STORED sequences (HyperStudy with forwardOnly=True averages the filtering
 * distributions, core.py:1353-1366): avg[t][g] += sum_b exp(log_weight[b]) * max(alpha_seq[b][t][g], 1e-300)
 * over the combos with alive[b] == 1.  Deterministic summation order (b ascending). */
int blg_accumulate(blg_plan *plan, const blg_inputs *in, const blg_outputs *out, uint32_t flags, void *stream);

/* x[i] *= factor, i < count.  Re-bases the running average when the reference log-weight of the sweep moves
 * (the streaming form of np.logaddexp in core.py:1363; also exp(m_r - M) between ranks). */
int blg_scale(blg_plan *plan, double *x, int64_t count, double factor, void *stream);

/* Averaging weights of one wave of a sweep without a host round trip -- the streaming form of np.logaddexp in
 * core.py:1358-1366.  With lw[b] = log_evidence[b] + log_prior[b], top = max of the finite lw and
 * new = max(shift[0], top):  if shift[0] is finite and new > shift[0] the running sum is re-based,
 * avg[i] *= exp(shift[0] - new) for i < count;  log_weight[b] = lw[b] - new (-inf where lw is not finite);
 * shift[0] = new.  shift: device [1], -inf before the first wave; avg may be NULL with count = 0. */
int blg_wave_weights(blg_plan *plan, const double *log_evidence, const double *log_prior, int64_t B, double *shift,
                     double *avg, int64_t count, double *log_weight, void *stream);

/* x[i] *= exp(from[0] - to[0]), i < count, with DEVICE scalars; nothing happens while from[0] is -inf (nothing was
 * accumulated).  Brings a rank's running average onto the common reference log-weight after an all-reduce(max) of
 * the ranks' shifts (merge of core.py:1339-1340) without reading the scalars back. */
int blg_rebase(blg_plan *plan, double *x, int64_t count, const double *from, const double *to, void *stream);

/* Row normalisation (optional) and posterior means of a [T][G] sequence: core.py:1379-1382, :480-483,
 * :1416-1419.  means is device [ndim][T]. */
int blg_finalize(blg_plan *plan, double *seq, int64_t T, double *means, uint32_t flags, void *stream);

/* Post-processing queries on a device-resident [T][G] sequence without downloading it:
 *   blg_marginal      out[t][i] = sum over the OTHER grid axis of seq[t][.]   (core.py:915, :979-980); out device
 *                     [T][n[axis]]; a 1-D grid is copied
 *   blg_time_average  out[g] = (1/T) sum_t seq[t][g]                          (core.py:886); out device [G]
 * Fixed summation order (deterministic). */
int blg_marginal(blg_plan *plan, const double *seq, int64_t T, int32_t axis, double *out, void *stream);
int blg_time_average(blg_plan *plan, const double *seq, int64_t T, double *out, void *stream);

/* Change-point prefix sharing.  A ChangePoint erases the history at tChange
 * (transitionModels.py:300-312), so all combinations that differ only in tChange share the filtering recursion BEFORE
 * it and the backward message AFTER it: both come from ONE change-point-free run per group of combinations.
 * The message itself needs no new pass: blg_backward on a sequence of ONES returns alpha * beta = beta row by row.
 *   blg_share_apply   the B = n_cp * n_groups combos of the call are laid out change-point-major (combo k * n_groups + g
 *                     = group g with its reset after step cp_step[k]).  For every row t > cp_step[k] of every combo:
 *                     u = alpha_seq[b][t][.] * ratio[g][t][.] in place -- the smoothed row up to its scale, which
 *                     goes to row_scale[b][t] = 1 / sum(u) -- and local_evidence[b][t] = 1 / (sum(post / lik) *
 *                     prod(lattice))  (core.py:436-441, :463).  ratio: device, the backward message (any positive
 *                     factor per row) of the n_groups change-point-free runs, `ratio_stride` doubles apart
 *                     (0 = T * G); the message and the likelihood row of (g, t) are read ONCE for all change-points.
 *                     cp_step: device [n_cp] int32.  Combos with alive[b] != 1 are skipped; a zero row sum sets
 *                     alive[b] = -1 (core.py:440-452). */
int blg_share_apply(blg_plan *plan, const blg_inputs *in, const blg_outputs *out, const double *ratio,
                    int64_t ratio_stride, int64_t n_groups, const int32_t *cp_step, void *stream);

/* Weighted sum of K rows: out[j] = sum_k weight[k] * state[k][j], j < n.  Used for the OnlineStudy
 * marginalisations (core.py:2195-2197, :2212; n = G) and for the averaged local evidence of a HyperStudy
 * (core.py:1410; K = B, n = T).  state device [K][n], weight device [K], out device [n]. */
int blg_mix(blg_plan *plan, const double *state, const double *weight, int64_t K, int64_t n, double *out,
            void *stream);

/* Number of kernels launched by this library since load (bench.py's gpu_launches claim). */
int64_t blg_launch_count(void);

/* Name of the kernel family the last blg_forward / blg_backward call launched ("fwd_fast1d_ws", "bwd_cluster2d",
 * "fwd_stream", ...; "oracle" for the CPU checker): lets the tests assert which device path produced a result. */
const char *blg_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* BLGRID_H_ */
