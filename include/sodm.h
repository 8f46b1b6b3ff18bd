/* sodm.h -- C ABI of the B200-native ensemble simulator behind pySODM's model API.
 *
 * pySODM (twallema/pySODM v0.2.8) has no plugin / FFI interface of its own: the drop-in boundary is
 * its Python class API.  This library is what a maintainer binds (ctypes) at
 * the two seams of the hot path:
 *
 *   seam S (simulate):  the loop over draws inside `JumpProcess.sim` / `ODE.sim`
 *                       src/pySODM/models/base.py:558-566 and :950-957
 *                       (N x `_mp_sim_single` -> `_sim_single` -> `_solve_discrete` | `solve_ivp`)
 *                       -> sodm_jump_run / sodm_ode_run          (one call for all N trajectories)
 *   seam L (posterior): `log_posterior_probability.compute_log_likelihood` + `ll_*`
 *                       src/pySODM/optimization/objective_functions.py:235-288, :550-666
 *                       -> sodm_loglik                           (one call for a whole population)
 *   output contract:    `_output_to_xarray_dataset` + concat along "draws", base.py:969-1042, :569-571
 *                       -> sodm_to_draws_major                   ([state][time][traj] -> [traj][time][state dims])
 *
 * Conventions
 *   - every function returns 0 on success or a negative SODM_E* code; sodm_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - no allocation and no ownership transfer: every pointer marked (device) is a CALLER-owned
 *     device buffer (the Python layer allocates them with PyTorch) that must stay alive until the
 *     work queued on `stream` has completed; pointers marked (host) are read before the call returns;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *     no hidden synchronisation; no global mutable state; one call never spans devices -- the caller
 *     shards trajectories over GPUs (rows [traj_id0, traj_id0 + n_traj));
 *   - trajectory-major arrays are structure-of-arrays: element (row r, trajectory i) lives at
 *     base[r * ld + i], so a warp of consecutive trajectories touches consecutive addresses.
 */
#ifndef SODM_H
#define SODM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SODM_ABI_VERSION 5

/* error codes */
#define SODM_OK            0
#define SODM_EINVAL       -1   /* bad argument */
#define SODM_EMODEL       -2   /* unknown model / model does not support the request */
#define SODM_ECUDA        -3   /* CUDA runtime error */
#define SODM_EUNSUPPORTED -4   /* valid request that this build does not implement */

/* model kinds */
#define SODM_KIND_JUMP 0
#define SODM_KIND_ODE  1

/* element types of output buffers */
#define SODM_I32 0
#define SODM_F32 1
#define SODM_F64 2

/* per-trajectory status bits (int32 status[n_traj]) */
#define SODM_ST_OK             0
#define SODM_ST_BAD_PROB       1   /* transition probability NaN or outside [0,1]: NumPy's binomial raises
                                      ValueError there (base.py:303) */
#define SODM_ST_NONFINITE      2   /* non-finite model output (objective_functions.py:264-265 raises) */
#define SODM_ST_STEP_TOO_SMALL 4   /* RK45: step size fell below 10 ulp(t) (scipy: status -1) */
#define SODM_ST_MAX_STEPS      8   /* step / event budget exhausted */
#define SODM_ST_TRUNCATED_Y0  16   /* non-integer initial count truncated to an integer */

/* compat flags of sodm_jump_args.compat */
#define SODM_COMPAT_STAY_QUIRK      1u  /* tau-leap: reference's stay probability 1 - sum_{i<m-1} p_i (base.py:339-342) */
#define SODM_COMPAT_SSA_FIRST_MATCH 2u  /* SSA: reaction looked up by propensity VALUE, first match wins (base.py:254) */
#define SODM_COMPAT_REFERENCE       3u
#define SODM_TUNE_QUEUE            256u  /* tuning, not semantics: the queue sampler (the draws of a warp's trajectories are
                                           items any lane may work on, batched by regime; counter-based words) instead of the
                                           per-lane walk (csrc/jump_run.cuh); same law, different random streams */

#define SODM_TUNE_LATENCY          512u  /* tuning: the 96-register build of the tau-leap kernel (no spills, 5 blocks per SM) for
                                           launches of about one wave or less (csrc/jump_kernels.cuh); same law, but about 3 in
                                           10^5 trajectories round differently than in the default build */

/* methods */
#define SODM_JUMP_TAU_LEAP 0
#define SODM_JUMP_SSA      1
#define SODM_ODE_RK45      0   /* scipy.integrate.solve_ivp(method='RK45') step for step (base.py:855) */
#define SODM_ODE_EULER     1   /* explicit Euler with step tau + interpolation (base.py:800-819) */
#define SODM_ODE_RK4       2   /* classical fixed-step RK4 (extension; the reference has none) */
#define SODM_ODE_RK23      3   /* scipy.integrate.solve_ivp(method='RK23') step for step (validation.py:158 lists it) */
#define SODM_ODE_DOP853    4   /* scipy.integrate.solve_ivp(method='DOP853') step for step, incl. the three extra evaluations of
                                  its dense output on steps that contain a grid point (validation.py:158) */

#define SODM_ODE_RADAU     5   /* scipy.integrate.solve_ivp(method='Radau') step for step: Radau IIA with scipy's finite-difference
                                  Jacobian (validation.py:158); thread-per-trajectory models of at most 32 equations */
#define SODM_ODE_BDF       6   /* scipy.integrate.solve_ivp(method='BDF') step for step: variable-order NDF, same Jacobian */

#define SODM_MAX_PARAM 64      /* flat parameter elements per model */
#define SODM_MAX_PEERS 7       /* peer destinations of a jump launch (an 8-GPU box) */
#define SODM_MAX_STATE 64      /* flat state elements of a jump model */

typedef struct sodm_model_info {
    int32_t kind;          /* SODM_KIND_* */
    int32_t n_state;       /* flat state elements (states in declaration order, C-order ravel; utils.py:45-55) */
    int32_t n_param;       /* flat parameter elements, in the order documented per model in DESIGN.md */
    int32_t n_channel;     /* jump models: transition channels (state, transition, element) */
    int32_t tdpf_offset;   /* first flat parameter element that is read from the time-dependent table */
    int32_t tdpf_count;    /* number of such elements (0: none) */
    int32_t supports_ssa;
    int32_t reserved;
} sodm_model_info;

int         sodm_abi_version(void);
const char* sodm_last_error(void);
/* Selects the device for the calling thread and checks that it is an sm_100 part. */
int         sodm_init(int device);
/* "sir_jump", "sir_si_jump", "influenza_1718_jump", "sir_ode", "sir_rate_ode", "sir_si_ode", "ppbb_ode",
 * "packed_pfr_ode"; < 0 if unknown. */
int         sodm_model_id(const char* name);
/* model_arg: packed_pfr_ode = number of axial nodes nx; ignored otherwise. */
int         sodm_model_info_get(int model, int model_arg, sodm_model_info* out);

/* ---- seam S, jump processes: replaces base.py:421-440 (+ :206-348) for n_traj trajectories ------------ */
typedef struct sodm_jump_args {
    int32_t  model;
    int32_t  method;            /* SODM_JUMP_* */
    uint32_t compat;            /* SODM_COMPAT_* (SODM_COMPAT_REFERENCE reproduces the reference's laws) */
    int32_t  out_dtype;         /* SODM_I32 only when every output time is hit exactly by a step (tau = 2^-k) */
    int64_t  n_traj;
    int64_t  traj_id0;          /* global id of the first trajectory: RNG stream = Philox(seed, traj_id0 + i) */
    uint64_t seed;
    int64_t  ld_in;             /* row stride of y0 / var_values, >= n_traj */
    int64_t  ld_out;            /* row stride of out, >= n_traj */
    const double*  y0;          /* (device) [n_state][ld_in], or [n_state] if y0_broadcast */
    int32_t  y0_broadcast;
    int32_t  n_var;             /* number of parameter elements that vary per trajectory */
    const double*  par;         /* (host)   [n_param] value of every parameter element (used unless overridden) */
    const int32_t* var_index;   /* (host)   [n_var] flat parameter element of each varying row */
    const double*  var_values;  /* (device) [n_var][ld_in] */
    const double*  tdpf_table;  /* (device) [n_tdpf_values][tdpf_count]; NULL if the model has tdpf_count == 0 */
    const int32_t* tdpf_index;  /* (device) [n_steps] table row in force during tau-leap step k (t_k = t0 + k*tau
                                   accumulated in FP64 as base.py:433).  SSA: the functions are evaluated at event
                                   times (base.py:376-380): entry 0 = row in force at t0 itself (first event), entry
                                   1+k = row in force on the open interval (t0 + k*tau, t0 + (k+1)*tau), on which the
                                   tabulated functions must be constant */
    int32_t  n_tdpf_values;
    int32_t  n_tdpf_index;      /* length of tdpf_index; must be >= the number of steps (see sodm_jump_n_steps) */
    double   t0;                /* first output time; outputs at t0, t0+1, ..., t0+n_out-1 (base.py:449) */
    int32_t  n_out;
    int32_t  n_sel;             /* 0: all states, else number of entries of state_sel */
    double   tau;               /* tau-leap step; SSA: width of a tdpf_index interval (ignored without a table) */
    const int32_t* state_sel;   /* (host)   [n_sel] flat state elements to output, in output-row order */
    void*    out;               /* (device) [n_rows][n_out][ld_out] of out_dtype, n_rows = n_sel ? n_sel : n_state */
    int32_t* status;            /* (device) [n_traj] SODM_ST_* bits, may be NULL */
    int64_t  max_events;        /* SSA: event budget per trajectory (0 = 2^62) */
    void*    stream;
    void* const* out_peers;     /* (host) [n_out_peers] device pointers: further destinations with the layout of `out`
                                   (buffers of PEER GPUs mapped into this process, see sodm_peer_enable).  Every output
                                   element is stored to `out` and to each of them by the simulation kernel itself: the
                                   "final gather of outputs" over NVLink happens while the trajectories run, there is no
                                   separate collective.  NULL / 0: plain local output. */
    int32_t  n_out_peers;       /* <= SODM_MAX_PEERS */
    int32_t  reserved2;
} sodm_jump_args;

int sodm_jump_run(const sodm_jump_args* args);
/* Lets kernels of the current device store into memory of `peer_device` (cudaDeviceEnablePeerAccess; "already enabled" and
 * the device itself are not errors).  The peer buffer itself is mapped by the host framework (CUDA IPC handle). */
int sodm_peer_enable(int peer_device);
/* Peer-visible buffers for `out_peers`.  The one exception to "no allocation": a CUDA IPC handle names a whole cudaMalloc
 * allocation, so the library allocates these itself.  Rank r calls sodm_peer_alloc on its device, sends the 64 handle bytes
 * to the other processes (any host channel), and every other process maps the buffer with sodm_peer_open on ITS current
 * device (peer access is enabled by the mapping).  Close every mapping before the owner frees the buffer. */
#define SODM_PEER_HANDLE_BYTES 64
int sodm_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle /* [SODM_PEER_HANDLE_BYTES] out */);
int sodm_peer_open(const unsigned char* handle, void** ptr);
int sodm_peer_close(void* ptr);
/* The same gather by DMA: copies `rows` rows of `width_bytes` (this rank's columns of a [row][time][n_total] buffer) from `src`
 * to `dst` (a mapped peer buffer) with the copy engines -- no SM is involved, so it overlaps a running simulation kernel
 * completely.  Asynchronous on `stream`. */
int sodm_peer_copy2d(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes, int64_t width_bytes,
                     int64_t rows, void* stream);
int sodm_peer_free(void* ptr);
/* Number of tau-leap steps the reference's loop `while t < t1: t = t + tau` takes (base.py:428-434). */
int64_t sodm_jump_n_steps(double t0, int32_t n_out, double tau);

/* ---- seam S, ODEs: replaces base.py:855 (solve_ivp) / :800-819 (Euler) for n_traj parameter draws ------ */
typedef struct sodm_ode_args {
    int32_t  model;
    int32_t  model_arg;         /* packed_pfr_ode: nx */
    int32_t  method;            /* SODM_ODE_* */
    int32_t  out_dtype;         /* SODM_F32 | SODM_F64 */
    int64_t  n_traj;
    int64_t  ld_in;
    int64_t  ld_out;
    const double*  y0;          /* (device) [n_state][ld_in] or [n_state] */
    int32_t  y0_broadcast;
    int32_t  n_var;
    const double*  par;         /* (host) [n_param] */
    const int32_t* var_index;   /* (host) [n_var] */
    const double*  var_values;  /* (device) [n_var][ld_in] */
    double   t0;
    int32_t  n_out;             /* reference grid t0 .. t0+n_out-1 */
    int32_t  out_stride;        /* keep grid points 0, out_stride, 2*out_stride, ... (1 = reference grid) */
    int32_t  out_last_only;     /* 1: keep only the final grid point (the tutorial's isel(time=-1)) */
    int32_t  reserved;
    double   rtol;              /* RK45 (reference default 1e-4, base.py:877) */
    double   atol;              /* RK45 (scipy default 1e-6: the reference does not pass atol) */
    double   tau;               /* Euler / RK4 step */
    void*    out;               /* (device) [n_state][n_kept][ld_out] */
    int32_t* nfev;              /* (device) [n_traj] right-hand-side evaluations (scipy's sol.nfev), may be NULL */
    int32_t* status;            /* (device) [n_traj], may be NULL */
    int64_t  max_steps;         /* 0 = 2^31-1 */
    void*    stream;
} sodm_ode_args;

int sodm_ode_run(const sodm_ode_args* args);
/* number of grid points kept for (n_out, out_stride, out_last_only) */
int32_t sodm_ode_n_kept(int32_t n_out, int32_t out_stride, int32_t out_last_only);

/* ---- seam L: replaces objective_functions.py:250-288 + :550-666 for n_traj trajectories --------------- */
#define SODM_LL_POISSON   0
#define SODM_LL_NORMAL    1
#define SODM_LL_LOGNORMAL 2
#define SODM_LL_NEGBIN    3
/* ll_arg broadcasting (sigma / alpha): one value, one per data column, one per data point */
#define SODM_ARG_NONE   0
#define SODM_ARG_SCALAR 1
#define SODM_ARG_COLUMN 2
#define SODM_ARG_FULL   3

typedef struct sodm_loglik_args {
    int32_t  kind;              /* SODM_LL_* */
    int32_t  traj_dtype;        /* SODM_I32 | SODM_F32 | SODM_F64 */
    int64_t  n_traj;
    int64_t  ld;                /* row stride of traj */
    const void*    traj;        /* (device) [n_row][n_out][ld] simulator output */
    int32_t  n_out;             /* grid points t0 .. t0+n_out-1 */
    int32_t  n_col;             /* data columns */
    const int32_t* col_ptr;     /* (device) [n_col+1] CSR offsets into col_rows: model output rows summed into a column
                                   (the sum over dimensions absent from the data, objective_functions.py:250-252) */
    const int32_t* col_rows;    /* (device) [col_ptr[n_col]] */
    const double*  col_weights; /* (device) [col_ptr[n_col]] weight of each row in its column, or NULL = 1: a LINEAR user
                                   `aggregation_function` (objective_functions.py:245-248) folded into the column sums */
    double   t0;
    int32_t  n_t;               /* data time points */
    int32_t  arg_mode;          /* SODM_ARG_* */
    const double*  t_data;      /* (device) [n_t] data times in simulation time units (linear interpolation, :254) */
    const double*  ydata;       /* (device) [n_t][n_col] row-major, time first (:259, :276-278) */
    const double*  ll_arg;      /* (device) sigma / alpha per arg_mode */
    double   weight;            /* dataset weight (:266-269) */
    double*  lp;                /* (device) [n_traj] log-probability, ACCUMULATED (+=) */
    int32_t* status;            /* (device) [n_traj] |= SODM_ST_NONFINITE when the model output holds NaN; may be NULL */
    void*    stream;
} sodm_loglik_args;

int sodm_loglik(const sodm_loglik_args* args);

/* *out |= status[0] | status[1] | .. | status[n-1] (SODM_ST_* flags of a whole launch; the caller zeroes *out).  The front
 * end turns the flags into the reference's exceptions (base.py:303 NumPy's ValueError; objective_functions.py:264-265). */
int sodm_status_or(const int32_t* status /* (device) [n] */, int64_t n, int32_t* out /* (device) [1] */, void* stream);

/* ---- output contract: [row][time][traj] (device layout) -> [traj][time][row] per state variable -------- */
/* For one state variable whose rows are r0 .. r0+n_dim-1:  dst[(i*n_t + j)*n_dim + d] = src[((r0+d)*n_t + j)*ld + i]
 * (base.py:1009-1010 moveaxis + :569-571 concat along "draws").  dst may be device or pinned host memory that is
 * device-accessible; dtype converts src_dtype -> dst_dtype (I32/F32/F64). */
int sodm_to_draws_major(const void* src, int32_t src_dtype, int64_t ld, int32_t r0, int32_t n_dim, int32_t n_t,
                        int64_t n_traj, void* dst, int32_t dst_dtype, void* stream);

/* ---- ensemble post-processing ----------------------------------------------------------------------------
 * What the reference's workflows do to a finished ensemble on the host, done where the ensemble lives:
 *   add_poisson_noise / add_gaussian_noise / add_negative_binomial_noise
 *                       src/pySODM/optimization/utils.py:11-33, :35-66, :68-95        -> sodm_noise
 *   out[state].mean(dim='draws')       tutorials/influenza_1718/calibration.py:245   -> sodm_draws_sum
 *   out[state].quantile(dim='draws', q=...)  tutorials/influenza_1718/calibration.py:246 -> sodm_draws_order_stats
 * All three read the simulators' [cell][trajectory] layout (cell = row * n_time + time; element (c, i) at
 * base[c * ld + i]).  Random streams are keyed by (seed, traj_id0 + i, c): the noisy ensemble does not depend on how
 * trajectories are sharded over GPUs. */
#define SODM_NOISE_POISSON       0    /* np.random.poisson(x)                                     utils.py:30    */
#define SODM_NOISE_GAUSSIAN_ABS  1    /* np.random.normal(x, sigma)            relative=False     utils.py:62    */
#define SODM_NOISE_GAUSSIAN_REL  2    /* np.random.normal(x, |sigma * x|)      relative=True      utils.py:64    */
#define SODM_NOISE_NEGBIN        3    /* np.random.negative_binomial(1/alpha, (1/alpha)/(x + 1/alpha))  utils.py:91-92 */

/* bits OR-ed into *status by sodm_noise */
#define SODM_NOISE_ST_BAD_MEAN   1    /* negative / NaN / infinite model output under Poisson or NB noise
                                         (NumPy: "lam < 0 or lam contains NaNs", "p < 0, p > 1 or p is NaN")   */

typedef struct sodm_noise_args {
    int32_t     kind;        /* SODM_NOISE_*                                                                      */
    double      param;       /* sigma (Gaussian, >= 0) or alpha (negative binomial, > 0); ignored for Poisson       */
    const void* in;          /* (device) [n_cells][ld] model output                                                */
    int32_t     in_dtype;    /* SODM_I32 | SODM_F32 | SODM_F64                                                     */
    void*       out;         /* (device) [n_cells][ld]; may be `in` itself when the dtypes are equal               */
    int32_t     out_dtype;
    int64_t     n_cells, n_traj, ld;
    int64_t     traj_id0;    /* global id of trajectory 0 of this shard                                            */
    uint64_t    seed;
    int32_t*    status;      /* (device) one int32, OR-ed with SODM_NOISE_ST_*; may be NULL                        */
    void*       stream;
} sodm_noise_args;
int sodm_noise(const sodm_noise_args* args);

/* sum[c] = sum_i in[c * ld + i] in FP64 (int32 input: accumulated in int64, i.e. exact). */
int sodm_draws_sum(const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld,
                   double* sum /* (device) [n_cells] */, void* stream);

/* out[j * n_cells + c] = the ranks[j]-th smallest (0-based) of the n_traj values of cell c -- the order statistics
 * np.quantile interpolates between; the caller applies NumPy's `_lerp`.  A cell holding a NaN yields NaN for every
 * rank, as np.quantile does.  `ranks` is a HOST array of at most SODM_MAX_RANKS entries. */
#define SODM_MAX_RANKS 4
int sodm_draws_order_stats(const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld,
                           const int64_t* ranks, int32_t n_ranks, double* out /* (device) [n_ranks][n_cells] */,
                           void* stream);

/* Order statistics of an ensemble SHARDED by trajectory over several GPUs: the same radix search in phases whose small
 * per-cell tables the caller combines across ranks (one all-reduce each) -- trajectories never move, histograms do.
 *   state arrays (device, int64 unless noted; keys are order-preserving integers with the sign bit flipped so that signed
 *   MIN / MAX order them):  lo[n_cells], hi[n_cells], nan_flag[n_cells] (int32), settled[n_cells] (int32, zero-initialised),
 *   prefix[2][n_cells][n_ranks] (second half is scratch), remaining[n_cells][n_ranks] (initialised with the GLOBAL ranks).
 *   1. SODM_SELECT_MINMAX     fills lo / hi / nan_flag from the local shard   -> all-reduce MIN(lo), MAX(hi), MAX(nan_flag)
 *   2. repeat ceil(key bits / 11) times (cells settle early and are then skipped):
 *        SODM_SELECT_HISTOGRAM  hist[n_cells][n_ranks][2048] (uint32) of the local shard   -> all-reduce SUM(hist)
 *        SODM_SELECT_PICK       every rank advances prefix / remaining / settled identically
 *   3. SODM_SELECT_DECODE     out[n_ranks][n_cells] as in sodm_draws_order_stats */
#define SODM_SELECT_MINMAX    0
#define SODM_SELECT_HISTOGRAM 1
#define SODM_SELECT_PICK      2
#define SODM_SELECT_DECODE    3
#define SODM_SELECT_BINS      2048
typedef struct sodm_select_state {
    int64_t* lo;
    int64_t* hi;
    int32_t* nan_flag;
    int32_t* settled;
    int64_t* prefix;
    int64_t* remaining;
} sodm_select_state;
int sodm_draws_select_phase(int32_t phase, const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld,
                            int32_t n_ranks, const sodm_select_state* state, uint32_t* hist, double* out, void* stream);

/* ---- optimiser populations on the device ----------------------------------------------------------------------
 * What the reference's optimisers do to the whole population on the host between two evaluations of the objective:
 *   swarm update + clamp          src/pySODM/optimization/pso.py:207-216            -> sodm_pso_move
 *   personal bests, arg min       src/pySODM/optimization/pso.py:229-236            -> sodm_pso_best
 *   proposal density of the KDE move of the default cocktail   src/pySODM/optimization/mcmc.py:26 (emcee KDEMove,
 *   scipy.stats.gaussian_kde.logpdf)                                               -> sodm_kde_logpdf
 * Populations are FP64 structure-of-arrays [n_dim][ld] on the device (member i, dimension d at base[d * ld + i]) -- the
 * layout of `var_values`, so a population is passed to sodm_jump_run / sodm_ode_run without a transpose. */
#define SODM_POPULATION_MAX_DIM 16

typedef struct sodm_pso_args {
    int32_t       n_dim;
    int64_t       n, ld;
    double*       x;           /* (device) [n_dim][ld] positions, updated in place                                  */
    double*       v;           /* (device) [n_dim][ld] velocities, updated in place                                 */
    const double* p;           /* (device) [n_dim][ld] best position of every particle                              */
    const double* g;           /* (device) [n_dim] best position of the swarm                                       */
    const double* lb;          /* (host) [n_dim]                                                                    */
    const double* ub;          /* (host) [n_dim], ub > lb                                                           */
    double        omega, phip, phig;
    uint64_t      seed;        /* r_p, r_g ~ U[0,1): Philox keyed by (seed, member_id0 + i, iteration)              */
    int64_t       iteration;   /* 0 .. 2^32 - 1                                                                     */
    int64_t       member_id0;  /* global index of particle 0 of this shard                                          */
    void*         stream;
} sodm_pso_args;
/* v = omega v + phip r_p (p - x) + phig r_g (g - x);  x = clamp(x + v, lb, ub) */
int sodm_pso_move(const sodm_pso_args* args);

/* p, fp <- x, fx where fx < fp (and feasible[i] != 0 when `feasible` is given); then i* = np.argmin(fp) (first index of the
 * minimum, NaN first):  best[0] = fp[i*], best[1 .. n_dim] = p[:, i*], *best_index = i*.  All pointers are device memory. */
int sodm_pso_best(int32_t n_dim, int64_t n, int64_t ld, const double* x, const double* fx, const uint8_t* feasible,
                  double* p, double* fp, double* best /* [1 + n_dim] */, int64_t* best_index, void* stream);

/* out[j] = log sum_k exp(-|q_j - x_k|^2 / 2) + log_norm for WHITENED points x (kernel centres) and queries q: with
 * L L^T = the kernel covariance, x = L^-1 data, q = L^-1 query and log_norm = -log(n_points) - log|L| - n_dim/2 log(2 pi)
 * this is gaussian_kde(data).logpdf(query). */
int sodm_kde_logpdf(int32_t n_dim, const double* points, int64_t n_points, int64_t ld_points, const double* queries,
                    int64_t n_queries, int64_t ld_queries, double log_norm, double* out /* [n_queries] */, void* stream);

/* Ensemble-MCMC moves of the reference's default cocktail (src/pySODM/optimization/mcmc.py:22-28; emcee's red-blue scheme:
 * the walkers listed in `mine` propose from the walkers listed in `others`, the random split of the iteration):
 *   SODM_MOVE_STRETCH  q = c + z (s - c), z = ((a-1) u + 1)^2 / a, log factor (D-1) log z          param = a  (emcee: 2)
 *   SODM_MOVE_DE       q = s + gamma (c_a - c_b) + eps, eps ~ N(0, sigma^2), log factor 0           param = gamma (<= 0: 2.38/sqrt(2D))
 *   SODM_MOVE_SNOOKER  q = s + gamma (u.z1 - u.z2) u, u = (s - z)/|s - z|, log factor (D-1) log(|q-z|/|s-z|)   param = gamma (emcee: 1.7)
 *   SODM_MOVE_KDE      see below
 * Random numbers: Philox keyed by (seed, walker index, 2 * iteration) for the proposal and (.., 2 * iteration + 1) for the
 * acceptance, so a run does not depend on launch geometry. */
#define SODM_MOVE_STRETCH 0
#define SODM_MOVE_DE      1
#define SODM_MOVE_SNOOKER 2
#define SODM_MOVE_KDE     3   /* q = c_k + L z, z ~ N(0, I): a draw from the Gaussian KDE of `others` whose kernel covariance has
                               * the lower Cholesky factor `chol`; log_factor is NOT written: the move also fills `white` with
                               * L^-1 times (the n_others kernel centres | the n_mine current positions | the n_mine proposals),
                               * the arguments of sodm_kde_logpdf, and the factor is density(s) - density(q) */
#define SODM_WALKERS_ST_NAN 1      /* a proposal scored NaN (emcee: "Probability function returned NaN") */

typedef struct sodm_walker_move_args {
    int32_t        kind;        /* SODM_MOVE_*                                                                       */
    int32_t        n_dim;
    int64_t        n_walkers, ld;
    const double*  x;           /* (device) [n_dim][ld] the ensemble                                                 */
    const int64_t* mine;        /* (device) [n_mine] walkers being moved                                             */
    int64_t        n_mine;
    const int64_t* others;      /* (device) [n_others] the complementary walkers                                     */
    int64_t        n_others;
    double         param, sigma;
    double*        q;           /* (device) [n_dim][ld_q] proposals, column j for walker mine[j]                     */
    int64_t        ld_q;
    double*        log_factor;  /* (device) [n_mine] log of the Hastings factor                                      */
    const double*  chol;        /* (HOST) [n_dim][n_dim] row-major lower Cholesky factor (SODM_MOVE_KDE), else NULL      */
    double*        white;       /* (device) [n_dim][ld_white] whitened coordinates (SODM_MOVE_KDE), else NULL            */
    int64_t        ld_white;    /* >= n_others + 2 n_mine                                                            */
    uint64_t       seed;
    int64_t        iteration;   /* 0 .. 2^31 - 1                                                                     */
    void*          stream;
} sodm_walker_move_args;
int sodm_walkers_propose(const sodm_walker_move_args* args);

/* Clip x into [lb, ub] (host arrays) in place; penalty[i] = -inf where anything was clipped, else 0: the reference's
 * treatment of an out-of-bounds theta (src/pySODM/optimization/objective_functions.py:301-307). */
int sodm_population_clip(int32_t n_dim, int64_t n, int64_t ld, double* x, const double* lb, const double* ub,
                         double* penalty /* (device) [n] */, void* stream);

/* Walker mine[j] moves to q[:, j] (x, lp updated in place, accepted[mine[j]] += 1) when
 * log_factor[j] + new_lp[j] - lp[mine[j]] > log u; a clipped proposal (penalty[j] < 0) or an infinite score never moves;
 * a NaN score sets SODM_WALKERS_ST_NAN. */
int sodm_walkers_accept(int32_t n_dim, int64_t n_walkers, int64_t ld, double* x, double* lp, const int64_t* mine,
                        int64_t n_mine, const double* q, int64_t ld_q, const double* new_lp, const double* penalty /* or NULL */,
                        const double* log_factor, uint64_t seed, int64_t iteration, int64_t* accepted /* [n_walkers] */,
                        int32_t* status /* or NULL */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SODM_H */
