/*
 * b200dde.h — C ABI of libb200dde.so, the B200-native batched MethodOfSteps ensemble solver.
 *
 * This is the drop-in boundary for the ONE hot path of SciML/DelayDiffEq.jl that this
 * repository replaces: solving an EnsembleProblem of DDEProblems with
 * MethodOfSteps(Tsit5()) / MethodOfSteps(Rosenbrock23()).  The reference has no FFI of its
 * own (pure Julia); each entry point below names the reference interface it replaces
 * (paths relative to the reference tree, DelayDiffEq.jl v5.69.0).  The Julia-side binding
 * a maintainer would add is shown in INTEGRATION.md and delaydiffeq.jl_b200/julia/B200DDE.jl.
 *
 * Conventions
 *   - all pointers are HOST pointers owned by the caller unless the function name ends in
 *     `_device`; the library never frees caller memory;
 *   - matrices are column-major exactly like Julia `Matrix{Float64}` / `Array{Float64,3}`;
 *   - return value 0 = ok, <0 = library error (text via b2d_last_error(), thread-local);
 *     per-trajectory failures are NOT errors, they are reported in retcode[];
 *   - there is no CPU fallback: if no sm_100 device is usable every solve call fails.
 */
#ifndef B200DDE_H
#define B200DDE_H

#ifdef __CUDACC_RTC__ /* runtime compilation of user models: no host headers */
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B2D_VERSION 100 /* 0.1.0 */

/* ---- algorithm selection: replaces MethodOfSteps(alg; constrained, fpsolve)  (src/algorithms.jl:4-36) */
#define B2D_ALG_TSIT5 0
#define B2D_ALG_ROSENBROCK23 1
#define B2D_ALG_BS3 2 /* the algorithm of the reference's analytic-error tests (test/interface/unconstrained.jl) */

/* ---- ahead-of-time compiled RHS registry (model_id).  A Julia closure cannot cross a C ABI;
 *      f(du,u,h,p,t), h(p,t) and the lag functions are device functors selected by id
 *      (or supplied as CUDA C++ source through b2d_create_user_model, compiled with NVRTC). */
#define B2D_MODEL_BC 0           /* README.md:23-40 bc_model, p = [p0,q0,v0]                  */
#define B2D_MODEL_MACKEY_GLASS 1 /* u' = beta u(t-tau)/(1+u(t-tau)^10) - 0.1 u, p=[beta,h0]   */
#define B2D_MODEL_STATE_DEP 2    /* test/regression/allocations.jl:185-202, lag 0.5+0.1|u|    */
#define B2D_MODEL_STIFF 3        /* SURVEY §8d C5 stiff delayed kinetics, p=[lambda]          */
#define B2D_MODEL_LOGISTIC 4     /* test/interface/parameters.jl:7-22 delayed logistic        */
#define B2D_MODEL_1DELAY 5       /* DDEProblemLibrary prob_dde_constant_1delay_ip             */
#define B2D_MODEL_2DELAYS 6      /* DDEProblemLibrary prob_dde_constant_2delays_ip            */
#define B2D_MODEL_1DELAY_LONG 7  /* DDEProblemLibrary prob_dde_constant_1delay_long_ip        */
#define B2D_MODEL_1DELAY_DEP 8   /* test/interface/dependent_delays.jl:18 lag (u,p,t)->1      */
#define B2D_MODEL_BACKWARDS 9    /* test/interface/backwards.jl:4-12  u' = h(t+1), reverse t  */
#define B2D_MODEL_SIR 10         /* test/interface/mass_matrix.jl:7-19 SIR, delayed recovery  */
#define B2D_MODEL_SIR_DDAE 11    /* test/interface/mass_matrix.jl:33-58 same as DDAE, M = diag(1,1,0) */
#define B2D_MODEL_COUNT 12

/* ---- per-trajectory return codes: SciMLBase.ReturnCode as tested in test/integrators/retcode.jl:13-19 */
#define B2D_RC_DEFAULT 0
#define B2D_RC_SUCCESS 1
#define B2D_RC_MAXITERS 2
#define B2D_RC_DTLESSTHANMIN 3
#define B2D_RC_UNSTABLE 4
#define B2D_RC_DTNAN 5
#define B2D_RC_HISTORY_OVERFLOW 6 /* device history ring too small (no reference analogue) */
#define B2D_RC_QUEUE_OVERFLOW 7   /* device discontinuity queue too small                   */
#define B2D_RC_TERMINATED 8       /* terminate!(integrator) from a callback: ReturnCode.Terminated */
#define B2D_RC_INEXACT_DIVISION 9 /* the lean kernel met a division operand outside the range of its branch-free sequence
                                     (denormal / huge values, zero or non-finite divisor): b2d_solve solves such a trajectory
                                     again in the general kernel, b2d_solve_device reports it (no reference analogue) */

/* ---- stats rows: SciMLBase.DEStats fields the path updates
 *      (src/solve.jl:242, src/fpsolve/fpsolve.jl:8-11, src/integrators/interface.jl:535) */
#define B2D_STAT_NACCEPT 0
#define B2D_STAT_NREJECT 1
#define B2D_STAT_NF 2
#define B2D_STAT_NFPITER 3
#define B2D_STAT_NFPCONVFAIL 4
#define B2D_STAT_NTRACKED 5 /* length(integrator.tracked_discontinuities)            */
#define B2D_STAT_MAXNODES 6 /* max live history-ring nodes (device) / 0 (oracle)      */
#define B2D_STAT_NSOLVE 7   /* linear solves (Rosenbrock23)                           */
#define B2D_N_STATS 8

/* ---- solver options: the keyword arguments of init/solve that the path reads
 *      (src/solve.jl:44-101) plus MethodOfSteps' own fields.  Zero/negative sentinels mean
 *      "reference default"; b2d_config_default() fills every field explicitly. */
typedef struct b2d_config {
    int32_t struct_size; /* = sizeof(b2d_config); checked by b2d_create                     */
    int32_t alg;         /* B2D_ALG_*                                                       */
    int32_t constrained; /* MethodOfSteps(...; constrained)  src/solve.jl:172-182            */
    int32_t fp_max_iter; /* NLFunctional max_iter (10)        src/fpsolve/utils.jl:35         */
    double fp_kappa;     /* NLFunctional kappa (1/100)                                      */
    int32_t fp_div_rule; /* nlsolve! divergence test: 1 (default) theta>2; 2: + precision-limit exit; 0: Hairer estimate */
    int32_t controller_pow; /* 1: FastPower.fastpower restatement (default), 0: exact pow  */

    int32_t model_id;
    int32_t order_discontinuity_t0; /* -1: model default (DDEProblem field, src/solve.jl:107) */
    int32_t neutral;                /* src/integrators/utils.jl:245-246                       */
    int32_t n_state, n_param;       /* validated against the model                            */
    int32_t n_const_lags, n_dep_lags;

    double abstol, reltol; /* src/utils.jl:91-133 defaults 1e-6 / 1e-3                       */
    double dt0;            /* initial dt; 0 => ode_determine_initdt (interface.jl:514-538)  */
    double dtmin;          /* <0 => prob2dtmin = max(eps(Float64), eps(t0))                 */
    double dtmax;          /* 0 => tf - t0 (src/solve.jl:60)                                 */
    int64_t maxiters;      /* 1000000 (src/solve.jl:76)                                      */
    double gamma, qmin, qmax, qsteady_min, qsteady_max, qoldinit, failfactor; /* solve.jl:63-73 */

    int32_t disc_interp_points; /* discontinuity_interp_points = 10  (src/solve.jl:97)       */
    int32_t track_theta_mode;   /* 0: reference-literal Theta=(s-tprev)/dt in track.jl, 1: (s-t)/dt */
    double disc_abstol;         /* 1e-12 (src/solve.jl:98)                                   */
    double disc_reltol;         /* 0     (src/solve.jl:99)                                   */

    int32_t rb23_dT_mode;  /* Rosenbrock23 k3 time-derivative factor: 0 dt*dT (upstream), 1 gamma*dT */
    int32_t ring_capacity; /* history nodes kept per trajectory on device; 0 => auto        */
    int32_t device;        /* CUDA device ordinal                                           */
    int32_t chunk_traj;    /* trajectories per wave for host-buffer solves; 0 => auto       */
    /* fixed-point solver of MethodOfSteps(alg; fpsolve = ...), src/algorithms.jl:4-36, src/fpsolve/utils.jl:2-79 */
    int32_t fp_alg;           /* B2D_FP_FUNCTIONAL (NLFunctional, default) or B2D_FP_ANDERSON (NLAnderson,
                                 src/fpsolve/functional.jl:15-75)                                               */
    int32_t fp_max_history;   /* NLAnderson max_history (10); effective size min(max_history, max_iter, n_state) */
    int32_t fp_aa_start;      /* NLAnderson aa_start (1)                                                       */
    int32_t fp_anderson_reset; /* 0 (default): the Anderson history survives from one fixed-point solve to the next,
                                 as on the reference's FPSolver path (no initialize! method resets it in-tree);
                                 1: history = 0 at the start of every fixed-point solve                        */
    /* solve(...; callback = DiscreteCallback(condition, affect!; save_positions)) — src/solve.jl:632-681,
     * src/integrators/interface.jl:585-632.  A Julia closure cannot cross the ABI: condition and affect! are either device
     * functors of the model (cb_condition / cb_affect, B2D_CB_MODEL) or the parametric preset callback built into the
     * kernels: condition `t == cb_par[0]` (B2D_CB_AT_TIME) or `t - t0 is a multiple of cb_par[0]` (B2D_CB_PERIODIC; the
     * DiffEqCallbacks PresetTimeCallback / PeriodicCallback shape — the host adds the times to tstops), affect! one of
     * B2D_CBA_*.  An affect! marks u as modified: handle_callback_modifiers! then re-evaluates the FSAL slope and pushes an
     * order-0 discontinuity at t, which propagates through the lags like any other. */
    int32_t callback;          /* B2D_CB_NONE (default) / B2D_CB_MODEL / B2D_CB_AT_TIME / B2D_CB_PERIODIC / B2D_CB_CONT_* (below) */
    int32_t cb_affect;         /* B2D_CBA_NONE / _TERMINATE / _ADD (u[cb_idx] += cb_par[1]) / _NEGATE (u = -u) */
    int32_t cb_idx;            /* component of B2D_CBA_ADD (0-based) */
    int32_t cb_save_positions; /* bit 0: save before affect!, bit 1: after (DiscreteCallback default (true, true) = 3).  The
                                  forced saves always enter the history; the points they add to sol.t / sol.u are returned
                                  by the save_everystep entry points and left out of a fixed-shape saveat block */
    double cb_par0, cb_par1, cb_par2, cb_par3; /* the parameters cb_par[0..3] of the preset callback / of the model's functors */
    /* solve(...; callback = ContinuousCallback(condition, affect!, affect_neg!; interp_points)): callback = B2D_CB_CONT_*.
     * The condition is a real function of (u, t) whose sign change inside a step is the event: the model's functor
     * cb_value(u, t, p, cb_par) (B2D_CB_CONT_MODEL), `t - cb_par[0]` (B2D_CB_CONT_TIME) or `u[cb_cond_idx] - cb_par[0]`
     * (B2D_CB_CONT_STATE).  The event time is located on the step's dense output (upstream find_callback_time: sign test
     * at interp_points samples, ITP bracketing, LeftRootFind), the step is cut there (change_t_via_interpolation!,
     * src/integrators/interface.jl:566-572, slopes recomputed by reeval_internals_due_to_modification!, :597-632), affect!
     * runs, and handle_callback_modifiers! (:585-594) pushes an order-0 discontinuity.  Tsit5 and BS3. */
    int32_t cb_cond_idx;      /* component of B2D_CB_CONT_STATE (0-based) */
    int32_t cb_direction;     /* 0: both crossings (affect_neg! = affect!, the default); +1: upcrossings only
                                 (affect_neg! = nothing); -1: downcrossings only (affect! = nothing) */
    int32_t cb_interp_points; /* 10 */
    int32_t cb_reserved;
} b2d_config;

#define B2D_CB_NONE 0
#define B2D_CB_MODEL 1
#define B2D_CB_AT_TIME 2
#define B2D_CB_PERIODIC 3
#define B2D_CB_CONT_MODEL 4 /* ContinuousCallback kinds: >= B2D_CB_CONT_MODEL */
#define B2D_CB_CONT_TIME 5
#define B2D_CB_CONT_STATE 6
#define B2D_CBA_NONE 0
#define B2D_CBA_TERMINATE 1
#define B2D_CBA_ADD 2
#define B2D_CBA_NEGATE 3

#define B2D_FP_FUNCTIONAL 0
#define B2D_FP_ANDERSON 1

#ifndef __CUDACC_RTC__ /* the entry points are host functions: hidden from run-time compiled device code */
typedef struct b2d_solver *b2d_handle;

/* Fill *cfg with the reference defaults for `alg` (src/solve.jl:44-101, OrdinaryDiffEqCore
 * gamma/qmin/qmax/qsteady defaults, NLFunctional defaults). */
void b2d_config_default(b2d_config *cfg, int32_t alg);

/* Static facts about a registry model (n_state, n_param, n_const_lags, n_dep_lags,
 * default order_discontinuity_t0); returns <0 for an unknown id. */
int b2d_model_info(int32_t model_id, int32_t *n_state, int32_t *n_param, int32_t *n_const_lags,
                   int32_t *n_dep_lags, int32_t *order_discontinuity_t0);

/* Diagonal of a registry model's mass matrix (DDEFunction(f; mass_matrix), src/functionwrapper.jl:31,60) into
 * diag[0..n_state).  Returns 0 for the identity, 1 for a non-singular and 2 for a singular matrix (the problem is
 * a DDAE: `isdae`, src/solve.jl:155-156; hosts then default `neutral` to true as DDEProblem does), <0 for an
 * unknown id.  Only Rosenbrock23 honours a mass matrix (test/interface/mass_matrix.jl:65-74). */
int b2d_model_mass_matrix(int32_t model_id, double *diag, int32_t n);

/* Create / destroy a solver.  Replaces SciMLBase.__init's allocation of integrators, caches,
 * heaps and options (src/solve.jl:38-586) — done once per ensemble instead of per trajectory. */
int b2d_create(b2d_handle *out, const b2d_config *cfg);
int b2d_destroy(b2d_handle h);

/* Per-component tolerances (abstol / reltol given as arrays, src/utils.jl:91-133) for the solves that follow;
 * n = n_state, or n = 0 to go back to the scalars of the config. */
int b2d_set_tolerances(b2d_handle h, const double *abstol, const double *reltol, int32_t n);

/* save_idxs (src/solve.jl:47,217-218): only the listed components (0-based, distinct, in this order) are written by b2d_solve /
 * b2d_solve_multi / b2d_solve_device, i.e. out_u becomes [n x n_out x n_traj]; n = 0 saves every component again.
 * The save_everystep entry points always return the full state. */
int b2d_set_save_idxs(b2d_handle h, const int32_t *idxs, int32_t n);

/* tstops / d_discontinuities keywords of solve (src/solve.jl:45-46,258-270), shared by every trajectory, for the
 * solves that follow.  Points outside (t0, tf] are ignored as upstream initialize_tstops does; every discontinuity time
 * is also a stop.  A discontinuity of order o propagates through the lags with order o + 1 (o for neutral problems)
 * while that does not exceed the order of the method + 1 (src/integrators/utils.jl:244-283).  n = 0 clears. */
int b2d_set_tstops(b2d_handle h, const double *tstops, int32_t n);

/* Per-trajectory constant lags: the reference ensemble solves a full problem per trajectory (src/solve.jl:159 unpacks
 * constant_lags of whatever prob_func returned), so a sweep over the delay itself is an ordinary ensemble.  lags is a
 * host array [n_traj][n_const_lags] (row = trajectory), copied; the host-buffer solves that follow (b2d_solve,
 * b2d_solve_multi — give every handle the whole array —, b2d_solve_dense / _everystep) must have n_traj trajectories and
 * use row i for trajectory i in place of their const_lags argument (which stays the nominal problem's).  Solved by the
 * general kernels.  n_traj = 0 clears.  Not with MethodOfSteps(...; constrained = true). */
int b2d_set_trajectory_lags(b2d_handle h, const double *lags, int64_t n_traj);
int b2d_set_d_discontinuities(b2d_handle h, const double *t, const int32_t *order, int32_t n);

/*
 * Solve n_traj independent trajectories with a shared saveat grid.
 * Replaces, for the whole ensemble, SciMLBase.__solve(prob::AbstractDDEProblem, alg::MethodOfSteps)
 * (src/solve.jl:1-9: __init + solve! src/solve.jl:588-630) called once per trajectory by
 * SciMLBase.solve_batch(EnsembleThreads) (upstream, not in tree).
 *
 *   u0         [n_state x n_traj]            initial states (prob.u0 per trajectory)
 *   p          [n_param x n_traj]            parameters     (prob.p  per trajectory)
 *   const_lags [n_const_lags]                prob.constant_lags (shared)
 *   saveat     [n_save]                      strictly inside (t0,tf), ascending in tdir
 *   save_start / save_end                    src/solve.jl:50-52,326-329
 *   out_t      [n_out]                       n_out = save_start + n_save + save_end
 *   out_u      [n_state x n_out x n_traj]    sol.u ; NaN where a failed trajectory never got
 *   retcode    [n_traj]                      B2D_RC_*
 *   stats      [B2D_N_STATS x n_traj]        may be NULL
 */
int b2d_solve(b2d_handle h, int64_t n_traj, double t0, double tf, const double *u0,
              const double *p, const double *const_lags, const double *saveat, int32_t n_save,
              int32_t save_start, int32_t save_end, double *out_u, double *out_t,
              int32_t *retcode, int64_t *stats);

/* The same solve sharded over several handles of ONE process (typically one handle per GPU of the box): contiguous
 * ranges of trajectories, one host thread per handle, every GPU copies its shard directly into the caller's arrays —
 * trajectories are independent, so there is no exchange and no gather.  This is the multi-GPU entry point
 * for a single-process host such as Julia. */
int b2d_solve_multi(b2d_handle *handles, int32_t n_handles, int64_t n_traj, double t0, double tf, const double *u0,
                    const double *p, const double *const_lags, const double *saveat, int32_t n_save,
                    int32_t save_start, int32_t save_end, double *out_u, double *out_t, int32_t *retcode,
                    int64_t *stats);

/* Same solve with every pointer in DEVICE memory of cfg.device (inputs already resident,
 * outputs left resident) on the given CUDA stream (cudaStream_t passed as void*; NULL =
 * default stream).  Used by multi-process hosts (one rank per GPU) that gather out_u with
 * NCCL, and for kernel-only timing.
 * Synchronisation contract: the KERNEL is asynchronous — the call returns after enqueueing it
 * on `stream` and the outputs are valid once work enqueued on `stream` has completed.  The
 * call itself blocks only for its set-up: the first solve of a (time span, lags, options)
 * combination runs a pilot launch to size the ring and waits for it, and a saveat grid that
 * differs from the previous call's is uploaded synchronously; repeated solves with the same
 * arguments enqueue without any host synchronisation.  A handle owns ONE ring: solves on
 * one handle execute one after the other — a device solve issued on a second stream, and
 * any b2d_solve / b2d_solve_dense that follows, is ordered after the launch in flight by
 * an event, never run concurrently with it.  Trajectories that overflow the ring are
 * reported (B2D_RC_HISTORY_OVERFLOW), not re-solved; b2d_solve re-solves them. */
int b2d_solve_device(b2d_handle h, int64_t n_traj, double t0, double tf, const double *d_u0,
                     const double *d_p, const double *const_lags_host, const double *saveat_host,
                     int32_t n_save, int32_t save_start, int32_t save_end, double *d_out_u,
                     int32_t *d_retcode, int64_t *d_stats, void *stream);

/* The one collective of the path (the reference has none — a Julia process owns all its trajectories): the
 * final gather of the saveat blocks of a MULTI-PROCESS host (one rank per GPU) to the root GPU over NCCL / NVLink.
 *   b2d_comm_unique_id   rank 0 creates the 128-byte ncclUniqueId; the host distributes it (any channel)
 *   b2d_comm_init        every rank joins with its handle (ncclCommInitRank on the handle's device)
 *   b2d_gather_saveat    counts[r] = doubles of rank r's block; the root receives every block at its offset of d_root
 *                        ([n_state x n_out x n_traj] in rank order = trajectory order) with one grouped ncclRecv per peer,
 *                        the others ncclSend theirs; asynchronous on `stream`
 * NCCL is loaded at run time (dlopen libnccl.so.2): inside a process that already uses it, e.g. through torch.distributed,
 * that same copy is used.  A single-process host (b2d_solve_multi) needs no gather: every GPU copies into the caller's array. */
int b2d_comm_unique_id(uint8_t *id128);
int b2d_comm_init(b2d_handle h, int32_t n_ranks, int32_t rank, const uint8_t *id128);
int b2d_gather_saveat(b2d_handle h, const double *d_local, const int64_t *counts, int32_t root, double *d_root, void *stream);

/*
 * save_everystep solve (saveat empty: the default `solve(prob, alg)` of README.md:23-40).
 * Every accepted step of every trajectory is returned, as sol.t / sol.u of the reference:
 *   out_t     [max_steps x n_traj]             sol.t  (first entry t0)
 *   out_u     [n_state x max_steps x n_traj]   sol.u
 *   n_steps   [n_traj]                         length(sol) (capped at max_steps)
 *   tracked_t / tracked_order [max_tracked x n_traj], n_tracked [n_traj]
 *             integrator.tracked_discontinuities (src/integrators/utils.jl:280); may be NULL
 */
int b2d_solve_everystep(b2d_handle h, int64_t n_traj, double t0, double tf, const double *u0,
                        const double *p, const double *const_lags, int32_t max_steps,
                        double *out_t, double *out_u, int32_t *n_steps, int32_t max_tracked,
                        double *tracked_t, int32_t *tracked_order, int32_t *n_tracked,
                        int32_t *retcode, int64_t *stats);

/* save_everystep solve that also returns the dense output: out_k [n_k x n_state x max_steps x n_traj] holds, for every
 * saved time, the slopes of the step ending there (n_k = 7 for Tsit5, 2 for BS3 / Rosenbrock23) — `sol.k` of the
 * reference (src/solve.jl:246-255), i.e. what `sol(t)` interpolates with between sol.t[i-1] and sol.t[i]. */
int b2d_solve_dense(b2d_handle h, int64_t n_traj, double t0, double tf, const double *u0, const double *p,
                    const double *const_lags, int32_t max_steps, double *out_t, double *out_u, double *out_k,
                    int32_t *n_steps, int32_t max_tracked, double *tracked_t, int32_t *tracked_order,
                    int32_t *n_tracked, int32_t *retcode, int64_t *stats);

/*
 * User-supplied models: `model_src` is CUDA C++ source defining a struct named `struct_name` with the
 * interface of csrc/models.cuh (N, NP, NCL, NDL, NH, ORDER0, HAS_JAC, hist_comp, f, h_pre, dep_lag[, jac, tgrad]).
 * It is compiled with NVRTC against the very kernel template the registry models use (same arithmetic rules:
 * -fmad=false) and behaves like a registry model afterwards; cfg->model_id is ignored.
 * This is how an arbitrary Julia f(du,u,h,p,t) / h(p,t) / lag(u,p,t) gets onto the device: the Julia shim passes
 * its CUDA C++ transcription as a string.
 * b2d_compile_user_model only compiles (no GPU needed) and returns the NVRTC log: 0 = ok.
 */
int b2d_create_user_model(b2d_handle *out, const b2d_config *cfg, const char *model_src, const char *struct_name);
int b2d_compile_user_model(const char *model_src, const char *struct_name, int32_t alg, char *log, int32_t log_cap);

/* b2d_destroy parks a handle's launch slots (streams, ring, device and pinned staging buffers) per device and the next
 * b2d_create on that device takes them over, so a host that creates and destroys a handle around every solve pays the
 * allocations once per process.  b2d_release_cached frees everything that is parked. */
int b2d_release_cached(void);

/* Pinned host allocations so that host-buffer solves copy at PCIe speed. */
void *b2d_host_alloc(uint64_t bytes);
void b2d_host_free(void *ptr);

/* CUDA-event time (ms) of the kernels of the last solve on this handle, the number of kernel
 * launches it made, and the ring capacity it used. */
int b2d_last_timing(b2d_handle h, double *kernel_ms, int32_t *n_launches, int32_t *ring_capacity);

/* Diagnostic: measured FP64 FMA peak (TFLOP/s) of `device`, the second roofline denominator of this path
 * (MEASURED_PEAKS.json only carries HBM and bf16). */
int b2d_measure_fp64_tflops(int32_t device, double *tflops);

/* Diagnostic: evaluate the device functors of a registry model alone — f, and where the model has them the Jacobian
 * (row-major n x n), the time gradient and the dependent-lag functions — at n_pts points, with the history handle returning
 * the caller's values: hv / dhv [n_state x slots x n_pts] are the history state and its time derivative the k-th lookup of
 * the model (constant lags first, then dependent lags; slots = max(1, n_const_lags + n_dep_lags)) shall see.  This is how the
 * tests compare csrc/models.cuh with the independent CPU restatement oracle/oracle_models.h, bit for bit, without a solve
 * in between (the transcription of README.md:23-40 and of the test problems of test/interface into device code). */
int b2d_eval_model(int32_t model_id, int32_t device, int32_t n_pts, const double *u, const double *p, const double *hv,
                   const double *dhv, double t, const double *lags, double *out_f, double *out_jac, double *out_tgrad,
                   double *out_lag);

/* Diagnostic: the branch-free division of the kernels (csrc/detmath.h dm_div_try — the compiler's own division sequence
 * with its range test returned as a flag) against the `/` operator on n_pairs generated operand pairs of one class
 * (0 random bit patterns, 1 solver magnitudes, 2 quotients next to a rounding tie, 3 powers of two / zeros / extremes).
 * mismatches counts pairs where the flag was true and the quotient differs from `a / b` in any bit (must be 0); flagged
 * counts pairs the sequence hands back to the ordinary operator. */
int b2d_check_division(int32_t device, int32_t operand_class, uint64_t n_pairs, uint64_t seed, uint64_t *mismatches,
                       uint64_t *flagged);

const char *b2d_last_error(void);
int b2d_version(void);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* B200DDE_H */
