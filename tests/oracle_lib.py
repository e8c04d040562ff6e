"""ctypes binding of the CPU oracle (oracle/liboracle_dde.so).  TEST INFRASTRUCTURE ONLY.

The oracle is the checker for the parity tests; it is never imported by the product package
(delaydiffeq.jl_b200/b200dde).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

ALG_TSIT5, ALG_ROSENBROCK23, ALG_BS3 = 0, 1, 2
(MODEL_BC, MODEL_MACKEY_GLASS, MODEL_STATE_DEP, MODEL_STIFF, MODEL_LOGISTIC, MODEL_1DELAY, MODEL_2DELAYS,
 MODEL_1DELAY_LONG, MODEL_1DELAY_DEP, MODEL_BACKWARDS, MODEL_SIR, MODEL_SIR_DDAE) = range(12)
RC_DEFAULT, RC_SUCCESS, RC_MAXITERS, RC_DTLESSTHANMIN, RC_UNSTABLE, RC_DTNAN, RC_HISTORY_OVERFLOW, RC_QUEUE_OVERFLOW = range(8)
STAT_NACCEPT, STAT_NREJECT, STAT_NF, STAT_NFPITER, STAT_NFPCONVFAIL, STAT_NTRACKED, STAT_MAXNODES, STAT_NSOLVE = range(8)
N_STATS = 8

# (n_state, n_param, n_const_lags, n_dep_lags)
MODEL_DIMS = {
    MODEL_BC: (3, 3, 1, 0), MODEL_MACKEY_GLASS: (1, 2, 1, 0), MODEL_STATE_DEP: (1, 1, 0, 1), MODEL_STIFF: (3, 1, 1, 0),
    MODEL_LOGISTIC: (1, 1, 1, 0), MODEL_1DELAY: (1, 1, 1, 0), MODEL_2DELAYS: (1, 1, 2, 0),
    MODEL_1DELAY_LONG: (1, 1, 1, 0), MODEL_1DELAY_DEP: (1, 1, 0, 1), MODEL_BACKWARDS: (1, 1, 1, 0),
    MODEL_SIR: (3, 1, 1, 0), MODEL_SIR_DDAE: (3, 1, 1, 0),
}


class Config(C.Structure):
    """Mirror of b2d_config (include/b200dde.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("alg", C.c_int32), ("constrained", C.c_int32), ("fp_max_iter", C.c_int32),
        ("fp_kappa", C.c_double), ("fp_div_rule", C.c_int32), ("controller_pow", C.c_int32),
        ("model_id", C.c_int32), ("order_discontinuity_t0", C.c_int32), ("neutral", C.c_int32),
        ("n_state", C.c_int32), ("n_param", C.c_int32), ("n_const_lags", C.c_int32), ("n_dep_lags", C.c_int32),
        ("abstol", C.c_double), ("reltol", C.c_double), ("dt0", C.c_double), ("dtmin", C.c_double),
        ("dtmax", C.c_double), ("maxiters", C.c_int64),
        ("gamma", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double), ("qsteady_min", C.c_double),
        ("qsteady_max", C.c_double), ("qoldinit", C.c_double), ("failfactor", C.c_double),
        ("disc_interp_points", C.c_int32), ("track_theta_mode", C.c_int32),
        ("disc_abstol", C.c_double), ("disc_reltol", C.c_double),
        ("rb23_dT_mode", C.c_int32), ("ring_capacity", C.c_int32), ("device", C.c_int32), ("chunk_traj", C.c_int32),
        ("fp_alg", C.c_int32), ("fp_max_history", C.c_int32), ("fp_aa_start", C.c_int32), ("fp_anderson_reset", C.c_int32),
        ("callback", C.c_int32), ("cb_affect", C.c_int32), ("cb_idx", C.c_int32), ("cb_save_positions", C.c_int32),
        ("cb_par0", C.c_double), ("cb_par1", C.c_double), ("cb_par2", C.c_double), ("cb_par3", C.c_double),
        ("cb_cond_idx", C.c_int32), ("cb_direction", C.c_int32), ("cb_interp_points", C.c_int32), ("cb_reserved", C.c_int32),
    ]


_libs = {}


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


# Which arithmetic layer calls without an explicit `libm` argument use: False = csrc/detmath.h (what the kernel computes with),
# True = glibc pow / log10 / exp2f.  tests/test_oracle_kat.py runs every known-answer test under both.
DEFAULT_LIBM = False


def load(libm=None):
    libm = DEFAULT_LIBM if libm is None else libm
    key = "libm" if libm else "det"
    if key not in _libs:
        name = "liboracle_dde_libm.so" if libm else ("liboracle_dde.so" if _cpu_has_fma() else "liboracle_dde_nofma.so")
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.oracle_config_default.argtypes = [C.POINTER(Config), C.c_int32]
        lib.oracle_solve.restype = C.c_int
        lib.oracle_solve_everystep.restype = C.c_int
        lib.oracle_hardware_threads.restype = C.c_int
        _libs[key] = lib
    return _libs[key]


def default_config(alg=ALG_TSIT5, model=MODEL_BC, libm=None, **kw):
    cfg = Config()
    load(libm).oracle_config_default(C.byref(cfg), alg)
    cfg.model_id = model
    n, npar, ncl, ndl = MODEL_DIMS[model]
    cfg.n_state, cfg.n_param, cfg.n_const_lags, cfg.n_dep_lags = n, npar, ncl, ndl
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def solve(cfg, u0, p, lags, t0, tf, saveat=(), save_start=True, save_end=True, nthreads=1, libm=None, out=None):
    """Ensemble solve.  u0: [n_traj, n_state], p: [n_traj, n_param] (row = trajectory, i.e. the
    transpose view of the column-major ABI matrices).  Returns dict(t, u[n_traj,n_out,n], retcode, stats)."""
    lib = load(libm)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    n_traj, n = u0.shape
    lags = np.ascontiguousarray(lags, dtype=np.float64)
    if lags.ndim == 2:  # per-trajectory constant lags: every trajectory is a problem of its own (src/solve.jl:159)
        parts = [solve(cfg, u0[i:i + 1], p[i:i + 1], lags[i], t0, tf, saveat, save_start, save_end, 1, libm) for i in range(n_traj)]
        return dict(t=parts[0]["t"], u=np.concatenate([q["u"] for q in parts]), retcode=np.concatenate([q["retcode"] for q in parts]),
                    stats=np.concatenate([q["stats"] for q in parts]))
    saveat = np.ascontiguousarray(saveat, dtype=np.float64)
    tdir = 1.0 if tf > t0 else -1.0
    inner = [s for s in saveat if tdir * t0 < tdir * s < tdir * tf]
    n_out = int(save_start) + len(inner) + int(save_end)
    ns = getattr(solve, "n_save_comp", 0) or n
    out_u = out if out is not None else np.empty((n_traj, n_out, ns), dtype=np.float64)
    assert out_u.shape == (n_traj, n_out, ns) and out_u.flags.c_contiguous
    out_t = np.empty(n_out, dtype=np.float64)
    rc = np.zeros(n_traj, dtype=np.int32)
    stats = np.zeros((n_traj, N_STATS), dtype=np.int64)
    r = lib.oracle_solve(C.byref(cfg), C.c_int64(n_traj), C.c_double(t0), C.c_double(tf), _dp(u0), _dp(p), _dp(lags),
                         _dp(saveat), C.c_int32(len(saveat)), C.c_int32(int(save_start)), C.c_int32(int(save_end)),
                         _dp(out_u), _dp(out_t), rc.ctypes.data_as(C.POINTER(C.c_int32)),
                         stats.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int32(nthreads))
    assert r == n_out, (r, n_out)
    return dict(t=out_t, u=out_u, retcode=rc, stats=stats)


class Sol:
    """Single-trajectory save_everystep solution with dense output (sol(t))."""

    def __init__(self, alg, t, u, k, tracked, retcode, stats):
        self.alg, self.t, self.u, self.k, self.tracked, self.retcode, self.stats = alg, t, u, k, tracked, retcode, stats

    def __len__(self):
        return len(self.t)

    def __call__(self, tq, libm=None):
        # upstream ode_interpolation(:left): i+ = first index >= 2 (1-based) with ts[i] >= tq
        ts = self.t
        tdir = 1.0 if ts[-1] >= ts[0] else -1.0
        ip = 1
        while ip < len(ts) - 1 and tdir * ts[ip] < tdir * tq:
            ip += 1
        im = ip - 1
        dt = ts[ip] - ts[im]
        th = 1.0 if dt == 0 else (tq - ts[im]) / dt
        n = self.u.shape[1]
        out = np.empty(n)
        y0 = np.ascontiguousarray(self.u[im])
        y1 = np.ascontiguousarray(self.u[ip])
        kk = np.ascontiguousarray(self.k[ip])
        load(libm).oracle_interpolate(C.c_int32(self.alg), C.c_int32(n), C.c_double(th), C.c_double(dt), _dp(y0), _dp(y1),
                                      _dp(kk), _dp(out))
        return out


def solve_everystep(cfg, u0, p, lags, t0, tf, max_steps=200000, max_tracked=4096, libm=None):
    lib = load(libm)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).ravel()
    p = np.ascontiguousarray(p, dtype=np.float64).ravel()
    lags = np.ascontiguousarray(lags, dtype=np.float64)
    n = len(u0)
    nk = 2 if cfg.alg in (ALG_ROSENBROCK23, ALG_BS3) else 7
    out_t = np.empty(max_steps)
    out_u = np.empty((max_steps, n))
    hist_k = np.zeros((max_steps, nk, n))
    n_steps = C.c_int32(0)
    tr_t = np.empty(max_tracked)
    tr_o = np.empty(max_tracked, dtype=np.int32)
    n_tr = C.c_int32(0)
    rc = C.c_int32(0)
    stats = np.zeros(N_STATS, dtype=np.int64)
    r = lib.oracle_solve_everystep(C.byref(cfg), C.c_double(t0), C.c_double(tf), _dp(u0), _dp(p), _dp(lags),
                                   C.c_int32(max_steps), _dp(out_t), _dp(out_u), C.byref(n_steps), C.c_int32(max_tracked),
                                   _dp(tr_t), tr_o.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n_tr), C.byref(rc),
                                   stats.ctypes.data_as(C.POINTER(C.c_int64)), _dp(hist_k))
    assert r == 0
    ns = min(n_steps.value, max_steps)
    nt = min(n_tr.value, max_tracked)
    tracked = [(float(tr_t[i]), int(tr_o[i])) for i in range(nt)]
    return Sol(cfg.alg, out_t[:ns].copy(), out_u[:ns].copy(), hist_k[:ns].copy(), tracked, rc.value, stats)


def set_options(abstol=None, reltol=None, save_idxs=None, libm=None):
    """Per-component tolerances / save_idxs for the solves that follow (None: back to the defaults)."""
    a = np.ascontiguousarray(abstol if abstol is not None else [], dtype=np.float64)
    r = np.ascontiguousarray(reltol if reltol is not None else [], dtype=np.float64)
    i = np.ascontiguousarray(save_idxs if save_idxs is not None else [], dtype=np.int32)
    load(libm).oracle_set_options(_dp(a), _dp(r), C.c_int32(len(a)), i.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(len(i)))
    solve.n_save_comp = len(i)


def set_stops(tstops=None, discs=None, libm=None):
    """solve(...; tstops, d_discontinuities) for the solves that follow (None: clear).  discs: (t, order) pairs."""
    t = np.ascontiguousarray(tstops if tstops is not None else [], dtype=np.float64)
    dt = np.ascontiguousarray([d[0] for d in (discs or [])], dtype=np.float64)
    do = np.ascontiguousarray([d[1] for d in (discs or [])], dtype=np.int32)
    load(libm).oracle_set_stops(_dp(t), C.c_int32(len(t)), _dp(dt), do.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(len(dt)))


def hardware_threads():
    return load().oracle_hardware_threads()


def eval_model(model, u, p, hv, dhv, t, lags, n_dep_lags=0, has_jac=False, libm=None):
    """f (and jac / tgrad / dependent lags) of a registry model at the points u, p with the history handle returning hv / dhv
    ([n_pts, slots, n]): the CPU counterpart of b2d_eval_model."""
    u = np.ascontiguousarray(u, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64)
    hv = np.ascontiguousarray(hv, dtype=np.float64); dhv = np.ascontiguousarray(dhv, dtype=np.float64)
    n_pts, n = u.shape
    lg = np.ascontiguousarray(list(lags) + [0.0] * 4, dtype=np.float64)
    f = np.zeros((n_pts, n)); jac = np.zeros((n_pts, n, n)); tg = np.zeros((n_pts, n)); lag = np.zeros((n_pts, max(n_dep_lags, 1)))
    lib = load(libm)
    lib.oracle_eval_model.restype = C.c_int
    rc = lib.oracle_eval_model(C.c_int32(model), C.c_int32(n_pts), _dp(u), _dp(p), _dp(hv), _dp(dhv), C.c_double(t), _dp(lg),
                               _dp(f), _dp(jac), _dp(tg), _dp(lag))
    assert rc == 0
    return f, (jac if has_jac else None), (tg if has_jac else None), lag[:, :n_dep_lags]


def history_probe(cfg, u0, p, lags, t0, tf, libm=None):
    """The lookup regimes of HistoryFunction as test/interface/history_function.jl walks them (oracle_history_probe).  Returns
    (records, flags): records = list of (deriv, isout, got[N], expected[N]) in the order user history (deriv 0, 1), constant
    extrapolation (0, 1), integrator interpolation (0, 1), solution interpolation (0, 1), integrator extrapolation (0, 1);
    flags = (solution still ends at t0 after perform_step!, integrator.t == sol.t[end] after loopfooter!)."""
    lib = load(libm)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    lags = np.ascontiguousarray(lags, dtype=np.float64)
    n = len(u0)
    rec = np.zeros(10 * (2 + 2 * n) + 2, dtype=np.float64)
    r = lib.oracle_history_probe(C.byref(cfg), C.c_double(t0), C.c_double(tf), _dp(u0), _dp(p), _dp(lags), _dp(rec), C.c_int32(len(rec)))
    assert r == len(rec), r
    out, flags, k = [], [], 0
    for i in range(10):
        if i == 6:
            flags.append(rec[k] == 1.0); k += 1
            flags.append(rec[k] == 1.0); k += 1
        out.append((int(rec[k]), rec[k + 1] == 1.0, rec[k + 2:k + 2 + n].copy(), rec[k + 2 + n:k + 2 + 2 * n].copy()))
        k += 2 + 2 * n
    return out, tuple(flags)
