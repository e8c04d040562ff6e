"""Host-side mirror of the reference's public API for the MethodOfSteps ensemble path.

Same names, argument meaning and error behaviour as the Julia API the reference exposes
(DDEProblem / MethodOfSteps / EnsembleProblem / solve with abstol, reltol, saveat ...), executed
through the C ABI of libb200dde.so.  File:line citations point into /root/reference.

The one structural difference a C ABI forces: `f`, `h` and the dependent-lag functions are not
host closures but device functors from the model registry (include/b200dde.h B2D_MODEL_*), named
by string, e.g. DDEProblem("bc_model", u0, None, tspan, p; constant_lags=[1.0]).
"""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _lib
from ._lib import B200DDEError, Config, MODELS, RETCODES, STAT_NAMES, N_STATS


# ----------------------------------------------------------------------------------------------
# algorithms (src/algorithms.jl:4-36)
# ----------------------------------------------------------------------------------------------
class Tsit5:
    alg_id = _lib.ALG_TSIT5


class Rosenbrock23:
    alg_id = _lib.ALG_ROSENBROCK23


class BS3:
    alg_id = _lib.ALG_BS3


class NLFunctional:
    """Fixed-point solver options (upstream OrdinaryDiffEqNonlinearSolve.NLFunctional; used at src/fpsolve/utils.jl:2-79)."""

    def __init__(self, kappa=1.0 / 100.0, max_iter=10):
        self.kappa, self.max_iter = float(kappa), int(max_iter)


class NLAnderson:
    """Anderson-accelerated fixed-point solver (upstream OrdinaryDiffEqNonlinearSolve.NLAnderson; driven by
    src/fpsolve/functional.jl:15-75).  droptol is not supported (upstream default: nothing)."""

    def __init__(self, kappa=1.0 / 100.0, max_iter=10, max_history=10, aa_start=1, droptol=None):
        if droptol is not None:
            raise B200DDEError("NLAnderson(droptol=...) is not implemented on the B200 path")
        self.kappa, self.max_iter = float(kappa), int(max_iter)
        self.max_history, self.aa_start = int(max_history), int(aa_start)


class Discontinuity:
    """Discontinuity(t, order) — src/discontinuity_type.jl:7-38."""

    def __init__(self, t, order):
        self.t, self.order = float(t), int(order)

    def __eq__(self, other):
        return (self.t, self.order) == (other.t, other.order) if isinstance(other, Discontinuity) else self.t == other

    def __repr__(self):
        return f"Discontinuity({self.t}, {self.order})"


class DiscreteCallback:
    """DiscreteCallback(condition, affect!; save_positions = (true, true)) of the reference's solve(...; callback = cb)
    (src/solve.jl:82,632-681; src/integrators/interface.jl:585-632).  A host closure cannot cross the C ABI: `condition`
    is "model" (the device functors cb_condition / cb_affect of the model, e.g. of a UserModel), ("at", t) — the
    PresetTimeCallback shape, `t == tc` — or ("every", period) — the PeriodicCallback shape; `affect` is "terminate",
    ("add", component, value), "negate" or None.  The times of the last two are added to tstops as DiffEqCallbacks does."""

    def __init__(self, condition, affect=None, save_positions=(True, True), par=()):
        self.condition, self.affect = condition, affect
        self.save_positions = (bool(save_positions[0]), bool(save_positions[1]))
        self.par = tuple(float(x) for x in par)

    def tstops(self, tspan):
        t0, tf = tspan
        if self.condition == "model":
            return []
        kind, x = self.condition
        if kind == "at":
            return [float(x)]
        n = int(np.floor(abs(tf - t0) / float(x) + 1e-12))
        return [t0 + np.sign(tf - t0) * k * float(x) for k in range(1, n + 1)]

    def apply(self, cfg):
        par = [0.0, 0.0, 0.0, 0.0]
        if self.condition == "model":
            cfg.callback = 1
            par[:len(self.par)] = self.par
        else:
            kind, x = self.condition
            cfg.callback = {"at": 2, "every": 3}[kind]
            par[0] = float(x)
            if self.affect is None:
                cfg.cb_affect = 0
            elif self.affect == "terminate":
                cfg.cb_affect = 1
            elif self.affect == "negate":
                cfg.cb_affect = 3
            else:
                _, idx, val = self.affect
                cfg.cb_affect, cfg.cb_idx, par[1] = 2, int(idx), float(val)
        cfg.cb_save_positions = int(self.save_positions[0]) | (int(self.save_positions[1]) << 1)
        cfg.cb_par0, cfg.cb_par1, cfg.cb_par2, cfg.cb_par3 = par


def _apply_affect(cfg, affect, par):
    if affect is None:
        cfg.cb_affect = 0
    elif affect == "terminate":
        cfg.cb_affect = 1
    elif affect == "negate":
        cfg.cb_affect = 3
    else:
        _, idx, val = affect
        cfg.cb_affect, cfg.cb_idx, par[1] = 2, int(idx), float(val)


class ContinuousCallback:
    """ContinuousCallback(condition, affect!, affect_neg! = affect!; interp_points = 10, save_positions = (true, true)) of
    solve(...; callback = cb): an event is a sign change of the real-valued `condition` inside a step; it is located on the
    step's dense output (upstream find_callback_time: sign test at `interp_points` samples, ITP bracketing, left root), the
    step is cut there (change_t_via_interpolation!, src/integrators/interface.jl:566-572; slopes recomputed, :597-632) and
    `affect!` runs, after which handle_callback_modifiers! (:585-594) adds an order-0 discontinuity.  `condition` is
    ("time", tc) — `t - tc`, the callback of test/integrators/events.jl:9-12 —, ("state", component, level) —
    `u[component] - level` — or "model" (the device functor cb_value(u, t, p, par) of a UserModel with HAS_CONT_CALLBACK);
    `affect` as for DiscreteCallback (None with "model": the model's cb_affect).  `direction` +1 / -1 keeps only up- /
    downcrossings (affect_neg! = nothing / affect! = nothing).  Tsit5 and BS3."""

    def __init__(self, condition, affect=None, direction=0, interp_points=10, save_positions=(True, True), par=()):
        self.condition, self.affect, self.direction, self.interp_points = condition, affect, int(direction), int(interp_points)
        self.save_positions = (bool(save_positions[0]), bool(save_positions[1]))
        self.par = tuple(float(x) for x in par)

    def tstops(self, tspan):
        return []

    def apply(self, cfg):
        par = [0.0, 0.0, 0.0, 0.0]
        if self.condition == "model":
            cfg.callback = 4
            par[:len(self.par)] = self.par
        elif self.condition[0] == "time":
            cfg.callback = 5
            par[0] = float(self.condition[1])
        elif self.condition[0] == "state":
            cfg.callback = 6
            cfg.cb_cond_idx = int(self.condition[1])
            par[0] = float(self.condition[2])
        else:
            raise B200DDEError(f"ContinuousCallback: unknown condition {self.condition!r}")
        if self.condition != "model" or self.affect is not None:
            _apply_affect(cfg, self.affect, par)  # (("add", i, v) stores v in par[1])
        cfg.cb_direction = self.direction
        cfg.cb_interp_points = self.interp_points
        cfg.cb_save_positions = int(self.save_positions[0]) | (int(self.save_positions[1]) << 1)
        cfg.cb_par0, cfg.cb_par1, cfg.cb_par2, cfg.cb_par3 = par


def PresetTimeCallback(t, affect, save_positions=(True, True)):
    return DiscreteCallback(("at", t), affect, save_positions)


def PeriodicCallback(affect, period, save_positions=(True, True)):
    return DiscreteCallback(("every", period), affect, save_positions)


class MethodOfSteps:
    """MethodOfSteps(alg; constrained=false, fpsolve=NLFunctional()) — src/algorithms.jl:34-36."""

    def __init__(self, alg, constrained=False, fpsolve=None):
        if isinstance(alg, type):
            alg = alg()
        if not isinstance(alg, (Tsit5, Rosenbrock23, BS3)):
            raise B200DDEError("MethodOfSteps: only Tsit5(), BS3() and Rosenbrock23() are implemented on the B200 path")
        self.alg, self.constrained = alg, bool(constrained)
        self.fpsolve = fpsolve if fpsolve is not None else NLFunctional()


# ----------------------------------------------------------------------------------------------
# user models (NVRTC)
# ----------------------------------------------------------------------------------------------
class UserModel:
    """f / h / lag functions of a DDEProblem as CUDA C++ source: a struct with the interface of csrc/models.cuh
    (N, NP, NCL, NDL, NH, ORDER0, HAS_JAC, hist_comp, f, h_pre, dep_lag[, jac, tgrad]), compiled with NVRTC against the
    same kernel template as the registry models.  `dims` = (n_state, n_param, n_const_lags, n_dep_lags)."""

    def __init__(self, source, struct_name, dims):
        self.source, self.struct_name = source, struct_name
        self.n_state, self.n_param, self.n_const_lags, self.n_dep_lags = (int(x) for x in dims)

    def compile_check(self, alg=_lib.ALG_TSIT5):
        """Compile only (works without a GPU); returns the NVRTC log, raises on errors."""
        buf = C.create_string_buffer(1 << 16)
        rc = _lib.lib().b2d_compile_user_model(self.source.encode(), self.struct_name.encode(), alg, buf, len(buf))
        if rc != 0:
            raise B200DDEError(buf.value.decode(errors="replace"))
        return buf.value.decode(errors="replace")


# ----------------------------------------------------------------------------------------------
# problems
# ----------------------------------------------------------------------------------------------
class DDEProblem:
    """DDEProblem(f, u0, h, tspan, p; constant_lags, dependent_lags, neutral, order_discontinuity_t0)
    (upstream SciMLBase; fields consumed at src/solve.jl:107,159).  `f` names a registry model; `h` and
    `dependent_lags` belong to that model and are accepted only for signature compatibility."""

    def __init__(self, f, u0, h=None, tspan=(0.0, 1.0), p=None, *, constant_lags=None, dependent_lags=None,
                 neutral=None, order_discontinuity_t0=None):
        self.user_model = f if isinstance(f, UserModel) else None
        if self.user_model is not None:
            self.model_id = -1
            self.n_state, self.n_param = f.n_state, f.n_param
            self.n_const_lags, self.n_dep_lags, self.model_order0 = f.n_const_lags, f.n_dep_lags, 0
        else:
            if isinstance(f, str):
                if f not in MODELS:
                    raise B200DDEError(f"unknown model {f!r}; registry: {sorted(MODELS)}")
                self.model_id = MODELS[f]
            else:
                self.model_id = int(f)
            n = [C.c_int32() for _ in range(5)]
            _check_model(self.model_id, n)
            self.n_state, self.n_param, self.n_const_lags, self.n_dep_lags, self.model_order0 = (x.value for x in n)
            # DDEFunction(f; mass_matrix): part of the compiled model; singular => isdae (src/solve.jl:155-156)
            diag = np.ones(self.n_state)
            kind = _lib.lib().b2d_model_mass_matrix(self.model_id, diag.ctypes.data, self.n_state)
            self.mass_matrix = diag if kind > 0 else None
            self.isdae = kind == 2
        self.f, self.h = f, h
        self.u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64)).copy()
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.p = np.zeros(self.n_param) if p is None else np.atleast_1d(np.asarray(p, dtype=np.float64)).copy()
        self.constant_lags = np.asarray([] if constant_lags is None else constant_lags, dtype=np.float64)
        self.dependent_lags = dependent_lags
        if neutral is None:  # upstream DDEProblem: neutral = mass_matrix !== I && det(mass_matrix) != 1
            mm = getattr(self, "mass_matrix", None)
            neutral = mm is not None and float(np.prod(mm)) != 1.0
        self.neutral = bool(neutral)
        self.order_discontinuity_t0 = order_discontinuity_t0
        if len(self.u0) != self.n_state:
            raise B200DDEError(f"u0 has {len(self.u0)} states, model expects {self.n_state}")
        if len(self.p) != self.n_param:
            raise B200DDEError(f"p has {len(self.p)} parameters, model expects {self.n_param}")
        if len(self.constant_lags) != self.n_const_lags:
            raise B200DDEError(f"constant_lags has {len(self.constant_lags)} entries, model expects {self.n_const_lags}")

    def remake(self, **kw):
        args = dict(f=self.f, u0=self.u0, h=self.h, tspan=self.tspan, p=self.p, constant_lags=self.constant_lags,
                    dependent_lags=self.dependent_lags, neutral=self.neutral,
                    order_discontinuity_t0=self.order_discontinuity_t0)
        args.update(kw)
        return DDEProblem(**args)


def _check_model(model_id, out):
    rc = _lib.lib().b2d_model_info(model_id, *[C.byref(x) for x in out])
    _lib.check(rc)


class EnsembleProblem:
    """EnsembleProblem(prob; prob_func) (upstream SciMLBase).  prob_func(prob, i, repeat) -> DDEProblem is
    evaluated on the host to build the u0/p matrices, exactly what the Julia shim does before its ccall.
    `u0`/`p` arrays ([trajectories, n]) are the vectorised alternative."""

    def __init__(self, prob, prob_func=None, output_func=None, u0=None, p=None, constant_lags=None):
        self.prob, self.prob_func, self.output_func = prob, prob_func, output_func
        self.u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
        self.p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
        # per-trajectory constant lags ([trajectories, n_const_lags]): given here, or collected from a prob_func that
        # changes constant_lags (matrices() leaves them in lags_matrix); None: the problem's lags for every trajectory
        self.lags_matrix = None if constant_lags is None else np.ascontiguousarray(constant_lags, dtype=np.float64)

    def matrices(self, trajectories):
        prob = self.prob
        if self.prob_func is not None:
            u0 = np.empty((trajectories, prob.n_state))
            p = np.empty((trajectories, prob.n_param))
            lags = np.empty((trajectories, len(prob.constant_lags)))
            for i in range(trajectories):
                pi = self.prob_func(prob, i, 1)
                # everything but u0, p and the values of the constant lags is shared by the ensemble on this path (the
                # reference unpacks the whole problem per trajectory, src/solve.jl:159): a prob_func that changes anything
                # else is refused instead of silently ignored
                for field in ("tspan", "neutral", "order_discontinuity_t0", "model_id", "dependent_lags"):
                    if getattr(pi, field) != getattr(prob, field):
                        raise B200DDEError(f"prob_func changed {field} for trajectory {i}; only u0, p and constant_lags may vary")
                if len(pi.constant_lags) != len(prob.constant_lags):
                    raise B200DDEError(f"prob_func changed the number of constant lags for trajectory {i}; only u0, p and constant_lags may vary")
                u0[i], p[i], lags[i] = pi.u0, pi.p, pi.constant_lags
            self.lags_matrix = lags if (lags != np.asarray(prob.constant_lags, dtype=np.float64)[None, :]).any() else None
            return u0, p
        u0 = self.u0 if self.u0 is not None else np.tile(prob.u0, (trajectories, 1))
        p = self.p if self.p is not None else np.tile(prob.p, (trajectories, 1))
        if self.lags_matrix is not None and self.lags_matrix.shape != (trajectories, len(prob.constant_lags)):
            raise B200DDEError(f"constant_lags has shape {self.lags_matrix.shape}; expected ({trajectories},{len(prob.constant_lags)})")
        if u0.shape != (trajectories, prob.n_state) or p.shape != (trajectories, prob.n_param):
            raise B200DDEError(f"ensemble matrices have shapes {u0.shape}, {p.shape}; expected "
                               f"({trajectories},{prob.n_state}), ({trajectories},{prob.n_param})")
        return u0, p


class EnsembleB200:
    """Ensemble algorithm selecting this library (the Julia shim's `struct EnsembleB200 <: EnsembleAlgorithm`).
    `devices=[0, 1, ...]` shards the ensemble over several GPUs from this one process (b2d_solve_multi)."""

    def __init__(self, device=0, devices=None):
        self.devices = [int(d) for d in devices] if devices else [int(device)]
        self.device = self.devices[0]


# ----------------------------------------------------------------------------------------------
# solutions
# ----------------------------------------------------------------------------------------------
class DEStats:
    def __init__(self, row):
        for name, v in zip(STAT_NAMES, row):
            setattr(self, name, int(v))

    def __repr__(self):
        return "DEStats(" + ", ".join(f"{k}={getattr(self, k)}" for k in STAT_NAMES) + ")"


_TS5_R = np.array([[1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
                   [0.0, 0.13169999999999998, -0.2234, 0.1017],
                   [0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253],
                   [0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902],
                   [0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928],
                   [0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661],
                   [0.0, 1.5, -4.0, 2.5]])


class DDESolution:
    """sol.t, sol.u, sol.retcode, sol.stats and — when the dense output was requested — sol(t)."""

    def __init__(self, t, u, retcode, stats, tracked_discontinuities=None, k=None, alg_id=0):
        self.t, self.u = t, u
        self.retcode = RETCODES.get(int(retcode), str(retcode))
        self.stats = DEStats(stats)
        self.tracked_discontinuities = tracked_discontinuities
        self.k, self.alg_id = k, alg_id

    def __len__(self):
        return len(self.t)

    def __call__(self, tq):
        """Dense output on the host (upstream ode_interpolation, left-continuous): convenience evaluation of the
        interpolant from the slopes the GPU returned; plain numpy arithmetic (not the kernel's fused order)."""
        if self.k is None:
            raise B200DDEError("sol(t) needs the dense output: solve(prob, alg, dense=True)")
        ts = self.t
        tdir = 1.0 if ts[-1] >= ts[0] else -1.0
        ip = 1
        while ip < len(ts) - 1 and tdir * ts[ip] < tdir * tq:
            ip += 1
        im = ip - 1
        dt = ts[ip] - ts[im]
        th = 1.0 if dt == 0 else (tq - ts[im]) / dt
        y0, y1, k = self.u[im], self.u[ip], self.k[ip]
        if self.alg_id == _lib.ALG_TSIT5:
            b = th * (_TS5_R[:, 0] + th * (_TS5_R[:, 1] + th * (_TS5_R[:, 2] + th * _TS5_R[:, 3])))
            return y0 + dt * (b[:, None] * k).sum(axis=0)
        if self.alg_id == _lib.ALG_BS3:
            return (1 - th) * y0 + th * y1 + th * (th - 1) * ((1 - 2 * th) * (y1 - y0) + (th - 1) * dt * k[0] + th * dt * k[1])
        d = 1.0 / (2.0 + np.sqrt(2.0))
        c1, c2 = th * (1 - th) / (1 - 2 * d), th * (th - 2 * d) / (1 - 2 * d)
        return y0 + dt * (c1 * k[0] + c2 * k[1])


class EnsembleSolution:
    """u[i] is trajectory i's DDESolution; `.t` is the shared save grid and `.array` the [traj, n_out, n] block."""

    def __init__(self, t, array, retcodes, stats, elapsed):
        self.t, self.array, self.retcodes, self.stats, self.elapsedTime = t, array, retcodes, stats, elapsed
        self.converged = bool(np.all(retcodes == 1))

    def __len__(self):
        return self.array.shape[0]

    def __getitem__(self, i):
        return DDESolution(self.t, self.array[i], self.retcodes[i], self.stats[i])


# ----------------------------------------------------------------------------------------------
# solve
# ----------------------------------------------------------------------------------------------
def _rationalize(x):
    f = Fraction(x).limit_denominator(10 ** 9)
    return f if float(f) == x else Fraction(x)


def saveat_grid(saveat, tspan):
    """The points the reference pushes into its saveat heap (upstream initialize_saveat, called at src/solve.jl:262):
    a Number means t0+s : s : tf; only points strictly inside (t0, tf) are kept.  Julia builds Float64 ranges from
    rationalised endpoints, so 0.1:0.1:10 yields the correctly rounded k/10; reproduced here with Fractions."""
    t0, tf = tspan
    tdir = 1.0 if tf > t0 else -1.0
    if np.isscalar(saveat):
        step = _rationalize(abs(float(saveat))) * (1 if tdir > 0 else -1)
        a, b = _rationalize(t0), _rationalize(tf)
        n = int((b - a) / step)
        pts = [float(a + k * step) for k in range(1, n + 1)]
    else:
        pts = [float(s) for s in saveat]
    return np.asarray([s for s in pts if tdir * t0 < tdir * s < tdir * tf], dtype=np.float64)


def _make_config(prob, alg, device, kw):
    L = _lib.lib()
    cfg = Config()
    L.b2d_config_default(C.byref(cfg), alg.alg.alg_id)
    cfg.constrained = int(alg.constrained)
    cfg.fp_max_iter = alg.fpsolve.max_iter
    cfg.fp_kappa = alg.fpsolve.kappa
    if isinstance(alg.fpsolve, NLAnderson):
        cfg.fp_alg = 1
        cfg.fp_max_history, cfg.fp_aa_start = alg.fpsolve.max_history, alg.fpsolve.aa_start
    cfg.model_id = max(prob.model_id, 0)
    cfg.n_state, cfg.n_param = prob.n_state, prob.n_param
    cfg.n_const_lags, cfg.n_dep_lags = prob.n_const_lags, prob.n_dep_lags
    cfg.neutral = int(prob.neutral)
    cfg.order_discontinuity_t0 = -1 if prob.order_discontinuity_t0 is None else int(prob.order_discontinuity_t0)
    cfg.device = device
    names = {"abstol": "abstol", "reltol": "reltol", "dt": "dt0", "dtmin": "dtmin", "dtmax": "dtmax",
             "maxiters": "maxiters", "gamma": "gamma", "qmin": "qmin", "qmax": "qmax", "qsteady_min": "qsteady_min",
             "qsteady_max": "qsteady_max", "qoldinit": "qoldinit", "failfactor": "failfactor",
             "discontinuity_interp_points": "disc_interp_points", "discontinuity_abstol": "disc_abstol",
             "discontinuity_reltol": "disc_reltol", "ring_capacity": "ring_capacity", "chunk_traj": "chunk_traj",
             "controller_pow": "controller_pow", "fp_div_rule": "fp_div_rule", "track_theta_mode": "track_theta_mode",
             "rb23_dT_mode": "rb23_dT_mode", "fp_anderson_reset": "fp_anderson_reset"}
    for k, v in kw.items():
        if k not in names:
            raise TypeError(f"solve() got an unexpected keyword argument {k!r}")
        if v is not None:
            setattr(cfg, names[k], v)
    return cfg


class Solver:
    """Owns one b2d_handle (device buffers, streams, ring) — reusable across solves of the same configuration."""

    def __init__(self, cfg, user_model=None):
        self.cfg = cfg
        self.h = C.c_void_p()
        if user_model is not None:
            _lib.check(_lib.lib().b2d_create_user_model(C.byref(self.h), C.byref(cfg), user_model.source.encode(),
                                                        user_model.struct_name.encode()))
        else:
            _lib.check(_lib.lib().b2d_create(C.byref(self.h), C.byref(cfg)))

    def close(self):
        if self.h:
            _lib.lib().b2d_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tolerances(self, abstol, reltol):
        """Per-component abstol / reltol arrays (src/utils.jl:91-133); None resets to the scalars of the config."""
        if abstol is None:
            _lib.check(_lib.lib().b2d_set_tolerances(self.h, None, None, 0))
            return
        a = np.ascontiguousarray(abstol, dtype=np.float64)
        r = np.ascontiguousarray(reltol, dtype=np.float64)
        _lib.check(_lib.lib().b2d_set_tolerances(self.h, a.ctypes.data, r.ctypes.data, len(a)))

    def set_save_idxs(self, idxs):
        """save_idxs (src/solve.jl:47,217-218), 0-based component indices; None saves every component."""
        self.n_save = 0
        if idxs is None:
            _lib.check(_lib.lib().b2d_set_save_idxs(self.h, None, 0))
            return
        i = np.ascontiguousarray(idxs, dtype=np.int32)
        _lib.check(_lib.lib().b2d_set_save_idxs(self.h, i.ctypes.data, len(i)))
        self.n_save = len(i)

    def set_trajectory_lags(self, lags):
        """Per-trajectory constant lags ([trajectories, n_const_lags]) for the host-buffer solves that follow: what a prob_func
        that remakes `constant_lags` gives the reference (src/solve.jl:159); None clears."""
        a = np.ascontiguousarray(lags if lags is not None else np.empty((0, 0)), dtype=np.float64)
        _lib.check(_lib.lib().b2d_set_trajectory_lags(self.h, a.ctypes.data if a.size else None, a.shape[0] if a.size else 0))

    def set_tstops(self, tstops):
        """solve(...; tstops) (src/solve.jl:45); None / empty clears."""
        a = np.ascontiguousarray(tstops if tstops is not None else [], dtype=np.float64)
        _lib.check(_lib.lib().b2d_set_tstops(self.h, a.ctypes.data if len(a) else None, len(a)))

    def set_d_discontinuities(self, discs):
        """solve(...; d_discontinuities) (src/solve.jl:46): Discontinuity objects or (t, order) pairs; numbers get order 0."""
        ts, os_ = [], []
        for d in (discs if discs is not None else []):
            if isinstance(d, Discontinuity):
                ts.append(d.t); os_.append(d.order)
            elif np.ndim(d) == 0:
                ts.append(float(d)); os_.append(0)
            else:
                ts.append(float(d[0])); os_.append(int(d[1]))
        t = np.ascontiguousarray(ts, dtype=np.float64)
        o = np.ascontiguousarray(os_, dtype=np.int32)
        _lib.check(_lib.lib().b2d_set_d_discontinuities(self.h, t.ctypes.data if len(t) else None,
                                                        o.ctypes.data if len(o) else None, len(t)))

    def comm_init(self, n_ranks, rank, unique_id):
        """Join the NCCL communicator of a multi-process host (b2d_comm_init); `unique_id`: the 128 bytes rank 0 got
        from `comm_unique_id()`."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        _lib.check(_lib.lib().b2d_comm_init(self.h, int(n_ranks), int(rank), C.addressof(buf)))

    def gather_saveat(self, d_local, counts, root, d_root, stream=0):
        """b2d_gather_saveat with raw device pointers (ints); counts: doubles per rank."""
        c = np.ascontiguousarray(counts, dtype=np.int64)
        _lib.check(_lib.lib().b2d_gather_saveat(self.h, d_local, c.ctypes.data, int(root), d_root, stream))

    def last_timing(self):
        ms, nl, cap = C.c_double(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().b2d_last_timing(self.h, C.byref(ms), C.byref(nl), C.byref(cap)))
        return ms.value, nl.value, cap.value

    def solve_host(self, u0, p, lags, tspan, grid, save_start, save_end, out=None, rc=None, stats=None):
        """b2d_solve with host numpy buffers.  u0: [traj, n], p: [traj, np] (C-order rows = Julia columns).
        `out`, `rc`, `stats` may be caller-owned arrays that are reused from call to call (the C ABI never allocates
        results; fresh 64 MB arrays cost page faults on every call)."""
        n_traj, n = u0.shape
        n = getattr(self, "n_save", 0) or n   # save_idxs: fewer components per saved point
        n_out = int(save_start) + len(grid) + int(save_end)
        if out is None:
            out = np.empty((n_traj, n_out, n), dtype=np.float64)
        out_t = np.empty(n_out, dtype=np.float64)
        if rc is None:
            rc = np.empty(n_traj, dtype=np.int32)
        if stats is None:
            stats = np.empty((n_traj, N_STATS), dtype=np.int64)
        assert rc.dtype == np.int32 and rc.shape == (n_traj,) and rc.flags.c_contiguous
        assert stats.dtype == np.int64 and stats.shape == (n_traj, N_STATS) and stats.flags.c_contiguous
        lags = np.ascontiguousarray(lags, dtype=np.float64)
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        r = _lib.lib().b2d_solve(self.h, n_traj, tspan[0], tspan[1], u0.ctypes.data, p.ctypes.data, lags.ctypes.data,
                                 grid.ctypes.data, len(grid), int(save_start), int(save_end), out.ctypes.data,
                                 out_t.ctypes.data, rc.ctypes.data, stats.ctypes.data)
        _lib.check(r)
        return out_t, out, rc, stats

    def solve_device(self, d_u0, d_p, lags, tspan, grid, save_start, save_end, d_out, d_rc, d_stats, n_traj, stream=0):
        """b2d_solve_device with raw device pointers (ints), asynchronous on `stream`."""
        lags = np.ascontiguousarray(lags, dtype=np.float64)
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        r = _lib.lib().b2d_solve_device(self.h, n_traj, tspan[0], tspan[1], d_u0, d_p, lags.ctypes.data,
                                        grid.ctypes.data, len(grid), int(save_start), int(save_end), d_out, d_rc,
                                        d_stats, stream)
        _lib.check(r)

    def solve_everystep(self, u0, p, lags, tspan, max_steps=4096, max_tracked=64, dense=False):
        n_traj, n = u0.shape
        nk = 7 if self.cfg.alg == _lib.ALG_TSIT5 else 2
        out_t = np.empty((n_traj, max_steps))
        out_u = np.empty((n_traj, max_steps, n))
        out_k = np.empty((n_traj, max_steps, nk, n)) if dense else None
        ns = np.empty(n_traj, dtype=np.int32)
        tr_t = np.empty((n_traj, max_tracked))
        tr_o = np.empty((n_traj, max_tracked), dtype=np.int32)
        tr_n = np.empty(n_traj, dtype=np.int32)
        rc = np.empty(n_traj, dtype=np.int32)
        stats = np.empty((n_traj, N_STATS), dtype=np.int64)
        lags = np.ascontiguousarray(lags, dtype=np.float64)
        r = _lib.lib().b2d_solve_dense(self.h, n_traj, tspan[0], tspan[1], u0.ctypes.data, p.ctypes.data,
                                       lags.ctypes.data, max_steps, out_t.ctypes.data, out_u.ctypes.data,
                                       out_k.ctypes.data if dense else None, ns.ctypes.data, max_tracked, tr_t.ctypes.data,
                                       tr_o.ctypes.data, tr_n.ctypes.data, rc.ctypes.data, stats.ctypes.data)
        _lib.check(r)
        sols = []
        for i in range(n_traj):
            m = min(int(ns[i]), max_steps)
            k = min(int(tr_n[i]), max_tracked)
            tracked = [(float(tr_t[i, j]), int(tr_o[i, j])) for j in range(k)]
            sols.append(DDESolution(out_t[i, :m].copy(), out_u[i, :m].copy(), rc[i], stats[i], tracked,
                                    k=out_k[i, :m].copy() if dense else None, alg_id=self.cfg.alg))
        return sols


def comm_unique_id():
    """The ncclUniqueId (128 bytes) a multi-process host distributes from rank 0 to every rank (b2d_comm_unique_id)."""
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.lib().b2d_comm_unique_id(C.addressof(buf)))
    return bytes(buf)


def devices_of(ensemblealg):
    return ensemblealg.devices if isinstance(ensemblealg, EnsembleB200) else [0]


def solve_multi(solvers, u0, p, lags, tspan, grid, save_start, save_end):
    """b2d_solve_multi over a list of Solver objects (one per GPU)."""
    n_traj, n = u0.shape
    n = getattr(solvers[0], "n_save", 0) or n   # save_idxs (set identically on every solver)
    n_out = int(save_start) + len(grid) + int(save_end)
    out = np.empty((n_traj, n_out, n), dtype=np.float64)
    out_t = np.empty(n_out, dtype=np.float64)
    rc = np.empty(n_traj, dtype=np.int32)
    stats = np.empty((n_traj, N_STATS), dtype=np.int64)
    lags = np.ascontiguousarray(lags, dtype=np.float64)
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    hs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    r = _lib.lib().b2d_solve_multi(hs, len(solvers), n_traj, tspan[0], tspan[1], u0.ctypes.data, p.ctypes.data,
                                   lags.ctypes.data, grid.ctypes.data, len(grid), int(save_start), int(save_end),
                                   out.ctypes.data, out_t.ctypes.data, rc.ctypes.data, stats.ctypes.data)
    _lib.check(r)
    return out_t, out, rc, stats


def solve(prob, alg, ensemblealg=None, *, trajectories=None, saveat=(), save_start=None, save_end=None,
          max_steps=4096, solver=None, dense=None, save_idxs=None, tstops=None, d_discontinuities=None, callback=None, **kw):
    """solve(prob, alg; kwargs...) / solve(ensembleprob, alg, EnsembleB200(); trajectories, kwargs...).

    Keyword defaults follow src/solve.jl:44-101: with an empty saveat every step is saved (save_everystep);
    save_start/save_end default to true when saveat is empty, a Number, or contains the end point (:50-52,326-329).
    """
    import time
    if not isinstance(alg, MethodOfSteps):
        raise B200DDEError("solve: algorithm must be MethodOfSteps(...)")
    ens = prob if isinstance(prob, EnsembleProblem) else None
    base = ens.prob if ens else prob
    device = ensemblealg.device if isinstance(ensemblealg, EnsembleB200) else 0
    t0, tf = base.tspan
    scalar_saveat = np.isscalar(saveat)
    empty_saveat = (not scalar_saveat) and len(saveat) == 0
    if save_start is None:
        save_start = empty_saveat or scalar_saveat or (t0 in list(saveat))
    if save_end is None:
        save_end = empty_saveat or scalar_saveat or (tf in list(saveat))
    # abstol / reltol may be arrays (one entry per component, src/utils.jl:91-133)
    tol_vec = None
    if any(np.ndim(kw.get(k)) > 0 for k in ("abstol", "reltol")):
        a = np.broadcast_to(np.asarray(kw.get("abstol", 1e-6) if kw.get("abstol") is not None else 1e-6, dtype=np.float64), (base.n_state,))
        r = np.broadcast_to(np.asarray(kw.get("reltol", 1e-3) if kw.get("reltol") is not None else 1e-3, dtype=np.float64), (base.n_state,))
        tol_vec = (a.copy(), r.copy())
        kw = {k: v for k, v in kw.items() if k not in ("abstol", "reltol")}
    own = solver is None
    if callback is not None:
        if not own:
            raise B200DDEError("solve(...; callback) configures the handle: pass no `solver`")
        tstops = list(tstops or []) + callback.tstops(base.tspan)

    def make_cfg(dev):
        cfg = _make_config(base, alg, dev, kw)
        if callback is not None:
            callback.apply(cfg)
        return cfg
    if own:
        solver = Solver(make_cfg(device), base.user_model)
    def apply_options(sv):
        if tol_vec is not None:
            sv.set_tolerances(*tol_vec)
        if save_idxs is not None:
            sv.set_save_idxs(save_idxs)
        if tstops is not None:
            sv.set_tstops(tstops)
        if d_discontinuities is not None:
            sv.set_d_discontinuities(d_discontinuities)
    apply_options(solver)
    try:
        if ens is None:
            u0 = base.u0[None, :].copy()
            p = base.p[None, :].copy()
            if empty_saveat:
                # dense = save_everystep && isempty(saveat) (src/solve.jl:55)
                return solver.solve_everystep(u0, p, base.constant_lags, base.tspan, max_steps=max_steps,
                                              dense=True if dense is None else dense)[0]
            grid = saveat_grid(saveat, base.tspan)
            t, u, rc, st = solver.solve_host(u0, p, base.constant_lags, base.tspan, grid, save_start, save_end)
            return DDESolution(t, u[0], rc[0], st[0])
        if trajectories is None:
            raise B200DDEError("solve(EnsembleProblem, ...): `trajectories` is required")
        u0, p = ens.matrices(int(trajectories))
        if ens.lags_matrix is not None:
            solver.set_trajectory_lags(ens.lags_matrix)
        tic = time.perf_counter()
        if empty_saveat:
            sols = solver.solve_everystep(u0, p, base.constant_lags, base.tspan, max_steps=max_steps, dense=bool(dense))
            return sols
        grid = saveat_grid(saveat, base.tspan)
        devices = ensemblealg.devices if isinstance(ensemblealg, EnsembleB200) else [0]
        if len(devices) > 1:
            extra = [Solver(make_cfg(d), base.user_model) for d in devices[1:]]
            try:
                for e in extra:
                    apply_options(e)
                    if ens.lags_matrix is not None:
                        e.set_trajectory_lags(ens.lags_matrix)  # every handle gets the whole array
                t, u, rc, st = solve_multi([solver] + extra, u0, p, base.constant_lags, base.tspan, grid, save_start, save_end)
            finally:
                for e in extra:
                    e.close()
        else:
            t, u, rc, st = solver.solve_host(u0, p, base.constant_lags, base.tspan, grid, save_start, save_end)
        return EnsembleSolution(t, u, rc, st, time.perf_counter() - tic)
    finally:
        if own:
            solver.close()
        elif ens is not None and ens.lags_matrix is not None:
            solver.set_trajectory_lags(None)  # a caller's handle does not keep this ensemble's lags
