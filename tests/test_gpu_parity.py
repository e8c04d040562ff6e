"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bar: BIT-EXACT.  The kernel and the oracle are two independent implementations (different data structures:
ring + cursors + sorted queues vs growing vectors + binary search + heaps) that share only the arithmetic
specification (explicit fma placement) and csrc/detmath.h, so sol.t, sol.u, saveat values, statistics, return
codes and discontinuity times must agree to the last bit.  north_star asks for 1e-6 relative at saveat points and
bit-exact discontinuity indexing; bit-exact everywhere is the stronger statement and is what is asserted.
"""
import json
import os

import numpy as np
import pytest

import b200dde as B
import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    # name: (registry model, oracle id, u0, p, lags, tspan)
    "bc_model": ("bc_model", O.MODEL_BC, [1.0, 1.0, 1.0], [0.2, 0.3, 1.0], [1.0], (0.0, 10.0)),
    "mackey_glass": ("mackey_glass", O.MODEL_MACKEY_GLASS, [1.2], [0.2, 1.2], [17.0], (0.0, 1000.0)),
    "logistic_p03": ("delayed_logistic", O.MODEL_LOGISTIC, [0.1], [0.3], [1.0], (0.0, 50.0)),
    "logistic_p14": ("delayed_logistic", O.MODEL_LOGISTIC, [0.1], [1.4], [1.0], (0.0, 50.0)),
    "one_delay": ("constant_1delay", O.MODEL_1DELAY, [1.0], [0.0], [1.0], (0.0, 10.0)),
    "two_delays_long": ("constant_2delays", O.MODEL_2DELAYS, [1.0], [0.0], [1 / 3, 1 / 5], (0.0, 100.0)),
    "one_delay_long": ("constant_1delay_long", O.MODEL_1DELAY_LONG, [1.0], [0.0], [0.2], (0.0, 100.0)),
    "backwards": ("backwards", O.MODEL_BACKWARDS, [1.0], [0.0], [-1.0], (2.0, 0.0)),
    "state_dependent": ("state_dependent", O.MODEL_STATE_DEP, [1.0], [0.0], [], (0.0, 10.0)),
    "one_delay_dependent": ("constant_1delay_dependent", O.MODEL_1DELAY_DEP, [1.0], [0.0], [], (0.0, 10.0)),
}


def gpu_everystep(case, alg=None, **kw):
    model, _, u0, p, lags, tspan = CASES[case]
    prob = B.DDEProblem(model, u0, None, tspan, p, constant_lags=lags)
    return B.solve(prob, alg or B.MethodOfSteps(B.Tsit5()), **kw)


def oracle_everystep(case, **kw):
    _, oid, u0, p, lags, tspan = CASES[case]
    okw = {k: v for k, v in kw.items() if k not in ("max_steps",)}
    if "dt" in okw:
        okw["dt0"] = okw.pop("dt")
    return O.solve_everystep(O.default_config(model=oid, **okw), u0, p, lags, tspan[0], tspan[1])


def assert_same_ensemble(sol, ref):
    assert np.array_equal(sol.t, ref["t"])
    assert np.array_equal(sol.retcodes, ref["retcode"])
    assert np.array_equal(sol.array, ref["u"], equal_nan=True)
    assert np.array_equal(sol.stats[:, :6], ref["stats"][:, :6])


def assert_same_solution(s, o):
    assert s.retcode == B.RETCODES[o.retcode]
    assert len(s) == len(o)
    assert np.array_equal(s.t, o.t), "sol.t differs"
    assert np.array_equal(s.u, o.u), "sol.u differs"
    assert s.tracked_discontinuities == o.tracked, "tracked discontinuities differ"
    got = [getattr(s.stats, k) for k in ("naccept", "nreject", "nf", "nfpiter", "nfpconvfail", "ntracked")]
    assert got == o.stats[:6].tolist()


@pytest.mark.parametrize("case", sorted(CASES))
def test_everystep_bit_exact(case):
    """Every accepted step of every KAT/benchmark problem: sol.t, sol.u, tracked discontinuities, stats."""
    assert_same_solution(gpu_everystep(case), oracle_everystep(case))


@pytest.mark.parametrize("case,kw", [
    ("bc_model", dict(abstol=1e-10, reltol=1e-10)),              # 91 live history nodes -> ring regrowth path
    ("mackey_glass", dict(abstol=1e-8, reltol=1e-8)),
    ("one_delay_long", dict(abstol=1e-8, reltol=1e-10)),
    ("two_delays_long", dict(controller_pow=0)),                   # exact-pow controller variant
    ("two_delays_long", dict(fp_div_rule=0)),                      # Hairer-type divergence estimate
    ("two_delays_long", dict(fp_div_rule=2)),
    ("state_dependent", dict(track_theta_mode=1)),
    ("one_delay_dependent", dict(abstol=1e-9, reltol=1e-6)),       # dependent_delays.jl:33 tolerances
    ("one_delay", dict(dt=0.01)),                                  # user initial dt
])
def test_everystep_options_bit_exact(case, kw):
    gk = dict(kw)
    assert_same_solution(gpu_everystep(case, max_steps=20000, **gk), oracle_everystep(case, **kw))


def test_constrained_bit_exact():
    alg = B.MethodOfSteps(B.Tsit5(), constrained=True)
    s = gpu_everystep("one_delay_long", alg=alg)
    o = oracle_everystep("one_delay_long", constrained=1)
    assert_same_solution(s, o)
    assert s.stats.nfpiter == 0 and len(set(s.t.tolist())) == len(s.t)   # unique_times.jl:6-10


def test_fpsolve_options_bit_exact():
    alg = B.MethodOfSteps(B.Tsit5(), fpsolve=B.NLFunctional(max_iter=100))
    s = gpu_everystep("one_delay_long", alg=alg, abstol=1e-12, reltol=1e-12, max_steps=20000)
    o = oracle_everystep("one_delay_long", abstol=1e-12, reltol=1e-12, fp_max_iter=100)
    assert_same_solution(s, o)


@pytest.mark.parametrize("name", ["c1_bc_model", "logistic_p03", "mackey_glass"])
def test_against_committed_golden_vectors(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    model = {"MODEL_BC": "bc_model", "MODEL_LOGISTIC": "delayed_logistic", "MODEL_MACKEY_GLASS": "mackey_glass"}[g["model"]]
    prob = B.DDEProblem(model, g["u0"], None, tuple(g["tspan"]), g["p"], constant_lags=g["lags"])
    s = B.solve(prob, B.MethodOfSteps(B.Tsit5()))
    assert s.t.tolist() == g["t"] and s.u.tolist() == g["u"]
    assert [list(x) for x in s.tracked_discontinuities] == g["tracked"]


def test_reference_kats_on_gpu():
    """The reference's own known answers, asserted directly on the GPU result (no oracle involved)."""
    s = gpu_everystep("logistic_p03")
    assert len(s) == 21 and abs(s.u[-1, 0] - 1.0) < 1e-5             # test/interface/parameters.jl:28,30
    assert abs(s(12.0)[0] - 0.884) < 1e-4                            # :29  sol1(12): dense output returned by the GPU
    s = gpu_everystep("logistic_p14")
    assert len(s) == 47 and abs(s.u[-1, 0] - 0.994) < 2e-5           # :35,37
    assert abs(s(12.0)[0] - 1.125) < 5e-4                            # :36
    s = gpu_everystep("one_delay")
    for t in range(6):                                               # test/interface/discontinuities.jl:92-95
        assert float(t) in s.t and (float(t), t) in s.tracked_discontinuities
    s = gpu_everystep("backwards")
    assert s.tracked_discontinuities == [(-2.0, 1), (-1.0, 2)]        # test/interface/backwards.jl:36-37
    assert gpu_everystep("one_delay", maxiters=1).retcode == "MaxIters"          # test/integrators/retcode.jl:16
    assert gpu_everystep("one_delay", dtmin=5.0).retcode == "DtLessThanMin"      # :19


# ---------------------------------------------------------------------------------------------------------
# MethodOfSteps(Rosenbrock23()): per-thread LU, analytic J / time gradient through the history derivative
# ---------------------------------------------------------------------------------------------------------
def rb23_both(model, oid, u0, p, lags, tspan, **kw):
    prob = B.DDEProblem(model, u0, None, tspan, p, constant_lags=lags)
    s = B.solve(prob, B.MethodOfSteps(B.Rosenbrock23()), max_steps=20000, **kw)
    o = O.solve_everystep(O.default_config(alg=O.ALG_ROSENBROCK23, model=oid, **kw), u0, p, lags, tspan[0], tspan[1])
    return s, o


@pytest.mark.parametrize("lam", [1e2, 1e3, 1e5])
def test_rosenbrock23_stiff_everystep_bit_exact(lam):
    """BASELINE configs[4] problem at three stiffness levels."""
    s, o = rb23_both("stiff_kinetics", O.MODEL_STIFF, [2.0, 0.5, 1.0], [lam], [1.0], (0.0, 10.0))
    assert_same_solution(s, o)
    assert s.stats.nsolve == int(o.stats[O.STAT_NSOLVE]) and s.stats.nsolve > 0
    assert abs(s.u[-1, 0] - 0.84) < 0.01 and abs(s.u[-1, 2] - 1.0014) < 1e-3   # SURVEY §8d scratch evidence


def test_rosenbrock23_one_delay_bit_exact():
    # test/integrators/rosenbrock.jl:15-25 problem (prob_dde_constant_1delay_ip) — fixed point + discontinuities under RB23
    s, o = rb23_both("constant_1delay", O.MODEL_1DELAY, [1.0], [0.0], [1.0], (0.0, 10.0))
    assert_same_solution(s, o)
    s, o = rb23_both("constant_1delay", O.MODEL_1DELAY, [1.0], [0.0], [1.0], (0.0, 10.0), rb23_dT_mode=1)
    assert_same_solution(s, o)


def test_rosenbrock23_needs_a_jacobian():
    prob = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])
    with pytest.raises(B.B200DDEError, match="Jacobian"):
        B.solve(prob, B.MethodOfSteps(B.Rosenbrock23()))


def test_rosenbrock23_stiff_ensemble_bit_exact():
    n = 3000
    lam = 10.0 ** (2.0 + 3.0 * np.arange(n) / (n - 1))
    u0 = np.tile([2.0, 0.5, 1.0], (n, 1))
    p = lam[:, None].copy()
    prob = B.DDEProblem("stiff_kinetics", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.Rosenbrock23()), B.EnsembleB200(0), trajectories=n,
                  saveat=0.1)
    ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_STIFF), u0, p, [1.0], 0.0, 10.0,
                  saveat=B.saveat_grid(0.1, (0.0, 10.0)), nthreads=O.hardware_threads())
    assert sol.converged
    assert_same_ensemble(sol, ref)
    assert np.array_equal(sol.stats[:, 7], ref["stats"][:, 7])


# ---------------------------------------------------------------------------------------------------------
# MethodOfSteps(BS3()): the algorithm of most of the reference's analytic-error KATs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["one_delay", "two_delays_long", "one_delay_long", "one_delay_dependent", "bc_model", "mackey_glass"])
def test_bs3_everystep_bit_exact(case):
    model, oid, u0, p, lags, tspan = CASES[case]
    prob = B.DDEProblem(model, u0, None, tspan, p, constant_lags=lags)
    s = B.solve(prob, B.MethodOfSteps(B.BS3()), max_steps=20000)
    o = O.solve_everystep(O.default_config(alg=O.ALG_BS3, model=oid), u0, p, lags, tspan[0], tspan[1])
    assert_same_solution(s, o)


def test_bs3_reference_kats_on_gpu():
    import math
    ana = lambda t: sum((-1) ** k * (t - k) ** k / math.factorial(k) for k in range(int(math.floor(t)) + 1))
    prob = B.DDEProblem("constant_1delay", [1.0], None, (0.0, 10.0), [0.0], constant_lags=[1.0])
    for alg, kw in ((B.MethodOfSteps(B.BS3()), {}), (B.MethodOfSteps(B.BS3(), constrained=True), dict(dt=0.1))):
        s = B.solve(prob, alg, **kw)
        e = np.array([abs(s.u[i, 0] - ana(s.t[i])) for i in range(len(s))])
        # test/interface/unconstrained.jl:31-33, constrained.jl:14-16
        assert e.max() < 3.0e-5 and e[-1] < 2.1e-5 and math.sqrt(np.mean(e ** 2)) < 1.3e-5
    prob2 = B.DDEProblem("constant_2delays", [1.0], None, (0.0, 1.0), [0.0], constant_lags=[1 / 3, 1 / 5])
    s = B.solve(prob2, B.MethodOfSteps(B.BS3()))
    want = [(0.0, 0), (1 / 5, 1), (1 / 3, 1), (2 / 5, 2), (8 / 15, 2), (3 / 5, 3), (2 / 3, 2), (11 / 15, 3), (13 / 15, 3)]
    assert len(s.tracked_discontinuities) == 9                          # test/interface/discontinuities.jl:47-58
    for (tg, og), (tw, ow) in zip(s.tracked_discontinuities, want):
        assert og == ow and math.isclose(tg, tw, rel_tol=1e-14)


def test_bs3_ensemble_bit_exact():
    u0, p = bc_sweep(5000, seed=21)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.BS3()), B.EnsembleB200(0), trajectories=5000, saveat=0.1)
    ref = O.solve(O.default_config(alg=O.ALG_BS3, model=O.MODEL_BC), u0, p, [1.0], 0.0, 10.0, saveat=B.saveat_grid(0.1, (0.0, 10.0)),
                  nthreads=O.hardware_threads())
    assert_same_ensemble(sol, ref)


@pytest.mark.parametrize("alg,oalg,case", [(B.Tsit5, O.ALG_TSIT5, "bc_model"), (B.BS3, O.ALG_BS3, "one_delay"),
                                           (B.Rosenbrock23, O.ALG_ROSENBROCK23, "one_delay")])
def test_dense_output_matches_oracle(alg, oalg, case):
    """sol.k returned by b2d_solve_dense is bit-identical to the oracle's dense history; sol(t) agrees to rounding."""
    model, oid, u0, p, lags, tspan = CASES[case]
    prob = B.DDEProblem(model, u0, None, tspan, p, constant_lags=lags)
    s = B.solve(prob, B.MethodOfSteps(alg()), max_steps=20000)
    o = O.solve_everystep(O.default_config(alg=oalg, model=oid), u0, p, lags, tspan[0], tspan[1])
    assert np.array_equal(s.k, o.k)
    for tq in np.linspace(tspan[0], tspan[1], 37):
        assert np.allclose(s(tq), o(tq), rtol=1e-12, atol=1e-14)


def test_wrong_sign_lag_is_an_error():
    # test/interface/backwards.jl:21-24: positive lag with reverse time throws
    prob = B.DDEProblem("backwards", [1.0], None, (2.0, 0.0), [0.0], constant_lags=[1.0])
    with pytest.raises(B.B200DDEError, match="wrong sign"):
        B.solve(prob, B.MethodOfSteps(B.Tsit5()))


# ---------------------------------------------------------------------------------------------------------
# ensembles with a saveat grid (b2d_solve)
# ---------------------------------------------------------------------------------------------------------
def bc_sweep(n, seed=0):
    rng = np.random.default_rng(seed)
    p = np.stack([0.1 + 0.2 * rng.random(n), 0.2 + 0.2 * rng.random(n), 0.5 + 1.0 * rng.random(n)], axis=1)
    return np.ones((n, 3)), p


def mg_sweep(n, seed=1):
    rng = np.random.default_rng(seed)
    p = np.stack([0.18 + 0.04 * rng.random(n), 0.5 + 1.0 * rng.random(n)], axis=1)
    return p[:, 1:2].copy(), p


def solve_both(model, oid, u0, p, lags, tspan, saveat, save_start=True, save_end=True, **kw):
    prob = B.DDEProblem(model, u0[0], None, tspan, p[0], constant_lags=lags)
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=len(u0), saveat=saveat,
                  save_start=save_start, save_end=save_end, **kw)
    grid = B.saveat_grid(saveat, tspan)
    okw = {k: v for k, v in kw.items() if k in ("abstol", "reltol", "maxiters", "dtmin")}
    ref = O.solve(O.default_config(model=oid, **okw), u0, p, lags, tspan[0], tspan[1], saveat=grid, save_start=save_start,
                  save_end=save_end, nthreads=O.hardware_threads())
    return sol, ref


def test_bc_ensemble_bit_exact():
    """BASELINE configs[1] at oracle-affordable size: random (unsorted!) parameters stress warp divergence."""
    u0, p = bc_sweep(20000)
    sol, ref = solve_both("bc_model", O.MODEL_BC, u0, p, [1.0], (0.0, 10.0), 0.1)
    assert sol.array.shape == (20000, 101, 3) and sol.converged
    assert_same_ensemble(sol, ref)


def test_mackey_glass_ensemble_bit_exact():
    """BASELINE configs[2]: chaotic, so anything short of identical arithmetic would show up here."""
    u0, p = mg_sweep(6000)
    sol, ref = solve_both("mackey_glass", O.MODEL_MACKEY_GLASS, u0, p, [17.0], (0.0, 1000.0), 10.0)
    assert sol.array.shape == (6000, 101, 1) and sol.converged
    assert_same_ensemble(sol, ref)


def test_mackey_glass_tight_ensemble_bit_exact():
    u0, p = mg_sweep(1500, seed=3)
    sol, ref = solve_both("mackey_glass", O.MODEL_MACKEY_GLASS, u0, p, [17.0], (0.0, 1000.0), 10.0, abstol=1e-8, reltol=1e-8)
    assert_same_ensemble(sol, ref)
    assert int(sol.stats[:, 6].max()) > 16   # needed more history nodes than the default ring: regrowth path ran


def test_state_dependent_ensemble_bit_exact():
    """BASELINE configs[3]: state-dependent lag with discontinuity tracking, initial-condition sweep."""
    n = 5000
    u0 = (0.5 + 1.0 * np.arange(n) / (n - 1))[:, None]
    p = np.zeros((n, 1))
    sol, ref = solve_both("state_dependent", O.MODEL_STATE_DEP, u0, p, [], (0.0, 10.0), 0.1)
    assert_same_ensemble(sol, ref)


@pytest.mark.parametrize("save_start", [False, True])
@pytest.mark.parametrize("save_end", [False, True])
def test_saveat_semantics(save_start, save_end):
    """test/interface/saveat.jl:22-80 on the GPU: grids and values, for scalar and vector saveat."""
    u0, p = np.ones((3, 1)), np.zeros((3, 1))
    for saveat, inner in ((25.0, [25.0, 50.0, 75.0]), ([25.0, 50.0, 75.0], [25.0, 50.0, 75.0])):
        sol, ref = solve_both("constant_1delay_long", O.MODEL_1DELAY_LONG, u0, p, [0.2], (0.0, 100.0), saveat,
                              save_start=save_start, save_end=save_end)
        assert list(sol.t) == ([0.0] if save_start else []) + inner + ([100.0] if save_end else [])
        assert_same_ensemble(sol, ref)
    prob = B.DDEProblem("constant_1delay_long", [1.0], None, (0.0, 100.0), [0.0], constant_lags=[0.2])
    assert list(B.solve(prob, B.MethodOfSteps(B.Tsit5()), saveat=40.0).t) == [0.0, 40.0, 80.0, 100.0]   # saveat.jl:134-136


def test_failed_trajectories_do_not_poison_their_warp():
    """maxiters small enough that most trajectories stop with MaxIters: retcodes, NaN padding and the surviving
    trajectories must all match the oracle (SURVEY §5 failure detection)."""
    u0, p = bc_sweep(4096, seed=5)
    sol, ref = solve_both("bc_model", O.MODEL_BC, u0, p, [1.0], (0.0, 10.0), 0.1, maxiters=38)
    assert_same_ensemble(sol, ref)
    assert 0 < int((sol.retcodes == 1).sum()) < 4096 and int((sol.retcodes == 2).sum()) > 0
    assert np.isnan(sol.array[sol.retcodes == 2][:, -1, :]).all()


def test_empty_and_ragged_sizes():
    prob = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])
    for n in (1, 31, 33, 1000):     # not multiples of the 32-lane tile
        u0, p = bc_sweep(n, seed=n)
        sol, ref = solve_both("bc_model", O.MODEL_BC, u0, p, [1.0], (0.0, 10.0), 1.0)
        assert_same_ensemble(sol, ref)
    s = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    t, u, rc, st = s.solve_host(np.empty((0, 3)), np.empty((0, 3)), [1.0], (0.0, 10.0), B.saveat_grid(1.0, (0.0, 10.0)), True, True)
    assert u.shape == (0, 11, 3) and len(t) == 11
    s.close()


# ---------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json sizes; the oracle cannot follow, so size-independent properties)
# ---------------------------------------------------------------------------------------------------------
def test_full_size_bc_properties():
    n = 1_000_000
    i = np.arange(n)
    a, b, c = i // 10000, (i // 100) % 100, i % 100
    p = np.ascontiguousarray(np.stack([0.1 + 0.2 * a / 99, 0.2 + 0.2 * b / 99, 0.5 + 1.0 * c / 99], axis=1))
    u0 = np.ones((n, 3))
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    t, u, rc, st = solver.solve_host(u0, p, [1.0], (0.0, 10.0), grid, True, True)
    assert u.shape == (n, 101, 3) and (rc == 1).all() and np.isfinite(u).all()
    assert np.array_equal(u[:, 0, :], u0)                              # save_start row is u0
    # idempotence + wave-size invariance: a different chunking must give the identical block
    solver2 = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {"chunk_traj": 77777}))
    _, u2, rc2, st2 = solver2.solve_host(u0, p, [1.0], (0.0, 10.0), grid, True, True)
    assert np.array_equal(u, u2) and np.array_equal(st[:, :6], st2[:, :6])
    solver2.close()
    # ... and so must another number of waves in flight (B2D_SLOTS is read when the handle is created; default 3)
    for slots in ("2", "4"):
        os.environ["B2D_SLOTS"] = slots
        try:
            solver3 = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
        finally:
            del os.environ["B2D_SLOTS"]
        _, u3, rc3, st3 = solver3.solve_host(u0, p, [1.0], (0.0, 10.0), grid, True, True)
        assert np.array_equal(u, u3) and np.array_equal(rc, rc3) and np.array_equal(st[:, :6], st3[:, :6])
        solver3.close()
    # a strided sample against the oracle (bit-exact) and the checksum of its statistics
    idx = np.arange(0, n, 997)
    ref = O.solve(O.default_config(model=O.MODEL_BC), u0[idx], p[idx], [1.0], 0.0, 10.0, saveat=grid, nthreads=O.hardware_threads())
    assert np.array_equal(u[idx], ref["u"]) and np.array_equal(st[idx, :6], ref["stats"][:, :6])
    # device-resident entry point gives the same block as the host-buffer entry point
    import torch
    d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
    d_out = torch.empty((n, 101, 3), dtype=torch.float64, device="cuda")
    d_rc = torch.empty(n, dtype=torch.int32, device="cuda")
    d_st = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), [1.0], (0.0, 10.0), grid, True, True, d_out.data_ptr(), d_rc.data_ptr(),
                        d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), u)
    # performance canary (not a benchmark): this launch takes ~5.6 ms on a B200; a state vector that silently moved
    # into local memory once made it 18.8 ms with every parity test still green
    ms, _, _ = solver.last_timing()
    assert ms < 12.0, f"bc_model 10^6 kernel took {ms:.1f} ms (expected ~5.6 ms)"
    solver.close()
    solver2.close()


def test_single_process_multi_handle_sharding():
    """b2d_solve_multi: two handles (both on device 0 here; one per GPU on a multi-GPU box) must give the bitwise
    identical block as one handle, including an odd split."""
    u0, p = bc_sweep(10001, seed=11)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    one = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=10001, saveat=0.5)
    two = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(devices=[0, 0, 0]), trajectories=10001, saveat=0.5)
    assert np.array_equal(one.t, two.t) and np.array_equal(one.array, two.array)
    assert np.array_equal(one.retcodes, two.retcodes) and np.array_equal(one.stats[:, :6], two.stats[:, :6])


def test_branch_free_division_is_ieee():
    """csrc/detmath.h dm_div_try is the compiler's own double-division sequence with the range test returned as a flag
    instead of branched on (models.cuh hmap, the interval window).  Wherever the flag is true the quotient must equal
    `a / b` in every bit: 2^28 operand pairs per class."""
    import ctypes as C
    for cls in range(4):
        bad, flagged = C.c_uint64(), C.c_uint64()
        rc = B.lib().b2d_check_division(0, cls, 1 << 28, 0xB200DDE + cls, C.byref(bad), C.byref(flagged))
        assert rc == 0 and bad.value == 0, (cls, bad.value)
        if cls in (1, 2):
            assert flagged.value == 0, (cls, flagged.value)    # ordinary magnitudes never leave the fast sequence
        else:
            assert 0 < flagged.value < (1 << 28)


def test_values_outside_the_fast_division_range_are_solved_exactly():
    """The lean Rosenbrock23 kernels divide with the compiler's own sequence minus its branch to a slow path; an operand
    outside the range of that sequence (here: states of magnitude 1e-300, whose error residuals are denormal-scale numerators)
    flags the trajectory, b2d_solve solves it again in the general kernel, and the result is the oracle's, bit for bit.  The
    device entry point, which cannot re-solve, reports B2D_RC_INEXACT_DIVISION instead of returning an unchecked result."""
    n = 512
    u0 = np.full((n, 1), 1.0)
    u0[::4, 0] = 1e-300           # every fourth trajectory leaves the fast range
    p = np.zeros((n, 1))
    prob = B.DDEProblem("constant_1delay", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    alg = B.MethodOfSteps(B.Rosenbrock23())
    sol = B.solve(ens, alg, B.EnsembleB200(0), trajectories=n, saveat=1.0, abstol=0.0)
    ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_1DELAY, abstol=0.0), u0, p, [1.0], 0.0, 10.0,
                  saveat=B.saveat_grid(1.0, (0.0, 10.0)), nthreads=O.hardware_threads())
    assert_same_ensemble(sol, ref)
    solver = B.Solver(B.api._make_config(prob, alg, 0, {"abstol": 0.0}))
    _, _, rc, _ = _device_solve(solver, u0, p, [1.0], (0.0, 10.0), 1.0)
    solver.close()
    assert set(np.unique(rc[::4])) == {9} and set(np.unique(np.delete(rc, np.arange(0, n, 4)))) == {1}


def test_many_user_discontinuities_inside_one_lag():
    """Every user discontinuity starts a chain t, t + lag, ... of its own; with one constant lag each chain keeps one
    pending entry in the propagated heaps.  The reference's BinaryMinHeap is unbounded (src/solve.jl:683-716,
    add_next_discontinuities! src/integrators/utils.jl:244-283): six discontinuities inside one lag window must simply work
    (the general kernels once had room for two), and more chains than the queue holds are refused, never silent."""
    discs = [(0.1, 0), (0.2, 1), (0.3, 0), (0.45, 2), (0.7, 0), (0.95, 1)]
    tstops = [0.15, 0.5]
    u0, p = bc_sweep(2000, seed=21)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=2000, saveat=0.1, tstops=tstops, d_discontinuities=discs)
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    try:
        O.set_stops(tstops, discs)
        ref = O.solve(O.default_config(model=O.MODEL_BC), u0, p, [1.0], 0.0, 10.0, saveat=grid, nthreads=O.hardware_threads())
        o = oracle_everystep("bc_model")
        o1 = oracle_everystep("one_delay")
    finally:
        O.set_stops()
    assert sol.converged
    assert_same_ensemble(sol, ref)
    s = gpu_everystep("bc_model", tstops=tstops, d_discontinuities=discs)
    assert_same_solution(s, o)
    assert len(s.tracked_discontinuities) > 20                      # t0 and six user chains, each followed through the lag
    assert_same_solution(gpu_everystep("one_delay", tstops=tstops, d_discontinuities=discs), o1)
    with pytest.raises(B.B200DDEError, match="d_discontinuities"):
        gpu_everystep("bc_model", d_discontinuities=[(0.01 * k, 0) for k in range(1, 41)])


def test_full_size_mackey_glass_properties():
    """BASELINE configs[2] at one GPU's share: 10^6 trajectories of the C3 sweep (beta x h0 grid, tau = 17, tspan 0..1000,
    saveat 10) through b2d_solve.  All Success, a strided sample bit-exact against the oracle including the statistics, and
    the same at 1e-8 / 1e-8 on 10^5 trajectories (27-node histories: the ring regrowth path).  The reference only *runs*
    its Mackey-Glass problem (test/integrators/verner.jl:74-88); parity here is parity with the KAT-pinned oracle."""
    for n, tol, stride in ((1_000_000, {}, 1999), (100_000, {"abstol": 1e-8, "reltol": 1e-8}, 1999)):
        i = np.arange(n)
        na = n // 4000
        p = np.ascontiguousarray(np.stack([0.18 + 0.04 * (i // 4000) / max(na - 1, 1), 0.5 + 1.0 * (i % 4000) / 3999], axis=1))
        u0 = p[:, 1:2].copy()
        prob = B.DDEProblem("mackey_glass", u0[0], None, (0.0, 1000.0), p[0], constant_lags=[17.0])
        solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, dict(tol)))
        grid = B.saveat_grid(10.0, (0.0, 1000.0))
        t, u, rc, st = solver.solve_host(u0, p, [17.0], (0.0, 1000.0), grid, True, True)
        assert u.shape == (n, 101, 1) and (rc == 1).all() and np.isfinite(u).all()
        assert np.array_equal(u[:, 0, :], u0) and t[0] == 0.0 and t[-1] == 1000.0
        idx = np.arange(0, n, stride)
        ref = O.solve(O.default_config(model=O.MODEL_MACKEY_GLASS, **tol), u0[idx], p[idx], [17.0], 0.0, 1000.0, saveat=grid,
                      nthreads=O.hardware_threads())
        assert np.array_equal(u[idx], ref["u"]) and np.array_equal(st[idx, :6], ref["stats"][:, :6])
        assert np.array_equal(rc[idx], ref["retcode"])
        # the device-resident entry point gives the same block; its single launch is the timing canary
        import torch
        d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
        d_out = torch.empty((n, 101, 1), dtype=torch.float64, device="cuda")
        d_rc = torch.empty(n, dtype=torch.int32, device="cuda")
        d_st = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        for _ in range(2):
            solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), [17.0], (0.0, 1000.0), grid, True, True, d_out.data_ptr(),
                                d_rc.data_ptr(), d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), u) and np.array_equal(d_st.cpu().numpy()[:, :6], st[:, :6])
        ms, _, _ = solver.last_timing()
        if not tol:
            assert ms < 30.0, f"Mackey-Glass 10^6 kernel took {ms:.1f} ms (expected < 20 ms)"
        solver.close()


def _splitmix64_permutation(n, seed=0xB200DDE):
    """the shuffled variant of the sweeps — a permutation drawn from splitmix64(seed = 0xB200DDE) keys."""
    x = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)).astype(np.uint64)
    x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return np.argsort(x, kind="stable")


@pytest.mark.parametrize("which", ["bc_model", "mackey_glass"])
def test_full_size_ensembles_are_permutation_equivariant(which):
    """Trajectories are independent, so solving the ensemble in a shuffled order must give the shuffled result, bit for bit:
    nothing a trajectory computes may depend on its neighbours in the warp (tile scheduling, ring slots, windows, the lanes a
    rejected step idles).  10^6 trajectories of the C2 / C3 sweeps in grid order and in splitmix64 order (the
    shuffled variant stresses divergence — neighbouring lanes no longer take similar steps); the kernel times of both
    orders are printed (pytest -s)."""
    import torch
    n = 1_000_000
    i = np.arange(n)
    if which == "bc_model":
        a, b, c = i // 10000, (i // 100) % 100, i % 100
        p = np.ascontiguousarray(np.stack([0.1 + 0.2 * a / 99, 0.2 + 0.2 * b / 99, 0.5 + 1.0 * c / 99], axis=1))
        u0, lags, tspan, dt_save = np.ones((n, 3)), [1.0], (0.0, 10.0), 0.1
    else:
        p = np.ascontiguousarray(np.stack([0.18 + 0.04 * (i // 4000) / 249, 0.5 + 1.0 * (i % 4000) / 3999], axis=1))
        u0, lags, tspan, dt_save = p[:, 1:2].copy(), [17.0], (0.0, 1000.0), 10.0
    perm = _splitmix64_permutation(n)
    assert np.array_equal(np.sort(perm), i)
    prob = B.DDEProblem(which, u0[0], None, tspan, p[0], constant_lags=lags)
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    grid = B.saveat_grid(dt_save, tspan)
    res = []
    for order in (i, perm):
        d_u0, d_p = torch.from_numpy(np.ascontiguousarray(u0[order])).cuda(), torch.from_numpy(np.ascontiguousarray(p[order])).cuda()
        d_out = torch.empty((n, len(grid) + 2, u0.shape[1]), dtype=torch.float64, device="cuda")
        d_rc = torch.empty(n, dtype=torch.int32, device="cuda")
        d_st = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        for _ in range(2):
            solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), lags, tspan, grid, True, True, d_out.data_ptr(), d_rc.data_ptr(),
                                d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res.append((d_out.cpu().numpy(), d_rc.cpu().numpy(), d_st.cpu().numpy(), solver.last_timing()[0]))
    solver.close()
    (u_a, rc_a, st_a, ms_a), (u_b, rc_b, st_b, ms_b) = res
    assert (rc_a == 1).all()
    assert np.array_equal(u_b, u_a[perm]) and np.array_equal(rc_b, rc_a[perm]) and np.array_equal(st_b[:, :6], st_a[perm][:, :6])
    print(f"{which}: grid order {ms_a:.2f} ms, splitmix64 order {ms_b:.2f} ms")


def test_multi_gpu_sharding_is_bitwise_equal_to_one_gpu():
    """the G-GPU result is bitwise equal to the 1-GPU result.  Uses every GPU the process can see
    (b2d_solve_multi, one handle per device); on a one-GPU box the test is skipped — tests/test_distributed_gloo.py covers
    the sharding arithmetic on CPU and test_single_process_multi_handle_sharding the multi-handle path on one device."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs more than one visible GPU")
    devices = list(range(g))
    for model, oid, sweep, lags, tspan, saveat, n in (("bc_model", O.MODEL_BC, bc_sweep, [1.0], (0.0, 10.0), 0.1, 200_003),
                                                       ("mackey_glass", O.MODEL_MACKEY_GLASS, mg_sweep, [17.0], (0.0, 1000.0), 10.0, 100_001)):
        u0, p = sweep(n, seed=31)
        prob = B.DDEProblem(model, u0[0], None, tspan, p[0], constant_lags=lags)
        ens = B.EnsembleProblem(prob, u0=u0, p=p)
        one = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=n, saveat=saveat)
        many = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(devices=devices), trajectories=n, saveat=saveat)
        last = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(devices[-1]), trajectories=n, saveat=saveat)
        for other in (many, last):
            assert np.array_equal(one.t, other.t) and np.array_equal(one.array, other.array)
            assert np.array_equal(one.retcodes, other.retcodes) and np.array_equal(one.stats[:, :6], other.stats[:, :6])
        idx = np.arange(0, n, 1999)
        ref = O.solve(O.default_config(model=oid), u0[idx], p[idx], lags, tspan[0], tspan[1], saveat=B.saveat_grid(saveat, tspan),
                      nthreads=O.hardware_threads())
        assert np.array_equal(many.array[idx], ref["u"]) and np.array_equal(many.stats[idx, :6], ref["stats"][:, :6])


def test_library_gather_single_rank():
    """b2d_comm_init / b2d_gather_saveat with a communicator of one rank (what a 1-GPU box can run): NCCL is loaded at run
    time, the root's own block is copied to its place in the gathered array.  Ranks > 1: bench.py under torchrun reports
    `gather_ok` (per-shard checksums on the root against the senders'), tests/test_distributed_gloo.py the set-up logic."""
    import torch
    from b200dde.api import comm_unique_id
    u0, p = bc_sweep(4096, seed=5)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    out = _device_solve(solver, u0, p, [1.0], (0.0, 10.0), 0.5)
    d_local = torch.from_numpy(out[1]).cuda()
    solver.comm_init(1, 0, comm_unique_id())
    d_root = torch.zeros_like(d_local)
    solver.gather_saveat(d_local.data_ptr(), [d_local.numel()], 0, d_root.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(d_root, d_local)
    with pytest.raises(B.B200DDEError):
        solver.gather_saveat(d_local.data_ptr(), [d_local.numel()], 3, d_root.data_ptr())
    solver.close()


def test_discrete_callbacks_bit_exact_and_reference_kats():
    """solve(...; callback = DiscreteCallback(...)) (test/integrators/events.jl: terminate! at a tstop, a jump that is saved
    before and after, a periodic callback): the general kernels against the oracle, bit for bit, and the KAT assertions on
    the GPU result.  Ensembles with a saveat grid: save_positions = (false, false) and the default."""
    cases = [("one_delay", B.PresetTimeCallback(4.0, "terminate"), dict(callback=2, cb_affect=1, cb_par0=4.0), [4.0]),
             ("one_delay", B.PresetTimeCallback(2.6, "negate"), dict(callback=2, cb_affect=3, cb_par0=2.6), [2.6]),
             ("one_delay", B.PeriodicCallback(None, 1.0), dict(callback=3, cb_affect=0, cb_par0=1.0), [float(k) for k in range(1, 11)]),
             ("bc_model", B.PresetTimeCallback(3.3, ("add", 1, 0.25)), dict(callback=2, cb_affect=2, cb_idx=1, cb_par0=3.3, cb_par1=0.25), [3.3]),
             ("state_dependent", B.PresetTimeCallback(2.0, "negate"), dict(callback=2, cb_affect=3, cb_par0=2.0), [2.0])]
    for case, cb, okw, stops in cases:
        s = gpu_everystep(case, callback=cb)
        try:
            O.set_stops(stops, [])
            o = oracle_everystep(case, **okw)
        finally:
            O.set_stops()
        assert_same_solution(s, o)
    s = gpu_everystep("one_delay", callback=B.PresetTimeCallback(4.0, "terminate"))
    assert s.t[-1] == 4.0 and s.retcode == "Terminated"                      # events.jl:52-55
    s = gpu_everystep("one_delay", callback=B.PresetTimeCallback(2.6, "negate"))
    i = [k for k, x in enumerate(s.t) if x == 2.6]
    assert len(i) == 2 and s.u[i[0], 0] == -s.u[i[1], 0]                     # events.jl:16-18
    s = gpu_everystep("one_delay", callback=B.PeriodicCallback(None, 1.0))
    assert s.retcode == "Success" and sum(1 for tt, oo in s.tracked_discontinuities if oo == 0 and tt > 0) >= 9   # events.jl:93-96
    # ensemble with a saveat grid
    u0, p = bc_sweep(3000, seed=41)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    cb = B.PresetTimeCallback(3.3, ("add", 1, 0.25), save_positions=(False, False))
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, callback=cb)
    try:
        O.set_stops([3.3], [])
        ref = O.solve(O.default_config(model=O.MODEL_BC, callback=2, cb_affect=2, cb_idx=1, cb_par0=3.3, cb_par1=0.25, cb_save_positions=0),
                      u0, p, [1.0], 0.0, 10.0, saveat=B.saveat_grid(0.1, (0.0, 10.0)), nthreads=O.hardware_threads())
    finally:
        O.set_stops()
    assert sol.converged
    assert_same_ensemble(sol, ref)
    base = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1)
    assert not np.array_equal(base.array[:, 40:], sol.array[:, 40:]) and np.array_equal(base.array[:, :28], sol.array[:, :28])
    # default save_positions = (true, true) with a saveat grid: the forced saves enter the history (the jump is seen by later
    # lookups exactly as on the reference), the two extra points at the event are not part of the fixed-shape block
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, callback=B.PresetTimeCallback(3.3, "negate"))
    try:
        O.set_stops([3.3], [])
        ref = O.solve(O.default_config(model=O.MODEL_BC, callback=2, cb_affect=3, cb_par0=3.3),
                      u0, p, [1.0], 0.0, 10.0, saveat=B.saveat_grid(0.1, (0.0, 10.0)), nthreads=O.hardware_threads())
    finally:
        O.set_stops()
    assert sol.converged
    assert_same_ensemble(sol, ref)


def test_continuous_callbacks_bit_exact_and_reference_kats():
    """solve(...; callback = ContinuousCallback(condition, affect!)) (test/integrators/events.jl:9-36): the event is located on
    the step's dense output, the step is cut there and its slopes recomputed, affect! runs, an order-0 discontinuity follows.
    General kernels against the oracle, bit for bit (save_everystep and saveat), and the KAT assertions on the GPU result."""
    cases = [("one_delay", B.ContinuousCallback(("time", 2.6), "negate"), dict(callback=5, cb_affect=3, cb_par0=2.6), {}),
             ("one_delay", B.ContinuousCallback(("time", 2.6), "negate"), dict(callback=5, cb_affect=3, cb_par0=2.6), dict(dtmax=0.01)),
             ("one_delay", B.ContinuousCallback(("state", 0, 0.0), "terminate"), dict(callback=6, cb_affect=1, cb_cond_idx=0, cb_par0=0.0), {}),
             ("one_delay", B.ContinuousCallback(("state", 0, 0.0), "terminate", direction=1),
              dict(callback=6, cb_affect=1, cb_cond_idx=0, cb_par0=0.0, cb_direction=1), {}),
             ("one_delay", B.ContinuousCallback(("state", 0, -0.2), ("add", 0, 0.5), interp_points=0),
              dict(callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=-0.2, cb_par1=0.5, cb_interp_points=0), {}),
             ("bc_model", B.ContinuousCallback(("state", 2, 0.8), ("add", 1, 0.25)),
              dict(callback=6, cb_affect=2, cb_idx=1, cb_cond_idx=2, cb_par0=0.8, cb_par1=0.25), {}),
             ("bc_model", B.ContinuousCallback(("state", 1, 0.9), "negate", direction=-1),
              dict(callback=6, cb_affect=3, cb_cond_idx=1, cb_par0=0.9, cb_direction=-1), {}),
             ("two_delays_long", B.ContinuousCallback(("state", 0, 0.1), "negate"), dict(callback=6, cb_affect=3, cb_cond_idx=0, cb_par0=0.1), {}),
             ("state_dependent", B.ContinuousCallback(("state", 0, 0.3), ("add", 0, 0.4)),
              dict(callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=0.3, cb_par1=0.4), {}),
             ("backwards", B.ContinuousCallback(("time", 1.3), "negate"), dict(callback=5, cb_affect=3, cb_par0=1.3), {}),   # reverse time:
             ("backwards", B.ContinuousCallback(("state", 0, 0.8), ("add", 0, 0.3)),                                         # "left" is later
              dict(callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=0.8, cb_par1=0.3), {}),
             ("logistic_p14", B.ContinuousCallback(("state", 0, 1.4), ("add", 0, -0.2), direction=1),
              dict(callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=1.4, cb_par1=-0.2, cb_direction=1), {}),
             ("mackey_glass", B.ContinuousCallback(("state", 0, 1.0), ("add", 0, -0.3), direction=1),
              dict(callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=1.0, cb_par1=-0.3, cb_direction=1), {})]
    n_events = 0
    for case, cb, okw, kw in cases:
        s = gpu_everystep(case, callback=cb, **kw)
        o = oracle_everystep(case, **okw, **kw)
        if len(o.tracked) > 64:   # (150 events on Mackey-Glass: the solve_everystep call returns the first 64 tracked entries)
            assert s.tracked_discontinuities == o.tracked[:64]
            o.tracked = o.tracked[:64]
        assert_same_solution(s, o)
        n_events += sum(1 for tt, oo in s.tracked_discontinuities if oo == 0 and tt > s.t[0])
    assert n_events >= 10                                                                # the cases do fire
    # BS3 (Hermite dense output)
    s = gpu_everystep("one_delay", alg=B.MethodOfSteps(B.BS3()), callback=B.ContinuousCallback(("time", 2.6), "negate"))
    o = oracle_everystep("one_delay", alg=O.ALG_BS3, callback=5, cb_affect=3, cb_par0=2.6)
    assert_same_solution(s, o)
    # events.jl:9-24 on the GPU result
    s1 = gpu_everystep("one_delay", callback=B.ContinuousCallback(("time", 2.6), "negate"))
    ts = [k for k, x in enumerate(s1.t) if np.isclose(x, 2.6, rtol=1.5e-8, atol=0)]
    assert len(ts) == 2 and s1.u[ts[0], 0] == -s1.u[ts[1], 0] != 0.0
    assert abs(s1(np.nextafter(s1.t[ts[0]], 10.0))[0] + s1(np.nextafter(s1.t[ts[0]], 0.0))[0]) < 2.0e-5
    # ensemble with a saveat grid: u0 / p sweep, every trajectory has its own event times
    u0, p = bc_sweep(3000, seed=43)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    cb = B.ContinuousCallback(("state", 2, 0.8), ("add", 1, 0.25))
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, callback=cb)
    ref = O.solve(O.default_config(model=O.MODEL_BC, callback=6, cb_affect=2, cb_idx=1, cb_cond_idx=2, cb_par0=0.8, cb_par1=0.25),
                  u0, p, [1.0], 0.0, 10.0, saveat=B.saveat_grid(0.1, (0.0, 10.0)), nthreads=O.hardware_threads())
    assert_same_ensemble(sol, ref)
    base = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1)
    assert (np.abs(base.array - sol.array).max(axis=(1, 2)) > 1e-3).mean() > 0.5        # most trajectories had an event
    with pytest.raises(B.B200DDEError, match="ContinuousCallback"):
        B.solve(B.DDEProblem("stiff_kinetics", [2.0, 0.5, 1.0], None, (0.0, 1.0), [100.0], constant_lags=[1.0]),
                B.MethodOfSteps(B.Rosenbrock23()), callback=B.ContinuousCallback(("time", 0.5), "negate"))


def test_per_trajectory_constant_lags():
    """A sweep over the delay itself: the reference ensemble solves what prob_func returns as a full problem per trajectory
    (src/solve.jl:159 unpacks constant_lags), so `remake(prob; constant_lags = [tau_i])` is an ordinary ensemble.  Every
    trajectory against the oracle solving it with its own lag, bit for bit; Mackey-Glass (tau 14 .. 20), two lags at once,
    the prob_func form, save_everystep, and the same ensemble over two handles."""
    rng = np.random.default_rng(77)
    n = 600
    p = np.stack([0.18 + 0.04 * rng.random(n), 0.5 + 1.0 * rng.random(n)], axis=1)
    u0 = p[:, 1:2].copy()
    tau = 14.0 + 6.0 * rng.random((n, 1))
    prob = B.DDEProblem("mackey_glass", u0[0], None, (0.0, 300.0), p[0], constant_lags=[17.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p, constant_lags=tau)
    alg = B.MethodOfSteps(B.Tsit5())
    sol = B.solve(ens, alg, B.EnsembleB200(0), trajectories=n, saveat=10.0)
    grid = B.saveat_grid(10.0, (0.0, 300.0))
    ref = O.solve(O.default_config(model=O.MODEL_MACKEY_GLASS), u0, p, tau, 0.0, 300.0, saveat=grid)
    assert sol.converged
    assert_same_ensemble(sol, ref)
    same = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), alg, B.EnsembleB200(0), trajectories=n, saveat=10.0)
    assert (np.abs(same.array - sol.array).max(axis=(1, 2)) > 1e-3).mean() > 0.9      # the lags do matter
    # the same ensemble sharded over two handles (every handle is given the whole lag matrix)
    two = B.solve(ens, alg, B.EnsembleB200(devices=[0, 0]), trajectories=n, saveat=10.0)
    assert np.array_equal(two.array, sol.array) and np.array_equal(two.stats, sol.stats)
    # prob_func form, two lags, fixed-point iterations (steps longer than the lags)
    m = 200
    l2 = np.stack([1 / 3 + 0.1 * rng.random(m), 1 / 5 + 0.05 * rng.random(m)], axis=1)
    prob2 = B.DDEProblem("constant_2delays", [1.0], None, (0.0, 20.0), [0.0], constant_lags=[1 / 3, 1 / 5])
    ens2 = B.EnsembleProblem(prob2, prob_func=lambda pr, i, rep: pr.remake(constant_lags=l2[i]))
    sol2 = B.solve(ens2, alg, B.EnsembleB200(0), trajectories=m, saveat=0.5)
    ref2 = O.solve(O.default_config(model=O.MODEL_2DELAYS), np.ones((m, 1)), np.zeros((m, 1)), l2, 0.0, 20.0,
                   saveat=B.saveat_grid(0.5, (0.0, 20.0)))
    assert_same_ensemble(sol2, ref2)
    assert ref2["stats"][:, 3].sum() > 0                                               # nfpiter: the fixed point did run
    # save_everystep
    sols = B.solve(B.EnsembleProblem(prob2, u0=np.ones((3, 1)), p=np.zeros((3, 1)), constant_lags=l2[:3]), alg, B.EnsembleB200(0), trajectories=3)
    for i in range(3):
        o = O.solve_everystep(O.default_config(model=O.MODEL_2DELAYS), [1.0], [0.0], l2[i], 0.0, 20.0)
        assert_same_solution(sols[i], o)
    # a lag against the direction of time is an error (test/interface/backwards.jl:21-24), whichever trajectory has it
    bad = tau.copy()
    bad[17, 0] = -1.0
    with pytest.raises(B.B200DDEError, match="wrong sign"):
        B.solve(B.EnsembleProblem(prob, u0=u0, p=p, constant_lags=bad), alg, B.EnsembleB200(0), trajectories=n, saveat=10.0)
    # the handle forgets the matrix when told to
    solver = B.Solver(B.api._make_config(prob, alg, 0, {}))
    solver.set_trajectory_lags(tau)
    solver.set_trajectory_lags(None)
    t, u, rc, st = solver.solve_host(u0, p, [17.0], (0.0, 300.0), grid, True, True)
    solver.close()
    assert np.array_equal(u, same.array)


def test_linearity_property_one_delay():
    """u' = -u(t-1) is linear; with reltol-only error control (abstol = 0, h = 0) the whole adaptive solve is
    scale-equivariant, so scaling u0 by a power of two scales the solution exactly."""
    n = 64
    u0 = np.ones((n, 1))
    u0[1::2] = 4.0
    p = np.zeros((n, 1))
    prob = B.DDEProblem("constant_1delay", [1.0], None, (0.0, 10.0), [0.0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=n, saveat=0.5, abstol=0.0)
    assert np.array_equal(sol.array[1::2], 4.0 * sol.array[0::2])


def test_full_size_state_dependent_and_stiff_properties():
    """BASELINE configs[3] and [4] at 10^6 trajectories: all Success, finite, strided sample bit-exact vs the oracle."""
    n = 1_000_000
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    idx = np.arange(0, n, 1999)
    # C4: state-dependent lag, initial-condition sweep
    u0 = (0.5 + 1.0 * np.arange(n) / (n - 1))[:, None].copy()
    p = np.zeros((n, 1))
    prob = B.DDEProblem("state_dependent", u0[0], None, (0.0, 10.0), p[0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    t, u, rc, st = solver.solve_host(u0, p, [], (0.0, 10.0), grid, True, True)
    solver.close()
    assert (rc == 1).all() and np.isfinite(u).all()
    ref = O.solve(O.default_config(model=O.MODEL_STATE_DEP), u0[idx], p[idx], [], 0.0, 10.0, saveat=grid, nthreads=O.hardware_threads())
    assert np.array_equal(u[idx], ref["u"]) and np.array_equal(st[idx, :6], ref["stats"][:, :6])
    # C5: stiff kinetics under Rosenbrock23, lambda sweep 1e2..1e5
    lam = 10.0 ** (2.0 + 3.0 * np.arange(n) / (n - 1))
    u0 = np.tile([2.0, 0.5, 1.0], (n, 1))
    p = lam[:, None].copy()
    prob = B.DDEProblem("stiff_kinetics", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Rosenbrock23()), 0, {}))
    t, u, rc, st = solver.solve_host(u0, p, [1.0], (0.0, 10.0), grid, True, True)
    ms, nl, cap = solver.last_timing()
    solver.close()
    assert (rc == 1).all() and np.isfinite(u).all()
    ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_STIFF), u0[idx], p[idx], [1.0], 0.0, 10.0, saveat=grid,
                  nthreads=O.hardware_threads())
    assert np.array_equal(u[idx], ref["u"]) and np.array_equal(st[idx, :6], ref["stats"][:, :6])


# ---------------------------------------------------------------------------------------------------------
# randomized differential tests: many trajectories x several tolerances per model, everything bit-exact
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("abstol,reltol", [(1e-6, 1e-3), (1e-4, 1e-2), (1e-9, 1e-7)])
def test_randomized_fixed_point_heavy_ensemble(abstol, reltol):
    """2 constant lags, order-0 discontinuity at t0, steps larger than the lags: fixed-point iterations, forced step
    failures, local-memory discontinuity queues — per-trajectory initial values so every lane walks its own path."""
    rng = np.random.default_rng(123)
    n = 1500
    u0 = (0.2 + 2.0 * rng.random(n))[:, None]
    p = np.zeros((n, 1))
    sol, ref = solve_both("constant_2delays", O.MODEL_2DELAYS, u0, p, [1 / 3, 1 / 5], (0.0, 30.0), 0.5, abstol=abstol, reltol=reltol)
    assert_same_ensemble(sol, ref)
    assert int(sol.stats[:, 3].sum()) > 0          # fixed-point iterations really happened


@pytest.mark.parametrize("abstol,reltol", [(1e-6, 1e-3), (1e-10, 1e-8)])
def test_randomized_state_dependent_ensemble(abstol, reltol):
    rng = np.random.default_rng(321)
    n = 3000
    u0 = (0.1 + 3.0 * rng.random(n))[:, None]
    p = np.zeros((n, 1))
    sol, ref = solve_both("state_dependent", O.MODEL_STATE_DEP, u0, p, [], (0.0, 10.0), 0.25, abstol=abstol, reltol=reltol)
    assert_same_ensemble(sol, ref)


def test_randomized_bc_tolerance_sweep():
    for k, (abstol, reltol) in enumerate([(1e-5, 1e-2), (1e-8, 1e-6), (1e-10, 1e-10)]):
        u0, p = bc_sweep(3000, seed=100 + k)
        u0 = u0 * (0.5 + np.random.default_rng(k).random((3000, 3)))
        sol, ref = solve_both("bc_model", O.MODEL_BC, u0, p, [1.0], (0.0, 10.0), 0.1, abstol=abstol, reltol=reltol)
        assert_same_ensemble(sol, ref)


def test_randomized_logistic_ensemble():
    """parameters.jl's delayed logistic over a parameter sweep that crosses its Hopf bifurcation (p = pi/2)."""
    n = 4000
    p = (0.2 + 2.3 * np.arange(n) / (n - 1))[:, None]
    u0 = np.full((n, 1), 0.1)
    sol, ref = solve_both("delayed_logistic", O.MODEL_LOGISTIC, u0, p, [1.0], (0.0, 50.0), 1.0)
    assert_same_ensemble(sol, ref)


def test_per_component_tolerances_and_save_idxs():
    """abstol / reltol as arrays (src/utils.jl:91-133) and save_idxs (src/solve.jl:47,217-218) — SURVEY §8 f1."""
    u0, p = bc_sweep(2000, seed=77)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    abstol, reltol = [1e-9, 1e-6, 1e-7], [1e-6, 1e-3, 1e-4]
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    try:
        O.set_options(abstol, reltol, [2, 0])
        ref = O.solve(O.default_config(model=O.MODEL_BC), u0, p, [1.0], 0.0, 10.0, saveat=grid, nthreads=O.hardware_threads())
    finally:
        O.set_options()
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=2000, saveat=0.1, abstol=abstol, reltol=reltol,
                  save_idxs=[2, 0])
    assert sol.array.shape == (2000, 101, 2)
    assert_same_ensemble(sol, ref)
    # and the vector tolerances really changed the integration
    base = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=2000, saveat=0.1)
    assert int(sol.stats[:, 0].sum()) > int(base.stats[:, 0].sum())
    prob1 = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])
    with pytest.raises(B.B200DDEError, match="out of range"):
        B.solve(prob1, B.MethodOfSteps(B.Tsit5()), saveat=1.0, save_idxs=[3])


# ---- NLAnderson (src/fpsolve/functional.jl:15-75; KAT test/interface/fpsolve.jl:59-75) ------------------------
@pytest.mark.parametrize("case", ["two_delays_long", "one_delay_long", "logistic_p14"])
def test_anderson_everystep_bit_exact(case):
    alg = B.MethodOfSteps(B.Tsit5(), fpsolve=B.NLAnderson())
    s, o = gpu_everystep(case, alg), oracle_everystep(case, fp_alg=1)
    assert_same_solution(s, o)
    assert s.stats.nsolve == o.stats[O.STAT_NSOLVE]
    if case != "logistic_p14":
        assert s.stats.nsolve > 0


def test_anderson_bs3_everystep_bit_exact():
    alg = B.MethodOfSteps(B.BS3(), fpsolve=B.NLAnderson(max_iter=8))
    s = gpu_everystep("two_delays_long", alg)
    o = oracle_everystep("two_delays_long", fp_alg=1, fp_max_iter=8, alg=O.ALG_BS3)
    assert_same_solution(s, o)
    assert s.stats.nsolve == o.stats[O.STAT_NSOLVE] > 0


@pytest.mark.parametrize("reset", [0, 1])
def test_anderson_system_ensemble_bit_exact(reset):
    """n_state = 3 (history window 3: Givens downdates, multi-column least squares) on a sweep of the coupling strength,
    saveat output, both history policies."""
    n = 3000
    rng = np.random.default_rng(5)
    p = (30.0 + 270.0 * rng.random(n))[:, None]
    u0 = np.tile([2.0, 0.5, 1.0], (n, 1))
    prob = B.DDEProblem("stiff_kinetics", u0[0], None, (0.0, 2.0), p[0], constant_lags=[0.005])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5(), fpsolve=B.NLAnderson()), B.EnsembleB200(0), trajectories=n, saveat=0.1,
                  fp_anderson_reset=reset)
    grid = B.saveat_grid(0.1, (0.0, 2.0))
    ref = O.solve(O.default_config(model=O.MODEL_STIFF, fp_alg=1, fp_anderson_reset=reset), u0, p, [0.005], 0.0, 2.0, saveat=grid,
                  nthreads=O.hardware_threads())
    assert_same_ensemble(sol, ref)
    assert np.array_equal(sol.stats[:, 7], ref["stats"][:, 7]) and int(sol.stats[:, 7].sum()) > n   # nsolve
    with pytest.raises(B.B200DDEError, match="Rosenbrock23"):
        B.solve(prob, B.MethodOfSteps(B.Rosenbrock23(), fpsolve=B.NLAnderson()), saveat=1.0)


# ---- solve(...; tstops, d_discontinuities) (src/solve.jl:45-46,258-270; KAT test/interface/discontinuities.jl:62-79) --
@pytest.mark.parametrize("algname", ["Tsit5", "BS3"])
def test_user_stops_everystep_bit_exact(algname):
    model, oid, u0, p, lags, _ = CASES["two_delays_long"]
    prob = B.DDEProblem(model, u0, None, (0.0, 1.0), p, constant_lags=lags)
    tstops, discs = [0.25, 0.7, 5.0, 0.7], [(0.6, 5), (0.3, 4)]
    s = B.solve(prob, B.MethodOfSteps(getattr(B, algname)()), tstops=tstops, d_discontinuities=[B.Discontinuity(*d) for d in discs])
    try:
        O.set_stops(tstops, discs)
        o = O.solve_everystep(O.default_config(model=oid, alg=O.ALG_TSIT5 if algname == "Tsit5" else O.ALG_BS3), u0, p, lags, 0.0, 1.0)
    finally:
        O.set_stops()
    assert_same_solution(s, o)
    for ts in (0.25, 0.3, 0.6, 0.7):
        assert ts in s.t


def test_user_stops_ensemble_and_state_dependent_bit_exact():
    """User stops merge with the propagated heaps (bc_model, saveat output) and with the stops that src/track.jl adds for
    state-dependent lags (the dynamic half of the same user heap)."""
    u0, p = bc_sweep(3000, seed=9)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    tstops, discs = [0.5, 2.0, 2.0, 7.25], [(1.5, 1), (3.0, 0), (4.0, 7)]
    sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, tstops=tstops, d_discontinuities=discs)
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    try:
        O.set_stops(tstops, discs)
        ref = O.solve(O.default_config(model=O.MODEL_BC), u0, p, [1.0], 0.0, 10.0, saveat=grid, nthreads=O.hardware_threads())
        o = oracle_everystep("state_dependent")
        o2 = oracle_everystep("backwards")
    finally:
        O.set_stops()
    assert_same_ensemble(sol, ref)
    base = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1)
    assert int(sol.stats[:, 0].sum()) > int(base.stats[:, 0].sum())     # the stops really were taken
    assert_same_solution(gpu_everystep("state_dependent", tstops=tstops, d_discontinuities=discs), o)
    # reverse time: only the stops inside (t0, tf] = (2, 0] survive, sorted along the direction of integration
    assert_same_solution(gpu_everystep("backwards", tstops=tstops, d_discontinuities=discs), o2)


# ---- the history ring: undersized rings are detected, automatic sizing --------------------------------------
def _device_solve(solver, u0, p, lags, tspan, saveat):
    import torch
    n, ns = u0.shape
    grid = B.saveat_grid(saveat, tspan)
    n_out = len(grid) + 2
    d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
    d_out = torch.full((n, n_out, ns), float("nan"), dtype=torch.float64, device="cuda")
    d_rc = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), lags, tspan, grid, True, True, d_out.data_ptr(), d_rc.data_ptr(),
                        d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return grid, d_out.cpu().numpy(), d_rc.cpu().numpy(), d_st.cpu().numpy()


@pytest.mark.parametrize("cap", [4, 8])
def test_undersized_ring_is_never_silent(cap):
    """With a ring smaller than a lag window needs, every trajectory must either be flagged B2D_RC_HISTORY_OVERFLOW or be
    bit-identical to the oracle — never a plausible wrong answer (the device entry point does not re-solve)."""
    for model, oid, (u0, p), lags, tspan, saveat in (
            ("bc_model", O.MODEL_BC, bc_sweep(4000, seed=21), [1.0], (0.0, 10.0), 0.1),
            ("mackey_glass", O.MODEL_MACKEY_GLASS, mg_sweep(4000, seed=22), [17.0], (0.0, 300.0), 10.0)):
        prob = B.DDEProblem(model, u0[0], None, tspan, p[0], constant_lags=lags)
        solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {"ring_capacity": cap}))
        grid, u, rc, st = _device_solve(solver, u0, p, lags, tspan, saveat)
        solver.close()
        ref = O.solve(O.default_config(model=oid), u0, p, lags, tspan[0], tspan[1], saveat=grid, nthreads=O.hardware_threads())
        ok = rc == 1
        assert set(np.unique(rc)) <= {1, 6}
        assert (rc == 6).any()                                     # the ring really is too small for some
        assert np.array_equal(u[ok], ref["u"][ok]) and np.array_equal(st[ok, :6], ref["stats"][ok, :6])


def test_ring_capacity_is_sized_by_a_pilot_launch():
    """ring_capacity = 0 (default): a strided pilot sample sizes the ring, so problems whose lag window holds many nodes
    (stiff Rosenbrock23 sweep: > 100) solve through the device entry point without overflow, and small ones keep 16."""
    n = 20000
    lam = 10.0 ** (2.0 + 3.0 * np.arange(n) / (n - 1))
    u0, p = np.tile([2.0, 0.5, 1.0], (n, 1)), lam[:, None].copy()
    prob = B.DDEProblem("stiff_kinetics", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Rosenbrock23()), 0, {}))
    grid, u, rc, st = _device_solve(solver, u0, p, [1.0], (0.0, 10.0), 0.1)
    cap = solver.last_timing()[2]
    solver.close()
    assert (rc == 1).all() and cap > 16 and cap >= st[:, 6].max()
    idx = np.arange(0, n, 37)
    ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_STIFF), u0[idx], p[idx], [1.0], 0.0, 10.0, saveat=grid,
                  nthreads=O.hardware_threads())
    assert np.array_equal(u[idx], ref["u"]) and np.array_equal(st[idx, :6], ref["stats"][:, :6])
    u0, p = bc_sweep(5000, seed=3)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    solver = B.Solver(B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {}))
    _, _, rc, _ = _device_solve(solver, u0, p, [1.0], (0.0, 10.0), 0.1)
    assert (rc == 1).all() and solver.last_timing()[2] == 16
    solver.close()


# ---- the lean kernel is a compile-time specialisation of the general one: same results wherever both apply ----------
def test_lean_and_general_kernels_agree_bit_for_bit(monkeypatch):
    """b2d_solve routes default forward solves to the lean kernel (output mode 0), solves with options to the general saveat
    kernel without the callback / Anderson code (mode 3) and those to the full general kernel (mode 2).  B2D_FORCE_GENERAL
    sends the same solves through mode 3, with B2D_FORCE_FULL through mode 2: results and statistics must be identical."""
    cases = []
    u0, p = bc_sweep(20000, seed=21)
    cases.append(("bc_model", u0, p, [1.0], (0.0, 10.0), 0.1, {}))
    u0, p = mg_sweep(6000, seed=22)
    cases.append(("mackey_glass", u0, p, [17.0], (0.0, 400.0), 10.0, {}))
    u0, p = bc_sweep(4000, seed=23)
    cases.append(("bc_model", u0, p, [1.0], (0.0, 10.0), 0.25, dict(abstol=1e-9, reltol=1e-7)))
    rng = np.random.default_rng(24)
    cases.append(("state_dependent", 0.5 + rng.random((5000, 1)), np.zeros((5000, 1)), [], (0.0, 10.0), 0.1, {}))
    for model, u0, p, lags, tspan, saveat, kw in cases:
        prob = B.DDEProblem(model, u0[0], None, tspan, p[0], constant_lags=lags)
        ens = B.EnsembleProblem(prob, u0=u0, p=p)
        monkeypatch.delenv("B2D_FORCE_GENERAL", raising=False)
        lean = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=len(u0), saveat=saveat, **kw)
        monkeypatch.setenv("B2D_FORCE_GENERAL", "1")
        gen = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=len(u0), saveat=saveat, **kw)
        monkeypatch.setenv("B2D_FORCE_FULL", "1")
        full = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=len(u0), saveat=saveat, **kw)
        monkeypatch.delenv("B2D_FORCE_GENERAL", raising=False)
        monkeypatch.delenv("B2D_FORCE_FULL", raising=False)
        assert lean.converged and gen.converged and full.converged
        assert np.array_equal(lean.array, gen.array) and np.array_equal(lean.stats, gen.stats), model
        assert np.array_equal(lean.array, full.array) and np.array_equal(lean.stats, full.stats), model
    # an option both general kernels serve (user stops + save_idxs): mode 3 by default, mode 2 when forced
    u0, p = bc_sweep(3000, seed=25)
    prob = B.DDEProblem("bc_model", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    ens = B.EnsembleProblem(prob, u0=u0, p=p)
    a = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, tstops=[0.37, 4.2], save_idxs=[2, 0])
    monkeypatch.setenv("B2D_FORCE_FULL", "1")
    b = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=3000, saveat=0.1, tstops=[0.37, 4.2], save_idxs=[2, 0])
    monkeypatch.delenv("B2D_FORCE_FULL", raising=False)
    assert a.converged and np.array_equal(a.array, b.array) and np.array_equal(a.stats, b.stats)


def test_rosenbrock23_fixed_point_ensemble_in_the_lean_kernel():
    """Rosenbrock23 with steps longer than the lag (loose tolerances): every pass of the fixed-point iteration re-fetches
    the history of the pass through the rolled prefetch (value at t, t + dt/2, t + dt; derivative at t) against the
    provisional node of the step in flight.  saveat output = the lean kernel."""
    rng = np.random.default_rng(31)
    n = 4000
    u0 = 0.25 + 2.0 * rng.random((n, 1))
    p = np.zeros((n, 1))
    for tol in (dict(abstol=1e-2, reltol=1e-1), dict(abstol=1e-5, reltol=1e-3)):
        prob = B.DDEProblem("constant_1delay", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
        sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.Rosenbrock23()), B.EnsembleB200(0), trajectories=n,
                      saveat=0.5, **tol)
        ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_1DELAY, **tol), u0, p, [1.0], 0.0, 10.0,
                      saveat=B.saveat_grid(0.5, (0.0, 10.0)), nthreads=O.hardware_threads())
        assert_same_ensemble(sol, ref)
        assert np.array_equal(sol.stats[:, 7], ref["stats"][:, 7])
    loose = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_1DELAY, abstol=1e-2, reltol=1e-1), u0[:64], p[:64], [1.0], 0.0, 10.0,
                    saveat=B.saveat_grid(0.5, (0.0, 10.0)), nthreads=1)
    assert int(loose["stats"][:, 3].sum()) > 0, "the loose-tolerance case was meant to exercise fixed-point iterations"


# ---------------------------------------------------------------------------------------------------------
# Mass-matrix DDAE under MethodOfSteps(Rosenbrock23()): test/interface/mass_matrix.jl
# ---------------------------------------------------------------------------------------------------------
def test_sir_ddae_everystep_bit_exact_and_reference_kat():
    u0, p, lags, tspan = [0.99, 0.01, 0.0], [0.5], [4.0], (0.0, 40.0)
    prob = B.DDEProblem("sir_ddae", u0, None, tspan, p, constant_lags=lags)
    assert prob.isdae and prob.neutral and np.array_equal(prob.mass_matrix, [1.0, 1.0, 0.0])   # mass_matrix.jl:58-62
    s = B.solve(prob, B.MethodOfSteps(B.Rosenbrock23()), max_steps=20000)
    o = O.solve_everystep(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_SIR_DDAE, neutral=1), u0, p, lags, 0.0, 40.0)
    assert_same_solution(s, o)
    assert s.stats.nsolve == int(o.stats[O.STAT_NSOLVE]) and s.stats.nsolve > 0
    assert s.t[1] - s.t[0] == 1e-6                                       # upstream initdt of a DAE
    assert np.max(np.abs(s.u.sum(axis=1) - 1.0)) < 1e-14                 # 0 = S + I + R - 1
    # mass_matrix.jl:65-74: L-inf < 5e-3 against an accurate solution of the equivalent DDE (GPU, Tsit5 at 1e-13)
    acc = B.solve(B.DDEProblem("sir", u0, None, tspan, p, constant_lags=lags), B.MethodOfSteps(B.Tsit5()), abstol=1e-13,
                  reltol=1e-13, saveat=s.t)
    assert acc.u.shape == s.u.shape
    assert np.max(np.abs(acc.u - s.u)) < 5.0e-3
    # the non-neutral propagation rule is the other kernel mode of the same model
    s2 = B.solve(prob.remake(neutral=False), B.MethodOfSteps(B.Rosenbrock23()), max_steps=20000)
    o2 = O.solve_everystep(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_SIR_DDAE), u0, p, lags, 0.0, 40.0)
    assert_same_solution(s2, o2)


def test_sir_ddae_ensemble_bit_exact():
    """Sweep over the infection rate and the initially infected fraction (u0 stays consistent: S + I + R = 1);
    neutral (general kernel) and non-neutral (lean kernel) propagation."""
    n = 3000
    g = 0.2 + 0.8 * np.arange(n) / (n - 1)
    i0 = 0.001 + 0.05 * ((np.arange(n) * 7919) % n) / n
    u0 = np.stack([1.0 - i0, i0, np.zeros(n)], axis=1)
    u0[:, 2] = 1.0 - u0[:, 0] - u0[:, 1]
    p = g[:, None].copy()
    grid = B.saveat_grid(0.5, (0.0, 40.0))
    for neutral in (True, False):
        prob = B.DDEProblem("sir_ddae", u0[0], None, (0.0, 40.0), p[0], constant_lags=[4.0], neutral=neutral)
        sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.Rosenbrock23()), B.EnsembleB200(0), trajectories=n,
                      saveat=0.5)
        ref = O.solve(O.default_config(alg=O.ALG_ROSENBROCK23, model=O.MODEL_SIR_DDAE, neutral=int(neutral)), u0, p, [4.0],
                      0.0, 40.0, saveat=grid, nthreads=O.hardware_threads())
        assert sol.converged
        assert_same_ensemble(sol, ref)
        assert np.array_equal(sol.stats[:, 7], ref["stats"][:, 7])
        assert np.max(np.abs(sol.array.sum(axis=2) - 1.0)) < 1e-13


def test_mass_matrix_needs_rosenbrock23():
    prob = B.DDEProblem("sir_ddae", [0.99, 0.01, 0.0], None, (0.0, 40.0), [0.5], constant_lags=[4.0])
    with pytest.raises(B.B200DDEError, match="mass matrix"):
        B.solve(prob, B.MethodOfSteps(B.Tsit5()))
