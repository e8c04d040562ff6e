"""User-supplied models: CUDA C++ source compiled with NVRTC against the registry's kernel template."""
import os

import numpy as np
import pytest

import b200dde as B
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _um(fname, struct, dims):
    return B.UserModel(open(os.path.join(HERE, "user_models", fname)).read(), struct, dims)


def test_user_models_compile_without_a_gpu():
    # NVRTC is only a compiler: this runs in the CPU container and proves the embedded kernel headers are self-contained
    _um("mackey_glass_user.cu", "MyMackeyGlass", (1, 2, 1, 0)).compile_check()
    _um("state_dep_user.cu", "MyStateDep", (1, 1, 0, 1)).compile_check()


def test_user_model_compile_errors_are_reported():
    with pytest.raises(B.B200DDEError, match='no member "NP"'):
        B.UserModel("struct Broken { static constexpr int N = 1; };", "Broken", (1, 1, 0, 0)).compile_check()
    with pytest.raises(B.B200DDEError, match="invalid struct name"):
        B.UserModel("struct A {};", "A; system(", (1, 1, 0, 0)).compile_check()


@pytest.mark.gpu
def test_user_mackey_glass_matches_the_oracle_bit_for_bit():
    rng = np.random.default_rng(7)
    n = 2000
    p = np.stack([0.18 + 0.04 * rng.random(n), 0.5 + 1.0 * rng.random(n)], axis=1)
    u0 = p[:, 1:2].copy()
    um = _um("mackey_glass_user.cu", "MyMackeyGlass", (1, 2, 1, 0))
    prob = B.DDEProblem(um, u0[0], None, (0.0, 1000.0), p[0], constant_lags=[17.0])
    sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=n, saveat=10.0)
    ref = O.solve(O.default_config(model=O.MODEL_MACKEY_GLASS), u0, p, [17.0], 0.0, 1000.0, saveat=B.saveat_grid(10.0, (0.0, 1000.0)),
                  nthreads=O.hardware_threads())
    assert sol.converged
    assert np.array_equal(sol.array, ref["u"]) and np.array_equal(sol.stats[:, :6], ref["stats"][:, :6])


@pytest.mark.gpu
def test_user_model_with_its_own_discrete_callback():
    """condition / affect! as device functors of a run-time compiled model (callback = "model"): a kick u += 0.3 at t = 40
    must give what the built-in preset callback gives, which the oracle restates."""
    um = _um("mackey_glass_user.cu", "MyMackeyGlassKick", (1, 2, 1, 0))
    prob = B.DDEProblem(um, [1.2], None, (0.0, 200.0), [0.2, 1.2], constant_lags=[17.0])
    cb = B.DiscreteCallback("model", par=(40.0, 0.3))
    s = B.solve(prob, B.MethodOfSteps(B.Tsit5()), callback=cb, tstops=[40.0])
    try:
        O.set_stops([40.0], [])
        o = O.solve_everystep(O.default_config(model=O.MODEL_MACKEY_GLASS, callback=2, cb_affect=2, cb_idx=0, cb_par0=40.0, cb_par1=0.3),
                              [1.2], [0.2, 1.2], [17.0], 0.0, 200.0)
    finally:
        O.set_stops()
    assert np.array_equal(s.t, o.t) and np.array_equal(s.u, o.u) and s.tracked_discontinuities == o.tracked
    i = [k for k, x in enumerate(s.t) if x == 40.0]
    assert len(i) == 2 and abs(s.u[i[1], 0] - s.u[i[0], 0] - 0.3) < 1e-15


@pytest.mark.gpu
def test_user_model_with_its_own_continuous_callback():
    """condition (a real function whose sign change is the event) and affect! as device functors of a run-time compiled model
    (ContinuousCallback("model")): with the preset affect! the result is the built-in state-threshold callback's, which the
    oracle restates; with the model's own affect! (u *= 0.5 at every upcrossing of 1.0) the state right after each event is
    half the state right before it."""
    um = _um("mackey_glass_user.cu", "MyMackeyGlassThreshold", (1, 2, 1, 0))
    prob = B.DDEProblem(um, [1.2], None, (0.0, 300.0), [0.2, 1.2], constant_lags=[17.0])
    s = B.solve(prob, B.MethodOfSteps(B.Tsit5()), callback=B.ContinuousCallback("model", ("add", 0, -0.3), direction=1, par=(1.0,)))
    o = O.solve_everystep(O.default_config(model=O.MODEL_MACKEY_GLASS, callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=1.0,
                                           cb_par1=-0.3, cb_direction=1), [1.2], [0.2, 1.2], [17.0], 0.0, 300.0)
    # (the solve_everystep call returns the first 64 tracked discontinuities; the events and their images are > 200 here)
    assert np.array_equal(s.t, o.t) and np.array_equal(s.u, o.u) and s.tracked_discontinuities == o.tracked[:64]
    s = B.solve(prob, B.MethodOfSteps(B.Tsit5()), callback=B.ContinuousCallback("model", direction=1, par=(1.0, 0.5)))
    ev = [k for k in range(len(s.t) - 1) if s.t[k] == s.t[k + 1]]
    assert len(ev) >= 3
    for k in ev:
        assert abs(s.u[k, 0] - 1.0) < 1e-9 and s.u[k + 1, 0] == 0.5 * s.u[k, 0]


@pytest.mark.gpu
def test_user_model_in_the_general_saveat_kernel():
    """User tstops send the solve to the general kernel (output mode 2), whose shared-memory layout is larger than the
    lean kernel's: the library must size the launch from that kernel's own row count."""
    rng = np.random.default_rng(11)
    n = 1500
    p = np.stack([0.18 + 0.04 * rng.random(n), 0.5 + 1.0 * rng.random(n)], axis=1)
    u0 = p[:, 1:2].copy()
    um = _um("mackey_glass_user.cu", "MyMackeyGlass", (1, 2, 1, 0))
    prob = B.DDEProblem(um, u0[0], None, (0.0, 300.0), p[0], constant_lags=[17.0])
    tstops = [33.3, 150.0]
    sol = B.solve(B.EnsembleProblem(prob, u0=u0, p=p), B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=n, saveat=10.0,
                  tstops=tstops)
    try:
        O.set_stops(tstops, [])
        ref = O.solve(O.default_config(model=O.MODEL_MACKEY_GLASS), u0, p, [17.0], 0.0, 300.0, saveat=B.saveat_grid(10.0, (0.0, 300.0)),
                      nthreads=O.hardware_threads())
    finally:
        O.set_stops()
    assert sol.converged
    assert np.array_equal(sol.array, ref["u"]) and np.array_equal(sol.stats[:, :6], ref["stats"][:, :6])


@pytest.mark.gpu
def test_user_state_dependent_model_matches_the_oracle():
    um = _um("state_dep_user.cu", "MyStateDep", (1, 1, 0, 1))
    prob = B.DDEProblem(um, [1.0], None, (0.0, 10.0), [0.0])
    s = B.solve(prob, B.MethodOfSteps(B.Tsit5()), abstol=1e-9, reltol=1e-9, max_steps=20000)
    o = O.solve_everystep(O.default_config(model=O.MODEL_STATE_DEP, abstol=1e-9, reltol=1e-9), [1.0], [0.0], [], 0.0, 10.0)
    assert s.retcode == "Success" and np.array_equal(s.t, o.t) and np.array_equal(s.u, o.u)
    assert s.tracked_discontinuities == o.tracked


@pytest.mark.gpu
def test_user_model_dimension_mismatch_is_an_error():
    um = _um("mackey_glass_user.cu", "MyMackeyGlass", (1, 3, 1, 0))     # wrong n_param
    prob = B.DDEProblem(um, [1.0], None, (0.0, 10.0), [0.2, 1.0, 0.0], constant_lags=[17.0])
    with pytest.raises(B.B200DDEError, match="n_param"):
        B.solve(prob, B.MethodOfSteps(B.Tsit5()))
