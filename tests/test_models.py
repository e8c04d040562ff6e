"""The model functors on their own: csrc/models.cuh (device) against oracle/oracle_models.h (CPU), and the analytic
Jacobians / time gradients of the CPU restatement against finite differences.

GPU == oracle parity of whole solves cannot see a transcription slip that both model files share; these tests compare the
two independently written files point by point and check the derivative members, which
Rosenbrock23 uses where the reference uses ForwardDiff, against the right-hand side itself.
"""
import ctypes as C

import numpy as np
import pytest

import b200dde as B
import oracle_lib as O

# registry name -> (oracle id, n_state, n_param, n_const_lags, n_dep_lags, has_jac)
MODELS = {name: mid for name, mid in B.MODELS.items()}
JAC = {"stiff_kinetics", "constant_1delay", "sir_ddae"}


def dims(name):
    n = [C.c_int32() for _ in range(5)]
    B._lib.check(B.lib().b2d_model_info(MODELS[name], *[C.byref(x) for x in n]))
    return tuple(x.value for x in n[:4])


def points(name, n_pts, seed):
    n, npar, ncl, ndl = dims(name)
    slots = max(ncl + ndl, 1)
    rng = np.random.default_rng(seed)
    u = rng.uniform(-2.0, 2.0, (n_pts, n))
    p = rng.uniform(0.1, 3.0, (n_pts, max(npar, 1)))[:, :npar] if npar else np.zeros((n_pts, 0))
    hv = rng.uniform(-2.0, 2.0, (n_pts, slots, n))
    dhv = rng.uniform(-2.0, 2.0, (n_pts, slots, n))
    lags = list(rng.uniform(0.1, 2.0, ncl))
    return u, np.ascontiguousarray(p), hv, dhv, lags, (n, npar, ncl, ndl)


@pytest.mark.parametrize("name", sorted(JAC))
def test_analytic_jacobian_and_time_gradient_match_finite_differences(name):
    u, p, hv, dhv, lags, (n, npar, ncl, ndl) = points(name, 64, 5)
    f0, jac, tg, _ = O.eval_model(MODELS[name], u, p, hv, dhv, 0.7, lags, ndl, True)
    eps = 1e-6
    for j in range(n):
        up, um = u.copy(), u.copy()
        up[:, j] += eps; um[:, j] -= eps
        fp = O.eval_model(MODELS[name], up, p, hv, dhv, 0.7, lags, ndl, True)[0]
        fm = O.eval_model(MODELS[name], um, p, hv, dhv, 0.7, lags, ndl, True)[0]
        assert np.allclose((fp - fm) / (2 * eps), jac[:, :, j], rtol=1e-6, atol=1e-6), (name, j)
    # f depends on t only through the history: d f / d t = (d f / d h) h'(t - lag), checked along the direction dhv
    fp = O.eval_model(MODELS[name], u, p, hv + eps * dhv, dhv, 0.7, lags, ndl, True)[0]
    fm = O.eval_model(MODELS[name], u, p, hv - eps * dhv, dhv, 0.7, lags, ndl, True)[0]
    assert np.allclose((fp - fm) / (2 * eps), tg, rtol=1e-6, atol=1e-6), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MODELS))
def test_device_models_equal_the_cpu_restatement_bit_for_bit(name):
    n_pts = 4096
    u, p, hv, dhv, lags, (n, npar, ncl, ndl) = points(name, n_pts, 11)
    has_jac = name in JAC
    f = np.zeros((n_pts, n)); jac = np.zeros((n_pts, n, n)); tg = np.zeros((n_pts, n)); lag = np.zeros((n_pts, max(ndl, 1)))
    lg = np.ascontiguousarray(lags + [0.0] * 4)
    rc = B.lib().b2d_eval_model(MODELS[name], 0, n_pts, u.ctypes.data, p.ctypes.data, hv.ctypes.data, dhv.ctypes.data, 0.7,
                                lg.ctypes.data, f.ctypes.data, jac.ctypes.data, tg.ctypes.data, lag.ctypes.data)
    B._lib.check(rc)
    rf, rj, rt, rl = O.eval_model(MODELS[name], u, p, hv, dhv, 0.7, lags, ndl, has_jac)
    assert np.array_equal(f, rf), name
    if has_jac:
        assert np.array_equal(jac, rj) and np.array_equal(tg, rt), name
    if ndl:
        assert np.array_equal(lag[:, :ndl], rl), name
    assert np.isfinite(f).all() and np.abs(f).max() > 0 or name == "constant_1delay"
