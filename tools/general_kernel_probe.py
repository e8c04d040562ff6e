#!/usr/bin/env python
"""What the general kernels cost: the bc_model and Mackey-Glass sweeps (10^6 trajectories, device-resident) solved by the lean
kernel, by the general saveat kernel without any option in use (B2D_FORCE_GENERAL), with a user tstop, with a ContinuousCallback
that fires in most trajectories, and with per-trajectory lags."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delaydiffeq.jl_b200"))
import b200dde as B
from b200dde.api import _make_config

n = 1_000_000
i = np.arange(n)
dev = torch.device("cuda:0")


def run(name, prob, u0, p, lags, tspan, dt_save, cfg_extra=None, cb=None, tstops=None, env=None, traj_lags=None):
    if env:
        os.environ.update(env)
    cfg = _make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, cfg_extra or {})
    if cb is not None:
        cb.apply(cfg)
    solver = B.Solver(cfg)
    if tstops:
        solver.set_tstops(tstops)
    grid = B.saveat_grid(dt_save, tspan)
    if traj_lags is not None:
        solver.set_trajectory_lags(traj_lags)
        t, u, rc, st = solver.solve_host(u0, p, lags, tspan, grid, True, True)
        t, u, rc, st = solver.solve_host(u0, p, lags, tspan, grid, True, True)
        ms = solver.last_timing()[0]
        acc, ok = int(st[:, 0].sum()), int((rc == 1).sum())
    else:
        d_u0, d_p = torch.from_numpy(u0).to(dev), torch.from_numpy(p).to(dev)
        d_out = torch.empty((n, len(grid) + 2, u0.shape[1]), dtype=torch.float64, device=dev)
        d_rc = torch.empty(n, dtype=torch.int32, device=dev)
        d_st = torch.empty((n, B._lib.N_STATS), dtype=torch.int64, device=dev)
        for _ in range(3):
            solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), lags, tspan, grid, True, True, d_out.data_ptr(), d_rc.data_ptr(),
                                d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        ms = solver.last_timing()[0]
        acc, ok = int(d_st[:, 0].sum().item()), int((d_rc == 1).sum().item())
    solver.close()
    if env:
        for k in env:
            os.environ.pop(k, None)
    print(f"{name:64s} {ms:8.2f} ms  {acc / ms / 1e6:7.2f}e9 steps/s  success {ok}/{n}", flush=True)


a, b, c = i // 10000, (i // 100) % 100, i % 100
p_bc = np.ascontiguousarray(np.stack([0.1 + 0.2 * a / 99, 0.2 + 0.2 * b / 99, 0.5 + 1.0 * c / 99], axis=1))
u_bc = np.ones((n, 3))
prob_bc = B.DDEProblem("bc_model", u_bc[0], None, (0.0, 10.0), p_bc[0], constant_lags=[1.0])
p_mg = np.ascontiguousarray(np.stack([0.18 + 0.04 * (i // 4000) / 249, 0.5 + 1.0 * (i % 4000) / 3999], axis=1))
u_mg = p_mg[:, 1:2].copy()
prob_mg = B.DDEProblem("mackey_glass", u_mg[0], None, (0.0, 1000.0), p_mg[0], constant_lags=[17.0])
for nm, prob, u0, p, lags, tspan, dts, cb, stop in (
        ("bc_model", prob_bc, u_bc, p_bc, [1.0], (0.0, 10.0), 0.1, B.ContinuousCallback(("state", 2, 0.8), ("add", 1, 0.25)), 3.3),
        ("mackey_glass", prob_mg, u_mg, p_mg, [17.0], (0.0, 1000.0), 10.0, B.ContinuousCallback(("state", 0, 1.0), ("add", 0, -0.3), direction=1), 333.3)):
    run(f"{nm}: lean kernel", prob, u0, p, lags, tspan, dts)
    run(f"{nm}: general kernel, no option in use", prob, u0, p, lags, tspan, dts, env={"B2D_FORCE_GENERAL": "1"})
    run(f"{nm}: general kernel, one user tstop", prob, u0, p, lags, tspan, dts, tstops=[stop])
    run(f"{nm}: general kernel, ContinuousCallback (state threshold)", prob, u0, p, lags, tspan, dts, cb=cb)
    run(f"{nm}: general kernel, per-trajectory lags (host path, last wave's kernel)", prob, u0, p, lags, tspan, dts,
        traj_lags=np.tile(np.asarray(lags), (n, 1)) * (1.0 + 0.1 * (i % 7)[:, None] / 7))
