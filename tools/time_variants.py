"""Kernel time of the bc_model ensemble for a few output/option variants (cost attribution experiments)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))
import b200dde as B
from b200dde.api import _make_config
import bench
n = 1_000_000
wl = bench.workload("bc", n, 0, 1)
prob = B.DDEProblem(wl["model"], wl["u0"][0], None, wl["tspan"], wl["p"][0], constant_lags=wl["lags"])
d_u0 = torch.from_numpy(wl["u0"]).cuda(); d_p = torch.from_numpy(wl["p"]).cuda()
d_rc = torch.empty(n, dtype=torch.int32, device="cuda"); d_st = torch.empty((n, 8), dtype=torch.int64, device="cuda")
for name, saveat, kw in (("saveat 0.1 (101 pts)", 0.1, {}), ("saveat 5 (3 pts)", 5.0, {}), ("saveat 0.1, exact pow controller", 0.1, dict(controller_pow=0))):
    solver = B.Solver(_make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, kw))
    grid = B.saveat_grid(saveat, wl["tspan"]); n_out = len(grid) + 2
    d_out = torch.empty((n, n_out, 3), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    f = lambda: solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), wl["lags"], wl["tspan"], grid, True, True, d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr(), n, st.cuda_stream)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/3:.3f} ms, accepted {int(d_st[:,0].sum())}")
    solver.close()
