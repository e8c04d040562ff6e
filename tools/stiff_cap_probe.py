#!/usr/bin/env python
"""Does the ring CAPACITY of the stiff sweep cost time or DRAM traffic, or does the live window?

The Rosenbrock23 sweep (lambda 1e2..1e5) needs 99 .. 312 ring nodes per trajectory (steps inside one lag); the pilot sizes the
launch for the worst trajectory (512).  This probe solves the lambda <= 190 tenth of the sweep (need <= 128) on 10^6 trajectories
with capacity 128, 256 and 512 on the device-resident path and prints the kernel time of each: if the capacity were what costs,
128 would be faster than 512.  (The bytes a trajectory keeps live are need x 88 B whatever the capacity: at 75 776 resident
threads that is 0.66 .. 2.1 GB against 126 MB of L2.)"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delaydiffeq.jl_b200"))
import b200dde as B
from b200dde.api import _make_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
for lo, hi in ((2.0, 2.28), (4.0, 5.0)):
    lam = 10.0 ** (lo + (hi - lo) * np.arange(n) / (n - 1))
    u0 = np.tile([2.0, 0.5, 1.0], (n, 1))
    p = lam[:, None].copy()
    grid = B.saveat_grid(0.1, (0.0, 10.0))
    prob = B.DDEProblem("stiff_kinetics", u0[0], None, (0.0, 10.0), p[0], constant_lags=[1.0])
    alg = B.MethodOfSteps(B.Rosenbrock23())
    d_u0, d_p = torch.from_numpy(u0).to(dev), torch.from_numpy(p).to(dev)
    d_out = torch.empty((n, len(grid) + 2, 3), dtype=torch.float64, device=dev)
    d_rc = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty((n, B._lib.N_STATS), dtype=torch.int64, device=dev)
    ref = None
    for cap in (0, 128, 256, 512, 1024):
        solver = B.Solver(_make_config(prob, alg, 0, {"ring_capacity": cap} if cap else {}))
        ms = []
        for _ in range(4):
            solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), [1.0], (0.0, 10.0), grid, True, True, d_out.data_ptr(),
                                d_rc.data_ptr(), d_st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ms.append(solver.last_timing()[0])
        rc = d_rc.cpu().numpy()
        out = d_out.cpu().numpy()
        same = True if ref is None else bool(np.array_equal(out[rc == 1], ref[rc == 1]))
        if ref is None:
            ref = out
        print(f"lambda 1e{lo}..1e{hi}  capacity {cap or 'auto'} -> {solver.last_timing()[2]}: kernel {min(ms[1:]):.3f} ms, "
              f"success {int((rc == 1).sum())}/{n}, overflow {int((rc == 6).sum())}, max need {int(d_st[:, 6].max().item())}, same results {same}",
              flush=True)
        solver.close()
