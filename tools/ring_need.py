"""Distribution of the live history nodes a trajectory needs (B2D_STAT_MAXNODES) for the bench workloads."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))
import b200dde as B
import bench
n = 20000
for name in ("bc", "mg", "sd", "stiff"):
    wl = bench.workload(name, n, 0, 1)
    prob = B.DDEProblem(wl["model"], wl["u0"][0], None, wl["tspan"], wl["p"][0], constant_lags=wl["lags"])
    ens = B.EnsembleProblem(prob, u0=wl["u0"], p=wl["p"])
    alg = B.MethodOfSteps(B.Rosenbrock23() if wl.get("alg") == "Rosenbrock23" else B.Tsit5())
    sol = B.solve(ens, alg, B.EnsembleB200(0), trajectories=n, saveat=wl["saveat"], ring_capacity=1024)
    m = sol.stats[:, 6]; acc = sol.stats[:, 0]
    print(name, "rc", np.unique(sol.retcodes), "maxnodes pct 50/90/99/99.9/max", [int(np.percentile(m, q)) for q in (50, 90, 99, 99.9, 100)],
          "accepted mean/max", float(acc.mean()), int(acc.max()))
    for q in (16, 32, 64, 128, 256):
        print("   frac needing >", q, float((m > q).mean()), " share of steps", float(acc[m > q].sum() / acc.sum()))
