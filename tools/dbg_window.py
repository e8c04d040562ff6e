import sys, os
sys.path.insert(0, "delaydiffeq.jl_b200"); sys.path.insert(0, ".")
import numpy as np, b200dde as B, bench
wl = bench.workload("mg", 100000, 0, 1)
prob = B.DDEProblem(wl["model"], wl["u0"][0], None, wl["tspan"], wl["p"][0], constant_lags=wl["lags"])
ens = B.EnsembleProblem(prob, u0=wl["u0"], p=wl["p"])
sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=100000, saveat=10.0)
d = sol.stats[:, 7]
att = sol.stats[:, 0] + sol.stats[:, 1]
print("attempts/traj", att.mean(), " fallback(ok-fail) per traj", (d & 0xff).mean(), " n_nodes==1:", ((d >> 8) & 0xff).mean(), " inflight:", ((d >> 16) & 0xff).mean(), " beyond-t:", ((d >> 24) & 0xff).mean())
print("hist of ok-fail count:", np.bincount((d & 0xff))[:12])
