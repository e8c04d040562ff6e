"""Full-size Mackey-Glass ensemble (10^6 trajectories) with a ContinuousCallback on a state threshold through b2d_solve: every
trajectory Success, a strided sample bit-exact against the oracle (GPU box; profiles/README.md)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delaydiffeq.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import b200dde as B
import oracle_lib as O
n = 1_000_000
i = np.arange(n)
p = np.ascontiguousarray(np.stack([0.18 + 0.04 * (i // 4000) / 249, 0.5 + 1.0 * (i % 4000) / 3999], axis=1))
u0 = p[:, 1:2].copy()
prob = B.DDEProblem("mackey_glass", u0[0], None, (0.0, 1000.0), p[0], constant_lags=[17.0])
ens = B.EnsembleProblem(prob, u0=u0, p=p)
cb = B.ContinuousCallback(("state", 0, 1.0), ("add", 0, -0.3), direction=1)
sol = B.solve(ens, B.MethodOfSteps(B.Tsit5()), B.EnsembleB200(0), trajectories=n, saveat=10.0, callback=cb)
bad = np.nonzero(sol.retcodes != 1)[0]
print("bad", bad[:10], sol.retcodes[bad][:10], sol.stats[bad][:3])
idx = np.concatenate([bad[:3], np.arange(0, n, 49999)])
ref = O.solve(O.default_config(model=O.MODEL_MACKEY_GLASS, callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=1.0, cb_par1=-0.3, cb_direction=1),
              u0[idx], p[idx], [17.0], 0.0, 1000.0, saveat=B.saveat_grid(10.0, (0.0, 1000.0)), nthreads=8)
print("oracle rc", ref["retcode"][:6], ref["stats"][:3])
ok = [bool(np.array_equal(sol.array[j], ref["u"][k], equal_nan=True)) and sol.retcodes[j] == ref["retcode"][k] for k, j in enumerate(idx)]
print("match", sum(ok), "of", len(ok), [int(idx[k]) for k in range(len(ok)) if not ok[k]])
