import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delaydiffeq.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import b200dde as B
import oracle_lib as O
prob = B.DDEProblem("state_dependent", [1.0], None, (0.0, 10.0), [0.0], constant_lags=[])
cb = B.ContinuousCallback(("state", 0, 0.3), ("add", 0, 0.4))
s = B.solve(prob, B.MethodOfSteps(B.Tsit5()), callback=cb)
o = O.solve_everystep(O.default_config(model=O.MODEL_STATE_DEP, callback=6, cb_affect=2, cb_idx=0, cb_cond_idx=0, cb_par0=0.3, cb_par1=0.4),
                      [1.0], [0.0], [], 0.0, 10.0)
for i in range(8):
    print(i, repr(float(s.t[i])), repr(float(o.t[i])), repr(float(s.u[i, 0])), repr(float(o.u[i, 0])))
print("gpu tracked", s.tracked_discontinuities[:6])
print("ora tracked", o.tracked[:6])
print("gpu stats", s.stats.naccept, s.stats.nreject, s.stats.nf, s.stats.nfpiter, "ora", o.stats[:5])
# dense history of the event step: compare slopes
print("gpu k[2]", s.k[2] if s.k is not None else None)
print("ora k[2]", o.k[2])
print("gpu k[3]", s.k[3] if s.k is not None else None)
print("ora k[3]", o.k[3])
