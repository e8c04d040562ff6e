"""Small end-to-end exercise of every kernel variant for compute-sanitizer (memcheck / racecheck): tiny ensembles only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))
import b200dde as B

def run(model, u0, p, lags, tspan, alg, n=70, saveat=None, vary=1e-3, **kw):
    prob = B.DDEProblem(model, u0, None, tspan, p, constant_lags=lags)
    U0 = np.tile(np.asarray(u0, float), (n, 1)) * (1.0 + vary * np.arange(n))[:, None]
    P = np.tile(np.asarray(p, float), (n, 1))
    sol = B.solve(B.EnsembleProblem(prob, u0=U0, p=P), alg, B.EnsembleB200(0), trajectories=n,
                  saveat=saveat if saveat is not None else (tspan[1] - tspan[0]) / 10, **kw)
    s1 = B.solve(prob, alg, max_steps=20000)
    print(model, type(alg.alg).__name__, "ok", int((sol.retcodes == 1).sum()), len(s1))

T, R, S = B.MethodOfSteps(B.Tsit5()), B.MethodOfSteps(B.Rosenbrock23()), B.MethodOfSteps(B.BS3())
run("bc_model", [1, 1, 1], [.2, .3, 1], [1.0], (0, 10), T)
run("bc_model", [1, 1, 1], [.2, .3, 1], [1.0], (0, 10), T, abstol=1e-9, reltol=1e-9)      # ring regrowth
run("mackey_glass", [1.2], [.2, 1.2], [17.0], (0, 300), T)
run("constant_2delays", [1.0], [0.0], [1 / 3, 1 / 5], (0, 30), T)                           # local-memory queues, fixed point
run("constant_1delay_long", [1.0], [0.0], [0.2], (0, 30), T)
run("state_dependent", [1.0], [0.0], [], (0, 10), T)                                         # tracking
run("constant_1delay_dependent", [1.0], [0.0], [], (0, 10), S)
run("backwards", [1.0], [0.0], [-1.0], (2, 0), T)
run("stiff_kinetics", [2, .5, 1], [1e3], [1.0], (0, 10), R)
run("constant_1delay", [1.0], [0.0], [1.0], (0, 10), R)
run("constant_1delay", [1.0], [0.0], [1.0], (0, 10), S)
run("sir_ddae", [.99, .01, 0.], [.5], [4.0], (0, 40), R, vary=0.0)                             # mass matrix, neutral: general kernel
run("sir", [.99, .01, 0.], [.5], [4.0], (0, 40), T)
um = B.UserModel(open(os.path.join(ROOT, "tests/user_models/mackey_glass_user.cu")).read(), "MyMackeyGlass", (1, 2, 1, 0))
run(um, [1.2], [.2, 1.2], [17.0], (0, 200), T)
print("sanitizer smoke done")
