"""Regenerate tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is Julia-only and cannot run in this image, so these vectors are NOT outputs of the reference; they
freeze the KAT-pinned oracle (floats are stored with repr(), i.e. bit-exactly) so that (a) an accidental change of
the oracle's arithmetic is caught on CPU and (b) the GPU tests can compare against committed values.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: README bc_model, single trajectory, Tsit5, defaults, every step
    "c1_bc_model": dict(model="MODEL_BC", u0=[1.0, 1.0, 1.0], p=[0.2, 0.3, 1.0], lags=[1.0], tspan=[0.0, 10.0]),
    # test/interface/parameters.jl:13-30
    "logistic_p03": dict(model="MODEL_LOGISTIC", u0=[0.1], p=[0.3], lags=[1.0], tspan=[0.0, 50.0]),
    # one trajectory of BASELINE.json configs[2]
    "mackey_glass": dict(model="MODEL_MACKEY_GLASS", u0=[1.2], p=[0.2, 1.2], lags=[17.0], tspan=[0.0, 1000.0]),
}

for name, c in CASES.items():
    cfg = O.default_config(model=getattr(O, c["model"]))
    s = O.solve_everystep(cfg, c["u0"], c["p"], c["lags"], c["tspan"][0], c["tspan"][1])
    out = dict(c)
    out.update(t=s.t.tolist(), u=s.u.tolist(), tracked=[list(x) for x in s.tracked], stats=s.stats[:6].tolist(),
               retcode=s.retcode)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=0)
    print(name, len(s), s.stats[:6].tolist())
