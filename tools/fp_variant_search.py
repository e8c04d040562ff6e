#!/usr/bin/env python
"""Search over what the un-vendored upstream `nlsolve!` might do differently: every combination of the
ORACLE_FPVAR switches of oracle/oracle_dde.cpp against the two reference KATs that disagree —
  test/interface/parameters.jl:28,35   exact step counts 21 / 47 of the delayed logistic problem
  test/interface/fpsolve.jl:24-27      nf > 3000, nfpiter > 300, nfpconvfail > 50 on prob_dde_constant_2delays_long
One process per variant (the switch is read once).  `fp_variant_search.py V` prints one variant as JSON;
without arguments all 512 are run and the distinct outcomes are tabulated (CPU only, about two minutes).
"""
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(v):
    os.environ["ORACLE_FPVAR"] = str(v)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    def es(model, u0, p, lags, t0, tf):
        return O.solve_everystep(O.default_config(model=model), u0, p, lags, t0, tf)
    s1 = es(O.MODEL_LOGISTIC, [0.1], [0.3], [1.0], 0.0, 50.0)
    s2 = es(O.MODEL_LOGISTIC, [0.1], [1.4], [1.0], 0.0, 50.0)
    s = es(O.MODEL_2DELAYS, [1.0], [0.0], [1 / 3, 1 / 5], 0.0, 100.0)
    return dict(V=v, steps_p03=len(s1), steps_p14=len(s2), nf=int(s.stats[O.STAT_NF]), nfpiter=int(s.stats[O.STAT_NFPITER]),
                nfpconvfail=int(s.stats[O.STAT_NFPCONVFAIL]))


def bits(v):
    return dict(conv_first=v & 1, keep_eta=(v >> 1) & 1, maxit_ok=(v >> 2) & 1, prec=(v >> 3) & 1, rule=(v >> 4) & 3,
                first=(v >> 6) & 3, reset_eta=(v >> 8) & 1)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(one(int(sys.argv[1]))))
        sys.exit(0)
    rows = []
    for v in range(512):
        r = subprocess.run([sys.executable, __file__, str(v)], capture_output=True, text=True)
        rows.append(json.loads(r.stdout))
    both = [r for r in rows if r["steps_p03"] == 21 and r["steps_p14"] == 47 and r["nf"] > 3000 and r["nfpiter"] > 300 and r["nfpconvfail"] > 50]
    print(f"{len(rows)} variants, {len(both)} satisfy both KATs")
    out = collections.defaultdict(list)
    for r in rows:
        out[(r["steps_p03"], r["steps_p14"], r["nf"], r["nfpiter"], r["nfpconvfail"])].append(r["V"])
    print("steps(p=0.3) steps(p=1.4) nf nfpiter nfpconvfail | variants | switch values seen")
    for k, vs in sorted(out.items(), key=lambda kv: -len(kv[1])):
        b = [bits(x) for x in vs]
        print(k, len(vs), {key: sorted(set(bb[key] for bb in b)) for key in b[0]})
