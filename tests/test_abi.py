"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200dde.h declares,
agrees with the Python mirror on the config layout, and fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import b200dde as B
from b200dde import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200dde.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"libb200dde.so does not export {name}"
    assert sorted(_lib.exported_symbols()) == declared  # the Python mirror binds exactly the declared ABI


def test_version_and_config_layout():
    L = B.lib()
    assert L.b2d_version() == 100
    cfg = B.Config()
    L.b2d_config_default(C.byref(cfg), 0)
    assert cfg.struct_size == C.sizeof(B.Config)  # same struct size on both sides of the ABI
    # reference defaults: src/solve.jl:63-76,97-99, src/utils.jl:91-133, NLFunctional()
    assert (cfg.abstol, cfg.reltol) == (1e-6, 1e-3)
    assert (cfg.gamma, cfg.qmin, cfg.qmax, cfg.failfactor, cfg.qoldinit) == (0.9, 0.2, 10.0, 2.0, 1e-4)
    assert cfg.maxiters == 1000000 and cfg.fp_max_iter == 10 and cfg.fp_kappa == 0.01
    assert (cfg.disc_interp_points, cfg.disc_abstol, cfg.disc_reltol) == (10, 1e-12, 0.0)
    L.b2d_config_default(C.byref(cfg), 1)
    assert cfg.qsteady_max == 1.2


def test_model_registry_matches_oracle():
    import oracle_lib as O
    for name, mid in B.MODELS.items():
        n = [C.c_int32() for _ in range(5)]
        assert B.lib().b2d_model_info(mid, *[C.byref(x) for x in n]) == 0
        assert tuple(x.value for x in n[:4]) == O.MODEL_DIMS[mid], name
    assert B.lib().b2d_model_info(99, *[C.byref(C.c_int32()) for _ in range(5)]) != 0
    assert "unknown model" in _lib.last_error()


def test_mass_matrix_of_the_registry_models():
    # DDEFunction(f; mass_matrix) is part of the compiled model (test/interface/mass_matrix.jl:46-62)
    import numpy as np
    for name, mid in B.MODELS.items():
        diag = np.full(3, -1.0)
        kind = B.lib().b2d_model_mass_matrix(mid, diag.ctypes.data, 3)
        if name == "sir_ddae":
            assert kind == 2 and diag.tolist() == [1.0, 1.0, 0.0]        # singular: isdae
        else:
            assert kind == 0, name
    assert B.lib().b2d_model_mass_matrix(99, None, 0) < 0
    # the host mirror derives isdae / neutral from it like DDEProblem does (mass_matrix.jl:58-62)
    dae = B.DDEProblem("sir_ddae", [0.99, 0.01, 0.0], None, (0.0, 40.0), [0.5], constant_lags=[4.0])
    assert dae.isdae and dae.neutral and dae.mass_matrix.tolist() == [1.0, 1.0, 0.0]
    assert not dae.remake(neutral=False).neutral
    ode = B.DDEProblem("sir", [0.99, 0.01, 0.0], None, (0.0, 40.0), [0.5], constant_lags=[4.0])
    assert not ode.isdae and not ode.neutral and ode.mass_matrix is None


def test_problem_validation_errors():
    with pytest.raises(B.B200DDEError):
        B.DDEProblem("no_such_model", [1.0], None, (0.0, 1.0))
    with pytest.raises(B.B200DDEError):
        B.DDEProblem("bc_model", [1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])  # wrong n_state
    with pytest.raises(B.B200DDEError):
        B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0])  # missing constant lag
    with pytest.raises(B.B200DDEError):
        B.MethodOfSteps("RK4")


def test_saveat_grid_semantics():
    # upstream initialize_saveat (src/solve.jl:262): a Number is a step; interior points only
    g = B.saveat_grid(0.1, (0.0, 10.0))
    assert len(g) == 99 and g[0] == 0.1 and g[2] == 0.3 and g[-1] == 9.9   # correctly rounded k/10 like Julia ranges
    assert list(B.saveat_grid(40.0, (0.0, 100.0))) == [40.0, 80.0]          # test/interface/saveat.jl:132-136
    assert list(B.saveat_grid([0.0, 25.0, 100.0], (0.0, 100.0))) == [25.0]
    assert list(B.saveat_grid(0.5, (2.0, 0.0))) == [1.5, 1.0, 0.5]          # reverse time


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    prob = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])
    with pytest.raises(B.B200DDEError, match="no CUDA device|no CPU fallback"):
        B.solve(prob, B.MethodOfSteps(B.Tsit5()))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "delaydiffeq.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle_dde" not in txt, os.path.join(dp, f)


def test_hot_kernels_keep_state_in_registers():
    """Codegen guard: the saveat kernels of the two BASELINE workloads must not have a stack frame or spills.

    A dynamically indexed register array (seen once with save_idxs) silently moves the state vector into local memory
    and costs 3x; ptxas reports it as a stack frame.  lib/ptxas.log is written by every `make`.
    """
    import re
    log = os.path.join(ROOT, "delaydiffeq.jl_b200", "lib", "ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas.log (library not built by make here)")
    text = open(log).read()
    seen = 0
    for m in re.finditer(r"Function properties for (\S*mos_kernelINS_(?:7BcModel|16MackeyGlassModel)ELi0ELi0\S*)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", text):
        seen += 1
        assert int(m.group(2)) == 0 and int(m.group(3)) == 0 and int(m.group(4)) == 0, m.group(0)
        assert int(m.group(5)) <= 128, m.group(0)
    assert seen >= 2
    # the stiff workload's kernel (Rosenbrock23, per-thread LU): pivot exchanges written as `if (p == r) swap` became a
    # dynamically indexed local array (72-byte frame, 58.4 vs 62.5 ms)
    m = re.search(r"Function properties for \S*mos_kernelINS_10StiffModelELi1ELi0\S*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores", text)
    assert m and int(m.group(1)) == 0 and int(m.group(2)) == 0
    # every other kernel: the only things allowed in local memory are the arrays of Traj::Local (large queues, tracked
    # discontinuities, Anderson history).  With those arrays inside Traj the whole trajectory state of the state-dependent
    # and of the general kernels was memory resident (2.7 KB frames; state-dependent sweep 8.6 ms instead of 4.4 ms).
    worst = 0
    for m in re.finditer(r"Function properties for \S*mos_kernelINS_\d+(\w+?)ELi\dELi\dELi\d\S*\n\s*(\d+) bytes stack frame", text):
        # general kernels, one lag: 32-entry propagated queues (1016 bytes of Traj::Local) + up to ~100 bytes of register spills
        limit = 2400 if m.group(1) in ("StateDepModel", "OneDelayDepModel") else 1200
        assert int(m.group(2)) <= limit, m.group(0)
        worst = max(worst, int(m.group(2)))
    assert worst > 0  # the pattern matched the general kernels at all


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's stdout is the driver's interface: exactly one line, valid JSON, with the keys of the contract."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--traj", "2000", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "accepted trajectory-steps/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def _c_prototypes():
    """name -> (return type, [parameter types]) of every function declared in include/b200dde.h (types normalised)."""
    import re
    txt = open(os.path.join(ROOT, "include", "b200dde.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(void|int|const char \*|void \*)\s*(b2d_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\bconst\b", "", a)
                a = re.sub(r"\s*\b[A-Za-z_]\w*\s*(\[\d*\])?$", lambda mm: "*" if mm.group(1) else "", a.strip()) if not a.strip().endswith("*") else a
                params.append(a.replace(" ", ""))
        protos[name] = (ret.replace(" ", ""), params)
    return protos


_JL = {"Ptr{Cvoid}": {"b2d_handle", "void*"}, "Ref{Ptr{Cvoid}}": {"b2d_handle*"}, "Ptr{Ptr{Cvoid}}": {"b2d_handle*"},
       "Ref{B2DConfig}": {"b2d_config*"}, "Int64": {"int64_t"}, "Int32": {"int32_t"}, "Cint": {"int", "int32_t"},
       "Float64": {"double"}, "Ptr{Float64}": {"double*"}, "Ptr{Int32}": {"int32_t*"}, "Ptr{Int64}": {"int64_t*"},
       "Cstring": {"char*", "constchar*"}, "Cvoid": {"void"}, "Ptr{UInt8}": {"uint8_t*"}}


def test_julia_shim_ccalls_match_the_header():
    """The Julia binding cannot run here (no julia in the image), so its interface is checked statically: every ccall of
    julia/B200DDE.jl names a function of include/b200dde.h with the same number of arguments and compatible types, and the
    B2DConfig mirror lists the fields of b2d_config in order with matching widths."""
    import re
    jl = open(os.path.join(ROOT, "delaydiffeq.jl_b200", "julia", "B200DDE.jl")).read()
    protos = _c_prototypes()
    assert {"b2d_solve", "b2d_solve_multi", "b2d_create", "b2d_create_user_model", "b2d_destroy"} <= set(protos)
    calls = re.findall(r"ccall\(\(:(b2d_\w+), lib\),\s*(\w+),\s*\(([^)]*)\)", jl, flags=re.S)
    assert len(calls) >= 12
    seen = set()
    for name, ret, args in calls:
        assert name in protos, name
        seen.add(name)
        cret, cparams = protos[name]
        jargs = [a.strip() for a in " ".join(args.split()).split(",") if a.strip()]
        assert len(jargs) == len(cparams), (name, jargs, cparams)
        assert cret in _JL[ret], (name, ret, cret)
        for ja, ca in zip(jargs, cparams):
            assert ca in _JL[ja], (name, ja, ca)
    # the calls a drop-in needs are all there
    assert {"b2d_config_default", "b2d_create", "b2d_create_user_model", "b2d_destroy", "b2d_release_cached", "b2d_solve",
            "b2d_solve_multi", "b2d_set_tolerances", "b2d_set_save_idxs", "b2d_set_tstops", "b2d_set_d_discontinuities",
            "b2d_model_mass_matrix", "b2d_last_error"} <= seen
    # struct mirror: same field names in the same order, Int32/Int64/Float64 <-> int32_t/int64_t/double
    hdr = open(os.path.join(ROOT, "include", "b200dde.h")).read()
    body = re.sub(r"/\*.*?\*/", " ", hdr[hdr.index("typedef struct b2d_config {"):hdr.index("} b2d_config;")], flags=re.S)
    cfields = []
    for m in re.finditer(r"\b(int32_t|int64_t|double)\s+([^;]+);", body):
        for nm in m.group(2).split(","):
            cfields.append((nm.strip(), m.group(1)))
    jbody = jl[jl.index("mutable struct B2DConfig"):jl.index("B2DConfig() = new()")]
    jfields = re.findall(r"(\w+)::(Int32|Int64|Float64)", jbody)
    width = {"Int32": "int32_t", "Int64": "int64_t", "Float64": "double"}
    assert [(n, width[t]) for n, t in jfields] == cfields
    from b200dde._lib import Config
    assert [f[0] for f in Config._fields_] == [n for n, _ in cfields]


def test_prob_func_may_only_change_u0_p_and_lag_values():
    # the reference solves every prob_func output as a full problem (src/solve.jl:159); on this path everything but u0 and
    # p is shared by the ensemble, and a prob_func that changes it is refused, not silently ignored
    prob = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])
    ok = B.EnsembleProblem(prob, prob_func=lambda pr, i, rep: pr.remake(p=[0.2, 0.3, 1.0 + 0.01 * i]))
    u0, p = ok.matrices(4)
    assert p[3, 2] == 1.03 and u0.shape == (4, 3)
    assert ok.lags_matrix is None
    for change in (dict(tspan=(0.0, 5.0)), dict(neutral=True), dict(order_discontinuity_t0=1)):
        bad = B.EnsembleProblem(prob, prob_func=lambda pr, i, rep, c=change: pr.remake(**c) if i == 2 else pr)
        with pytest.raises(B.B200DDEError, match="only u0, p and constant_lags"):
            bad.matrices(4)
    # the values of the constant lags may vary: they become the per-trajectory lag matrix (b2d_set_trajectory_lags)
    sweep = B.EnsembleProblem(prob, prob_func=lambda pr, i, rep: pr.remake(constant_lags=[1.0 + 0.25 * i]))
    sweep.matrices(4)
    assert sweep.lags_matrix.tolist() == [[1.0], [1.25], [1.5], [1.75]]


def test_callback_objects_fill_the_config_as_the_header_says():
    # the host-side mirror of solve(...; callback = ...): every callback shape lands in the b2d_config fields include/b200dde.h
    # documents (B2D_CB_* / B2D_CBA_*), and the preset-time shapes add their times to tstops as DiffEqCallbacks does
    hdr = open(os.path.join(ROOT, "include", "b200dde.h")).read()
    cst = {k: int(v) for k, v in re.findall(r"#define (B2D_CBA?_\w+) (\d+)", hdr)}
    prob = B.DDEProblem("bc_model", [1.0, 1.0, 1.0], None, (0.0, 10.0), [0.2, 0.3, 1.0], constant_lags=[1.0])

    def cfg_of(cb):
        cfg = B.api._make_config(prob, B.MethodOfSteps(B.Tsit5()), 0, {})
        cb.apply(cfg)
        return cfg
    c = cfg_of(B.PresetTimeCallback(3.3, ("add", 1, 0.25), save_positions=(False, True)))
    assert (c.callback, c.cb_affect, c.cb_idx, c.cb_par0, c.cb_par1, c.cb_save_positions) == (cst["B2D_CB_AT_TIME"], cst["B2D_CBA_ADD"], 1, 3.3, 0.25, 2)
    assert B.PresetTimeCallback(3.3, "negate").tstops((0.0, 10.0)) == [3.3]
    c = cfg_of(B.PeriodicCallback("terminate", 2.5))
    assert (c.callback, c.cb_affect, c.cb_par0) == (cst["B2D_CB_PERIODIC"], cst["B2D_CBA_TERMINATE"], 2.5)
    assert B.PeriodicCallback(None, 2.5).tstops((0.0, 10.0)) == [2.5, 5.0, 7.5, 10.0]
    c = cfg_of(B.ContinuousCallback(("time", 2.6), "negate"))
    assert (c.callback, c.cb_affect, c.cb_par0, c.cb_direction, c.cb_interp_points, c.cb_save_positions) == \
        (cst["B2D_CB_CONT_TIME"], cst["B2D_CBA_NEGATE"], 2.6, 0, 10, 3)
    c = cfg_of(B.ContinuousCallback(("state", 2, 0.8), ("add", 1, 0.25), direction=-1, interp_points=0))
    assert (c.callback, c.cb_cond_idx, c.cb_par0, c.cb_affect, c.cb_idx, c.cb_par1, c.cb_direction, c.cb_interp_points) == \
        (cst["B2D_CB_CONT_STATE"], 2, 0.8, cst["B2D_CBA_ADD"], 1, 0.25, -1, 0)
    c = cfg_of(B.ContinuousCallback("model", par=(1.0, 0.5)))
    assert (c.callback, c.cb_affect, c.cb_par0, c.cb_par1) == (cst["B2D_CB_CONT_MODEL"], cst["B2D_CBA_NONE"], 1.0, 0.5)
    assert B.ContinuousCallback(("time", 2.6), "negate").tstops((0.0, 10.0)) == []
    c = cfg_of(B.DiscreteCallback("model", par=(40.0, 0.3)))
    assert (c.callback, c.cb_par0, c.cb_par1) == (cst["B2D_CB_MODEL"], 40.0, 0.3)
    with pytest.raises(B.B200DDEError, match="unknown condition"):
        cfg_of(B.ContinuousCallback(("velocity", 0, 1.0), "negate"))
