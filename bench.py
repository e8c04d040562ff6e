#!/usr/bin/env python
"""bench.py — accepted trajectory-steps/s of the MethodOfSteps ensemble path (BASELINE.json metric).

One "step" = one pass of the hot path over one batch: solving the whole bc_model ensemble of
BASELINE.json configs[1] (10^6 trajectories per GPU, 100x100x100 sweep over (p0,q0,v0), FP64, Tsit5,
abstol 1e-6 / reltol 1e-3, saveat = 0.1 -> 101 points) from t = 0 to t = 10.

  value      Σ naccept / time, inputs resident in HBM, kernel launched through b2d_solve_device
  e2e        same through b2d_solve with pinned HOST buffers (H2D of u0/p and D2H of the saveat block inside the region)
  roofline   algorithmic bytes of SURVEY §8(d) per launch / CUDA-event duration of the launch, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement proxy for EnsembleThreads) on the host cores, bounded sample

`--impl reference` times the CPU oracle port instead (the reference is Julia and cannot run here).
N > 1: one rank per GPU (torchrun), trajectories sharded with no collective in the data path (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))

METRIC = "accepted trajectory-steps/sec"
UNIT = "steps/s"
T0, TF, SAVEAT = 0.0, 10.0, 0.1
# node write 8(1+8n) + two delayed-node reads; "sd" uses the scalar Tsit5 figure, "stiff" the same formula
# with Rosenbrock23's 2 slopes per node (8(1+3n) + 2*8(2+3n_h)); the last two are this repo's estimates, not SURVEY's
ALG_BYTES_PER_STEP = {"bc": 360.0, "mg": 232.0, "sd": 232.0, "stiff": 256.0}
ALG_FLOPS_PER_ATTEMPT = {"bc": 900.0, "mg": 650.0, "sd": 650.0, "stiff": 500.0}  # SURVEY §8(d) / estimates as above


def workload(name, n_traj, rank, world):
    """Deterministic parameter sweeps of SURVEY §8(d); the sweep is extended along its slowest axis for world > 1
    (weak scaling) and each rank takes a contiguous slice."""
    n_total = n_traj * world
    i = np.arange(rank * n_traj, (rank + 1) * n_traj, dtype=np.int64)
    if name == "bc":
        side = 100
        na = max(1, n_total // (side * side))
        a, b, c = i // (side * side), (i // side) % side, i % side
        p = np.stack([0.1 + 0.2 * a / max(na - 1, 1), 0.2 + 0.2 * b / (side - 1), 0.5 + 1.0 * c / (side - 1)], axis=1)
        u0 = np.ones((n_traj, 3))
        return dict(model="bc_model", u0=u0, p=np.ascontiguousarray(p), lags=[1.0], tspan=(0.0, 10.0), saveat=0.1,
                    desc=f"bc_model ensemble (BASELINE configs[1]), {n_traj} trajectories/GPU, grid {na}x100x100 over (p0,q0,v0)")
    if name == "mg":
        nb = 4000
        na = max(1, n_total // nb)
        a, b = i // nb, i % nb
        p = np.stack([0.18 + 0.04 * a / max(na - 1, 1), 0.5 + 1.0 * b / (nb - 1)], axis=1)
        u0 = p[:, 1:2].copy()
        return dict(model="mackey_glass", u0=u0, p=np.ascontiguousarray(p), lags=[17.0], tspan=(0.0, 1000.0), saveat=10.0,
                    desc=f"Mackey-Glass tau=17 (BASELINE configs[2]), {n_traj} trajectories/GPU, grid {na}x4000 over (beta,h0)")
    if name == "sd":
        u0 = (0.5 + 1.0 * i / max(n_total - 1, 1))[:, None].copy()
        return dict(model="state_dependent", u0=u0, p=np.zeros((n_traj, 1)), lags=[], tspan=(0.0, 10.0), saveat=0.1,
                    desc=f"state-dependent lag u'(t) = -u(t - 0.5 - 0.1|u(t)|) (BASELINE configs[3]), {n_traj} trajectories/GPU, u0 sweep")
    if name == "stiff":
        lam = 10.0 ** (2.0 + 3.0 * i / max(n_total - 1, 1))
        return dict(model="stiff_kinetics", u0=np.tile([2.0, 0.5, 1.0], (n_traj, 1)), p=lam[:, None].copy(), lags=[1.0],
                    tspan=(0.0, 10.0), saveat=0.1, alg="Rosenbrock23",
                    desc=f"stiff delay kinetics under Rosenbrock23 (BASELINE configs[4]), {n_traj} trajectories/GPU, lambda 1e2..1e5")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        return self.__enter__()

    def stop(self):
        self.__exit__()

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)
            self.proc = None

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        smax = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(name, n_traj):
    """dram__bytes_read.sum + dram__bytes_write.sum of one ncu-captured launch (10^6 trajectories), scaled to this launch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            v = json.load(f).get(name)
        return None if v is None else v * n_traj / 1.0e6
    return None


def cpu_oracle_rate(wl, n_sample, threads):
    """Accepted steps/s of the CPU oracle on an evenly strided sample of the workload (restatement proxy for
    EnsembleThreads: static contiguous chunks over `threads` std::threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from b200dde import saveat_grid
    model = {"bc_model": O.MODEL_BC, "mackey_glass": O.MODEL_MACKEY_GLASS, "state_dependent": O.MODEL_STATE_DEP,
             "stiff_kinetics": O.MODEL_STIFF}[wl["model"]]
    n = wl["u0"].shape[0]
    idx = np.linspace(0, n - 1, n_sample).astype(np.int64)
    cfg = O.default_config(alg=O.ALG_ROSENBROCK23 if wl.get("alg") == "Rosenbrock23" else O.ALG_TSIT5, model=model, fp_div_rule=1)
    grid = saveat_grid(wl["saveat"], wl["tspan"])
    u0s, ps = np.ascontiguousarray(wl["u0"][idx]), np.ascontiguousarray(wl["p"][idx])
    out = np.zeros((n_sample, len(grid) + 2, u0s.shape[1]))  # allocated and touched outside the timed region
    tic = time.perf_counter()
    r = O.solve(cfg, u0s, ps, wl["lags"], wl["tspan"][0], wl["tspan"][1], saveat=grid, nthreads=threads, out=out)
    sec = time.perf_counter() - tic
    return float(r["stats"][:, 0].sum()) / sec, sec


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Julia and no julia binary
    exists in the image, so the timed code is the KAT-pinned oracle port with every host thread."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    wl = workload(args.workload, args.traj, 0, 1)
    threads = O.hardware_threads()
    # the whole workload per step: 10^6 bc_model trajectories take the oracle ~1 s on 16 threads, so no sampling is
    # needed to stay within minutes (a 20000-trajectory sample lasts 17 ms and mostly measures thread start-up)
    n_sample = min(args.cpu_sample or args.traj, args.traj)
    for _ in range(args.warmup):
        cpu_oracle_rate(wl, max(256, n_sample // 8), threads)
    rates, secs = [], []
    for _ in range(args.steps):
        r, s = cpu_oracle_rate(wl, n_sample, threads)
        rates.append(r)
        secs.append(s)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "sample": f"{n_sample} of {args.traj} trajectories per step (evenly strided)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_sample} of {args.traj} trajectories per step, all of t in [0,10]"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_json(line)


_JSON_OUT = None


def emit_json(line):
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def shard_parity(wl, d_out, d_stats, d_rc, grid, n_sample):
    """A strided sample of THIS rank's shard against the CPU oracle, bit for bit (saveat block, statistics, return codes).
    The oracle is the checker here, never the thing measured.  Returns (ok, sample size)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import torch
    model = {"bc_model": O.MODEL_BC, "mackey_glass": O.MODEL_MACKEY_GLASS, "state_dependent": O.MODEL_STATE_DEP,
             "stiff_kinetics": O.MODEL_STIFF}[wl["model"]]
    n = wl["u0"].shape[0]
    idx = np.unique(np.linspace(0, n - 1, n_sample).astype(np.int64))
    cfg = O.default_config(alg=O.ALG_ROSENBROCK23 if wl.get("alg") == "Rosenbrock23" else O.ALG_TSIT5, model=model, fp_div_rule=1)
    ref = O.solve(cfg, np.ascontiguousarray(wl["u0"][idx]), np.ascontiguousarray(wl["p"][idx]), wl["lags"], wl["tspan"][0],
                  wl["tspan"][1], saveat=grid, nthreads=min(8, O.hardware_threads()))
    ti = torch.from_numpy(idx).to(d_out.device)
    got_u, got_st, got_rc = d_out[ti].cpu().numpy(), d_stats[ti].cpu().numpy(), d_rc[ti].cpu().numpy()
    ok = bool(np.array_equal(got_u, ref["u"], equal_nan=True) and np.array_equal(got_st[:, :6], ref["stats"][:, :6]) and
              np.array_equal(got_rc, ref["retcode"]))
    return ok, int(len(idx))


def run_workload(args, name, n, rank, world, local, dev, torch, dist, B, _lib, full):
    """All legs of one workload: device-resident throughput (`value`), roofline, shard parity, end to end through the C
    ABI with host buffers (+ the in-run host-I/O ceiling), the gather of the saveat blocks at N > 1 and, when `full`, the
    cold call, the pageable-output call and the CPU baseline.  Returns the dict of results (rank 0 prints it)."""
    from b200dde.api import _make_config
    wl = workload(name, n, rank, world)
    nstate, npar = wl["u0"].shape[1], wl["p"].shape[1]
    grid = B.saveat_grid(wl["saveat"], wl["tspan"])
    n_out = len(grid) + 2
    prob = B.DDEProblem(wl["model"], wl["u0"][0], None, wl["tspan"], wl["p"][0], constant_lags=wl["lags"])
    alg = B.MethodOfSteps(B.Rosenbrock23() if wl.get("alg") == "Rosenbrock23" else B.Tsit5())
    solver = B.Solver(_make_config(prob, alg, local, {}))

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident path ------------------------------------------------------------------------------
    d_u0 = torch.from_numpy(wl["u0"]).to(dev)
    d_p = torch.from_numpy(wl["p"]).to(dev)
    d_out = torch.empty((n, n_out, nstate), dtype=torch.float64, device=dev)
    d_rc = torch.empty(n, dtype=torch.int32, device=dev)
    d_stats = torch.empty((n, B._lib.N_STATS), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def device_step():
        solver.solve_device(d_u0.data_ptr(), d_p.data_ptr(), wl["lags"], wl["tspan"], grid, True, True, d_out.data_ptr(),
                            d_rc.data_ptr(), d_stats.data_ptr(), n, stream.cuda_stream)

    for _ in range(args.warmup):
        device_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the device-resident region lasts only K x a few ms, shorter than nvidia-smi's sampling period, so the sampler keeps
    # running through the end-to-end region below (same kernels, same clocks) and is stopped after it
    clocks = ClockSampler(local)
    clocks.start()
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    last_kernel_ms, _, ring_cap = solver.last_timing()  # library events around the last launch
    acc = int(d_stats[:, 0].sum().item())
    rej = int(d_stats[:, 1].sum().item())
    nf = int(d_stats[:, 2].sum().item())
    ok = int((d_rc == 1).sum().item())
    total_ms_max = allmax(total_ms)
    acc_all = allsum(acc)
    ms_per_step = total_ms_max / args.steps
    value = acc_all / (ms_per_step * 1e-3)
    res = {"workload": wl["desc"], "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "gpu_launches": args.steps,
           "accepted_steps_per_step": acc, "accepted_steps_all_ranks": int(acc_all), "rejected": rej, "nf": nf, "success": ok,
           "trajectories_per_gpu": n, "ring_capacity": ring_cap, "n_out": n_out,
           "alg": f"MethodOfSteps({wl.get('alg', 'Tsit5')}())", "saveat": wl["saveat"]}

    # ---- parity of this rank's shard against the oracle (every rank checks its own trajectories) -----------------
    par_ok, par_n = shard_parity(wl, d_out, d_stats, d_rc, grid, 256 if name in ("bc", "sd") else 96)
    all_ok = allsum(1.0 if par_ok else 0.0) == float(world)
    res["shard_parity"] = bool(all_ok)
    res["shard_parity_sample"] = f"{par_n} evenly strided trajectories of every rank's shard, bit-exact vs the CPU oracle (saveat block, statistics, retcodes)"
    res["all_success"] = bool(allsum(ok) == float(n) * world)

    # ---- roofline of the dominant (only) kernel --------------------------------------------------------------
    peak, peak_src = peaks()
    alg_bytes = ALG_BYTES_PER_STEP[name] * acc + 8.0 * nstate * n_out * n + 8.0 * (nstate + npar) * n
    kdur_ms = total_ms / args.steps  # one launch per step; events on the launching stream bracket only launches
    achieved = alg_bytes / (kdur_ms * 1e-3) / 1e9
    fp64 = C.c_double(0.0)
    _lib.lib().b2d_measure_fp64_tflops(local, C.byref(fp64))
    flops = ALG_FLOPS_PER_ATTEMPT[name] * (acc + rej)
    fp64_ach = flops / (kdur_ms * 1e-3) / 1e12
    traffic = traffic_from_profile(name, n)
    res["roofline"] = {
        "bound": "fp64", "achieved": fp64_ach, "peak": fp64.value, "unit": "TFLOP/s",
        "frac": fp64_ach / fp64.value if fp64.value > 0 else None, "traffic": traffic,
        "peak_source": "FP64 FMA microbenchmark run in this process (b2d_measure_fp64_tflops); MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_attempt": ALG_FLOPS_PER_ATTEMPT[name], "attempts_per_launch": acc + rej, "kernel_ms": kdur_ms,
        "hbm": {"algorithmic_gbs": achieved, "peak_gbs": peak, "algorithmic_frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "dram_gbs": None if traffic is None else traffic / (kdur_ms * 1e-3) / 1e9,
                "dram_frac": None if traffic is None else traffic / (kdur_ms * 1e-3) / 1e9 / peak,
                "note": "SURVEY §8(d) algorithmic bytes; the history ring is L2-resident by design, measured DRAM traffic "
                        "(ncu, profiles/traffic.json) is what dram_frac uses"},
        "note": "FP64 issue/latency bound (ncu: profiles/); algorithmic flops of SURVEY §8(d) per attempt over the measured FMA peak"}

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        L = _lib.lib()
        nbytes_out = n * n_out * nstate * 8
        ptr = L.b2d_host_alloc(nbytes_out)
        if not ptr:
            raise SystemExit("b2d_host_alloc failed")
        h_out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n, n_out, nstate))
        pu = L.b2d_host_alloc(n * nstate * 8)
        pp = L.b2d_host_alloc(n * npar * 8)
        h_u0 = np.ctypeslib.as_array(C.cast(pu, C.POINTER(C.c_double)), shape=(n, nstate))
        h_p = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double)), shape=(n, npar))
        h_u0[:] = wl["u0"]
        h_p[:] = wl["p"]
        # the ceiling of this leg, measured in this run: every rank copies its saveat block device -> pinned host at the
        # same time (barrier-bracketed, best of 3), nothing else running
        t_h = torch.from_numpy(h_out)
        t_h.copy_(d_out)
        best = float("inf")
        for _ in range(3):
            barrier()
            tic = time.perf_counter()
            t_h.copy_(d_out)
            torch.cuda.synchronize()
            best = min(best, allmax(time.perf_counter() - tic))
        ceiling_gbs = world * nbytes_out / best / 1e9
        h_out[:] = 0.0
        e2e_solver = B.Solver(_make_config(prob, alg, local, {}))
        # caller-owned result arrays, reused by every step like the saveat block (ordinary pageable memory)
        rc_h = np.zeros(n, dtype=np.int32)
        st_h = np.zeros((n, B._lib.N_STATS), dtype=np.int64)
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_solver.solve_host(h_u0, h_p, wl["lags"], wl["tspan"], grid, True, True, out=h_out, rc=rc_h, stats=st_h)
        barrier()
        st_h[:, 0] = -1  # the timed steps must deliver the statistics themselves
        tic = time.perf_counter()
        for _ in range(args.steps):
            e2e_solver.solve_host(h_u0, h_p, wl["lags"], wl["tspan"], grid, True, True, out=h_out, rc=rc_h, stats=st_h)
        torch.cuda.synchronize()
        sec = time.perf_counter() - tic
        # like the reference arm, the metric's numerator is reduced outside the timed region: every step solves the same
        # ensemble, so it is the accepted-step count of the last step times the number of steps (checked against the device path)
        acc_e2e = int(st_h[:, 0].sum()) * args.steps
        if int(st_h[:, 0].sum()) != acc or int((rc_h == 1).sum()) != ok:
            raise SystemExit(f"bench.py [{name}, rank {rank}]: end-to-end statistics differ from the device-resident path: accepted "
                             f"{int(st_h[:, 0].sum())} vs {acc}, Success {int((rc_h == 1).sum())} vs {ok} of {n} "
                             f"(return codes of the device path: {torch.bincount(d_rc.long()).tolist()})")
        sec_max = allmax(sec)
        e2e = {"value": allsum(acc_e2e) / sec_max, "unit": UNIT,
               "h2d_bytes_per_step": int(n * (nstate + npar) * 8),
               "d2h_bytes_per_step": int(nbytes_out + n * 4 + n * B._lib.N_STATS * 8),
               "ms_per_step": 1e3 * sec_max / args.steps,
               "api": "b2d_solve (C ABI) on pinned host buffers, waves of one resident grid of trajectories, three waves in flight, D2H overlapped with the next waves"}
        io_gbs = world * (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["host_io"] = {"achieved_gbs": io_gbs, "ceiling_gbs": ceiling_gbs, "frac": io_gbs / ceiling_gbs,
                          "scope": "all ranks together (the byte counts above are per rank)",
                          "ceiling_source": f"measured in this run: {world} rank(s) copying their {nbytes_out / 1e9:.2f} GB saveat block "
                                            "device -> pinned host at the same time, best of 3"}
        e2e["matches_device_path"] = bool(np.array_equal(h_out, d_out.cpu().numpy()))
        if full:
            # the cold call: create + solve + destroy, what a host that does not keep the handle pays (streams, device and
            # pinned staging allocations, the pilot launch that sizes the ring); rank-local, not part of `value`
            def one_cold_call():
                barrier()
                tic = time.perf_counter()
                cs = B.Solver(_make_config(prob, alg, local, {}))
                cs.solve_host(h_u0, h_p, wl["lags"], wl["tspan"], grid, True, True, out=h_out, rc=rc_h, stats=st_h)
                cs.close()
                return time.perf_counter() - tic
            cold = [one_cold_call() for _ in range(2)]
            os.environ["B2D_NO_POOL"] = "1"  # what the very first call of a process pays: no parked slots to take over
            try:
                first = one_cold_call()
            finally:
                del os.environ["B2D_NO_POOL"]
            cold_max = allmax(min(cold))
            e2e["cold"] = {"value": acc_all / cold_max, "unit": UNIT, "ms_per_call": 1e3 * cold_max,
                           "first_call_ms": 1e3 * allmax(first),
                           "what": "b2d_create + b2d_solve + b2d_destroy around every solve (pinned host buffers), best of 2; the "
                                   "library parks a destroyed handle's streams and buffers for the next create, first_call_ms is "
                                   "the same call without that (every allocation paid)"}
        # the same call with an ordinary (pageable) output array, as a Julia Array would be: the library stages it
        # through pinned buffers and copies with host threads while the next wave runs (single-GPU runs only: N ranks
        # would hold N more 2.4 GB host arrays and compete for the same host cores)
        if world == 1 and full:
            pg_out = np.empty((n, n_out, nstate))
            e2e_solver.solve_host(wl["u0"], wl["p"], wl["lags"], wl["tspan"], grid, True, True, out=pg_out)
            tic = time.perf_counter()
            for _ in range(2):
                _, _, _, st_pg = e2e_solver.solve_host(wl["u0"], wl["p"], wl["lags"], wl["tspan"], grid, True, True, out=pg_out)
            e2e["pageable_value"] = 2.0 * float(st_pg[:, 0].sum()) / (time.perf_counter() - tic)
            del pg_out
        e2e_solver.close()
        del t_h
        for q in (ptr, pu, pp):
            L.b2d_host_free(q)
    clocks.stop()
    res["clocks"] = clocks.summary()
    res["e2e"] = e2e

    # ---- the only collective of the path: the library's NCCL gather of the saveat blocks to rank 0 -------------------
    # (b2d_gather_saveat: one grouped ncclRecv per peer on the root, one ncclSend elsewhere; not part of `value`, which
    # has no collective in it — the block is also what a host-buffer solve delivers without any gather)
    if world > 1 and not args.no_gather:
        from b200dde.distributed import init_library_comm, gather_saveat_library
        init_library_comm(solver)
        full_block = torch.empty((n * world, n_out, nstate), dtype=torch.float64, device=dev) if rank == 0 else None
        gather_saveat_library(solver, d_out, n * world, out=full_block)  # warm-up: NCCL sets its channels up lazily
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        g0.record()
        for _ in range(reps):
            gather_saveat_library(solver, d_out, n * world, out=full_block)
        g1.record()
        torch.cuda.synchronize()
        res["gather_ms"] = allmax(g0.elapsed_time(g1) / reps)
        res["gather_bytes"] = int((world - 1) * d_out.numel() * 8)
        res["gather_gbs"] = res["gather_bytes"] / (res["gather_ms"] * 1e-3) / 1e9
        res["gather_api"] = "b2d_gather_saveat (library, NCCL send/recv over NVLink, dlopen of the process's libnccl.so.2)"
        res["value_with_gather"] = acc_all / ((ms_per_step + res["gather_ms"]) * 1e-3)
        # every rank's block must arrive where its trajectories belong: per-shard checksums on the root against the senders'
        mine = torch.nan_to_num(d_out).sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            got = [torch.nan_to_num(full_block[r * n:(r + 1) * n]).sum() for r in range(world)]
            res["gather_ok"] = bool(all(bool(torch.equal(g.reshape(1), s_)) for g, s_ in zip(got, sums)) and
                                    torch.equal(full_block[:n], d_out))
        del full_block

    # ---- CPU baseline on rank 0 ---------------------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        threads = O.hardware_threads()
        pilot_rate, _ = cpu_oracle_rate(wl, 2048, threads)
        steps_per_traj = acc / n
        budget_s = 12.0 if full else 6.0
        n_sample = args.cpu_sample or int(min(n, max(4096, budget_s * pilot_rate / max(steps_per_traj, 1.0))))
        rate, sec = cpu_oracle_rate(wl, n_sample, threads)
        res["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n_sample} evenly strided trajectories of the same ensemble, full tspan, {sec:.1f} s; "
                                         "C++ oracle = restatement proxy for EnsembleThreads (no Julia in the image)"}
    else:
        res["cpu_baseline"] = None
    solver.close()
    del d_out, d_u0, d_p, d_rc, d_stats
    torch.cuda.empty_cache()
    return res


def main():
    # stdout carries exactly ONE line, the JSON result: everything else that libraries write to file descriptor 1 (NCCL
    # prints "NCCL version ..." there when NCCL_DEBUG=VERSION is set in the environment) is sent to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bc", choices=["bc", "mg", "sd", "stiff"])
    ap.add_argument("--traj", type=int, default=1_000_000, help="trajectories per GPU")
    ap.add_argument("--secondary", default="auto", choices=["auto", "none", "bc", "mg", "sd", "stiff"],
                    help="second workload reported under \"secondary\" (auto: Mackey-Glass next to the default bc_model line)")
    ap.add_argument("--secondary-traj", type=int, default=1_250_000,
                    help="trajectories per GPU of the secondary workload (1.25e6 x 8 GPUs = the 10^7 of BASELINE configs[2])")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip timing the NCCL gather of the saveat blocks at N > 1")
    ap.add_argument("--gather", action="store_true", help="(kept for compatibility: the gather is always timed at N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import b200dde as B
    from b200dde import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    prim = run_workload(args, args.workload, args.traj, rank, world, local, dev, torch, dist, B, _lib, full=True)
    second = args.secondary
    if second == "auto":
        second = "mg" if args.workload == "bc" else "none"
    sec = None
    if second != "none" and second != args.workload:
        sec = run_workload(args, second, args.secondary_traj, rank, world, local, dev, torch, dist, B, _lib, full=False)

    if rank == 0:
        line = {"metric": METRIC, "value": prim["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": prim["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": prim["workload"], "alg": prim["alg"], "abstol": 1e-6, "reltol": 1e-3,
                           "saveat": prim["saveat"], "n_out": prim["n_out"], "trajectories_per_gpu": prim["trajectories_per_gpu"],
                           "ring_capacity": prim["ring_capacity"],
                           "l2": f"per-step working set {(prim['roofline']['hbm']['algorithmic_bytes_per_launch'] / 1e9):.2f} GB of inputs+outputs > 126 MB L2, no flush needed",
                           "accepted_steps_per_step": prim["accepted_steps_per_step"], "rejected": prim["rejected"], "nf": prim["nf"],
                           "success": prim["success"]},
                "clocks": prim["clocks"], "gpu_launches": prim["gpu_launches"] + (sec["gpu_launches"] if sec else 0),
                "e2e": prim["e2e"], "roofline": prim["roofline"], "cpu_baseline": prim["cpu_baseline"],
                "shard_parity": prim["shard_parity"] and (sec["shard_parity"] if sec else True),
                "shard_parity_sample": prim["shard_parity_sample"], "all_success": prim["all_success"],
                "accepted_steps_checksum": prim["accepted_steps_all_ranks"]}
        for k in ("gather_ms", "gather_bytes", "gather_ok"):
            if k in prim:
                line[k] = prim[k]
        if sec is not None:
            line["secondary"] = sec
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
