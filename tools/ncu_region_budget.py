#!/usr/bin/env python
"""Per-source-region instruction budget of one kernel from an ncu report.

Input: `ncu -i REPORT --page source --csv --print-source cuda,sass` (the report must have been captured with
`--import-source on` from a `-lineinfo` build).  Every aggregated source-line row (Address "-") carries the executed warp
instructions and the stall samples of the SASS attributed to that line (innermost inlined frame).  Lines are grouped by the
device function that encloses them in the *embedded* source (so the numbers stay valid when the tree moves on), functions
into the regions DESIGN.md talks about.

usage: ncu_region_budget.py both.csv cuda.csv [attempts]
  both.csv  --print-source cuda,sass (metrics per source line; lists only lines that own SASS)
  cuda.csv  --print-source cuda      (the full embedded text of mos_device.cuh: needed to find the enclosing functions)
  attempts  accepted + rejected steps of the launch: adds per-attempt figures
"""
import collections
import csv
import re
import sys

REGIONS = [  # (region, file regex, function regex) — first match wins
    ("history lookup: locate / interval load", r"mos_device", r"^(locate|load_interval|advance_interval|prefetch_interval|R|async_.*|cbase|c_y0|c_k|c_y1|c_thi_next|w_issue|w_leave|wbase|wthi)$"),
    ("history lookup: interpolant (weights + sum)", r"mos_device", r"^(interp|ts5_bweights|ts5_bweights_d|hist|hist_all_slots.*|prefetch_history.*|hpre|H|HPre.*|eval|deriv|w_pass)$"),
    ("model f (inlined RHS)", r"models", r".*"),
    ("stage sums (Tsit5 / RB23 / BS3)", r"mos_device", r"^(perform_step_.*|feval.*|lu_solve)$"),
    ("error estimate + norm", r"mos_device", r"^(error_estimate_tsit5|rms|resid)$"),
    ("controller: loopfooter + fastpower", r"mos_device", r"^(loopfooter)$"),
    ("controller: loopfooter + fastpower", r"detmath", r"(fastpower|exp2|log2|pow)"),
    ("loopheader + check_error + eps", r"mos_device", r"^(loopheader|check_error|timedep_dtmin|handle_discontinuities|add_next_discontinuities|push_tracked|neutral)$"),
    ("divisions: dm_rcp_newton / dm_div_try / dm_div (lookup theta, history map, controller)", r"detmath", r"^(dm_rcp_newton|dm_div_try|dm_div|dm_div_slow)$"),
    ("divisions: dm_rcp_newton / dm_div_try / dm_div (lookup theta, history map, controller)", r"mos_device", r"^(dv|hmap_apply)$"),
    ("loopheader + check_error + eps", r"detmath", r".*"),
    ("savevalues: ring node store", r"mos_device", r"^(savevalues|publish_inflight|publish_inflight_at|force_save)$"),
    ("savevalues: saveat interpolation + output stores", r"mos_device", r"^(interp_step|emit|emit_at|save_src)$"),
    ("queues / tstops / bookkeeping", r"mos_device", r"^(bookkeeping|finish|has_.*|first_.*|pop_.*|uts_.*|udd_.*|user_disc_first|push|pop|top|empty|n|Nr|T|O|clear|gt|sm_int|bind)$"),
    ("fixed point / Anderson / track.jl", r"mos_device", r"^(fixed_point_check|anderson|qradd|qrdelete|givens|track_propagated_discontinuities|discontinuity_function|sgn)$"),
    ("init + auto_dt_reset", r"mos_device", r"^(init|auto_dt_reset|Traj)$"),
    ("attempt glue + warp loop", r"mos_device", r"^(attempt|mos_kernel|SMR|usnap|operator\*|forward)$"),
]


def enclosing_functions(lines):
    """line number -> name of the device function whose body contains it (brace counting on the embedded source)."""
    out, stack, depth = {}, [], 0
    pend = None
    for no in sorted(lines):
        text = lines[no]
        code = text.split("//")[0]
        m = re.search(r"(?:__device__|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*|operator\*)\s*\(", code)
        if m and not code.strip().startswith("#"):
            pend = m.group(1)
        ctor = re.match(r"\s*__device__ __forceinline__ (Traj)\(", code)
        if ctor:
            pend = "Traj"
        for ch in code:
            if ch == "{":
                depth += 1
                if pend is not None:
                    stack.append((pend, depth))
                    pend = None
            elif ch == "}":
                if stack and stack[-1][1] == depth:
                    stack.pop()
                depth -= 1
        out[no] = stack[-1][0] if stack else (pend or "?")
        if stack and stack[-1][1] <= depth:
            out[no] = stack[-1][0]
    return out


def main():
    path = sys.argv[1]
    full_text = {}
    for r in csv.reader(open(sys.argv[2])):
        if len(r) == 2 and r[0].strip().isdigit():
            full_text[int(r[0])] = r[1]
    attempts = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path)))
    per_file = collections.defaultdict(dict)   # file -> line -> (text, inst, samples, thread_inst)
    cur, hdr = None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            # the header has two "Source" columns; index by position
            i_inst, i_samp, i_tinst = r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
            continue
        if cur is None or hdr is None or len(r) < 8 or not r[0].strip().isdigit() or r[2] != "-":
            continue
        per_file[cur][int(r[0])] = (r[1], int(r[i_inst] or 0), int(r[i_samp] or 0), int(r[i_tinst] or 0))
    agg = collections.OrderedDict((name, [0, 0, 0]) for name, _, _ in REGIONS)
    agg["other"] = [0, 0, 0]
    other_lines = []
    for f, lines in per_file.items():
        # function map needs every line of the file; the csv only lists lines with code attributed -> read text we have
        if "mos_device" in f:
            fmap = enclosing_functions(full_text)
        else:
            try:
                fmap = enclosing_functions({i + 1: t.rstrip("\n") for i, t in enumerate(open(f))})
            except OSError:
                fmap = {}
        for no, (text, inst, samp, tinst) in lines.items():
            fn = fmap.get(no, "?")
            for name, fre, fnre in REGIONS:
                if re.search(fre, f) and re.search(fnre, fn):
                    a = agg[name]
                    break
            else:
                a = agg["other"]
                other_lines.append((inst, f.split("/")[-1], no, fn, text.strip()[:70]))
            a[0] += inst; a[1] += samp; a[2] += tinst
    ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values()); tt = sum(a[2] for a in agg.values())
    print(f"total warp instructions {ti:,}   thread instructions {tt:,}   stall samples {ts:,}")
    if attempts:
        print(f"attempts {attempts:,.0f}: {tt / attempts:.0f} thread instructions per attempt")
    print(f"{'region':88s} {'warp inst':>14s} {'%inst':>6s} {'%samples':>8s}" + ("  thread-inst/attempt" if attempts else ""))
    for name, (i, s, t) in agg.items():
        if i == 0 and s == 0:
            continue
        extra = f"  {t / attempts:8.0f}" if attempts else ""
        print(f"{name:88s} {i:14,d} {100 * i / ti:6.1f} {100 * s / max(ts, 1):8.1f}{extra}")
    if other_lines:
        print("\nlargest unclassified lines:")
        for inst, f, no, fn, text in sorted(other_lines, reverse=True)[:12]:
            print(f"  {inst:12,d} {f}:{no} [{fn}] {text}")


if __name__ == "__main__":
    main()
