#!/usr/bin/env python
"""Host-I/O ceiling of the end-to-end path on an N-GPU box, one process driving every GPU:
aggregate device->host rate of N concurrent copies into (a) cudaHostAlloc pinned memory, (b) transparent-hugepage-backed
memory registered with cudaHostRegister, (c) MAP_HUGETLB memory when the box has hugetlb pages, (d) write-combined pinned
memory, (e) the same bytes in 8 pieces per GPU; then the library's own single-process path, b2d_solve_multi with one handle
per GPU writing into one pinned array (bc_model, 10^6 trajectories per GPU).  Diagnostic, not product."""
import ctypes as C
import mmap
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200")); sys.path.insert(0, ROOT)
ng = torch.cuda.device_count()
NB = 1 << 30
rt = torch.cuda.cudart()
libc = C.CDLL(None, use_errno=True)
print("gpus", ng, "| hugetlb:", [l.strip() for l in open("/proc/meminfo") if l.startswith(("HugePages_Total", "Hugepagesize"))],
      "| THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
devs = [torch.full((NB,), 7, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
streams = [torch.cuda.Stream(device=g) for g in range(ng)]


def rate(hosts, pieces=1, reps=4):
    def go():
        for g, h in enumerate(hosts):
            with torch.cuda.stream(streams[g]):
                step = NB // pieces
                for k in range(pieces):
                    h[k * step:(k + 1) * step].copy_(devs[g][k * step:(k + 1) * step], non_blocking=True)
    go()
    for g in range(ng):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    for g in range(ng):
        torch.cuda.synchronize(g)
    return reps * NB * len(hosts) / (time.perf_counter() - t0) / 1e9


def registered(flags_mmap, advise):
    out = []
    for g in range(ng):
        try:
            m = mmap.mmap(-1, NB, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | flags_mmap)
        except OSError as e:
            print("   mmap failed:", e)
            return None
        arr = np.frombuffer(m, dtype=np.uint8)
        if advise:
            r = libc.madvise(C.c_void_p(arr.ctypes.data), C.c_size_t(NB), 14)  # MADV_HUGEPAGE
            if r != 0:
                print("   madvise(MADV_HUGEPAGE) failed, errno", C.get_errno())
        arr[::4096] = 1  # fault the pages in
        err = rt.cudaHostRegister(arr.ctypes.data, NB, 0)
        if int(err) != 0:
            print("   cudaHostRegister failed:", err)
            return None
        t = torch.from_numpy(arr)
        t._keep = m
        out.append(t)
    return out


pinned = [torch.empty(NB, dtype=torch.uint8).pin_memory() for _ in range(ng)]
print(f"(a) cudaHostAlloc pinned, one GPU alone:            {rate(pinned[:1]):7.1f} GB/s")
print(f"(a) cudaHostAlloc pinned, {ng} GPUs concurrently:       {rate(pinned):7.1f} GB/s aggregate")
print(f"(e) same, 8 pieces per GPU:                          {rate(pinned, pieces=8):7.1f} GB/s aggregate")
thp = registered(0, True)
if thp:
    print(f"(b) THP-backed + cudaHostRegister, {ng} GPUs:           {rate(thp):7.1f} GB/s aggregate")
    anon = [l for l in open("/proc/meminfo") if l.startswith("AnonHugePages")]
    print("    ", anon[0].strip() if anon else "")
plain = registered(0, False)
if plain:
    print(f"(b') 4 KB pages + cudaHostRegister, {ng} GPUs:          {rate(plain):7.1f} GB/s aggregate")
MAP_HUGETLB = 0x40000
htlb = registered(MAP_HUGETLB, False)
if htlb:
    print(f"(c) MAP_HUGETLB + cudaHostRegister, {ng} GPUs:          {rate(htlb):7.1f} GB/s aggregate")
# (d) write-combined pinned memory through the runtime
wc = []
for g in range(ng):
    ptr = C.c_void_p()
    lib = C.CDLL("libcudart.so.12")
    if lib.cudaHostAlloc(C.byref(ptr), C.c_size_t(NB), 4) != 0:  # cudaHostAllocWriteCombined
        wc = None
        break
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(NB,))
    wc.append(torch.from_numpy(arr))
if wc:
    print(f"(d) write-combined pinned, {ng} GPUs:                   {rate(wc):7.1f} GB/s aggregate")

# ---- the library's single-process multi-GPU path ------------------------------------------------------------------
import b200dde as B
import bench
from b200dde.api import _make_config, solve_multi
n = 1_000_000
wl = bench.workload("bc", n * ng, 0, 1)
prob = B.DDEProblem(wl["model"], wl["u0"][0], None, wl["tspan"], wl["p"][0], constant_lags=wl["lags"])
alg = B.MethodOfSteps(B.Tsit5())
solvers = [B.Solver(_make_config(prob, alg, g, {})) for g in range(ng)]
grid = B.saveat_grid(wl["saveat"], wl["tspan"])
L = B.lib()
nbytes = n * ng * 101 * 3 * 8
ptr = L.b2d_host_alloc(nbytes)
h_out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n * ng, 101, 3))
rc = np.zeros(n * ng, dtype=np.int32); st = np.zeros((n * ng, 8), dtype=np.int64); out_t = np.zeros(101)
hs = (C.c_void_p * ng)(*[s.h for s in solvers])
lags = np.ascontiguousarray(wl["lags"])


def multi():
    r = L.b2d_solve_multi(hs, ng, n * ng, 0.0, 10.0, wl["u0"].ctypes.data, wl["p"].ctypes.data, lags.ctypes.data, grid.ctypes.data, len(grid),
                          1, 1, h_out.ctypes.data, out_t.ctypes.data, rc.ctypes.data, st.ctypes.data)
    B._lib.check(r)


multi(); multi()
t0 = time.perf_counter()
for _ in range(3):
    multi()
dt = (time.perf_counter() - t0) / 3
print(f"b2d_solve_multi, one process, {ng} handles, bc_model {n} trajectories per GPU, pinned output: {dt * 1e3:.1f} ms per solve, "
      f"{st[:, 0].sum() / dt:.3e} steps/s, {nbytes / dt / 1e9:.1f} GB/s of output, all Success: {bool((rc == 1).all())}")
