#!/usr/bin/env python
"""Host topology probe for the multi-GPU end-to-end path: where do the GPUs hang, and does the NUMA node of the pinned
output buffer matter for the device->host copy rate when several GPUs copy at the same time?  (diagnostic, not product)"""
import ctypes, glob, os, subprocess, sys, time
import torch

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node=None):
    if node is None:
        r = libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)
    return r, ctypes.get_errno()


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa
        return f"<{e}>"


print("== nodes", sh("ls -d /sys/devices/system/node/node* | xargs -n1 basename | tr '\\n' ' '"))
for nd in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(os.path.basename(nd), "cpus", open(nd + "/cpulist").read().strip(), "|", sh(f"grep -E 'MemTotal|MemFree' {nd}/meminfo | tr '\\n' ' '"))
print("== affinity", sorted(os.sched_getaffinity(0)))
print("== cpuset.mems", sh("cat /sys/fs/cgroup/cpuset.mems.effective /sys/fs/cgroup/cpuset/cpuset.mems 2>/dev/null"))
print("== cpuset.cpus", sh("cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset/cpuset.cpus 2>/dev/null"))
print("== Mems_allowed", sh("grep -E 'Mems_allowed_list|Cpus_allowed_list' /proc/self/status"))
print("== topo\n" + sh("nvidia-smi topo -m"))
ng = torch.cuda.device_count()
print("== gpus", ng)
nodes = sorted(int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node*"))
for g in range(ng):
    bus = torch.cuda.get_device_properties(g).pci_bus_id if hasattr(torch.cuda.get_device_properties(g), "pci_bus_id") else None
    q = sh(f"nvidia-smi -i {g} --query-gpu=pci.bus_id --format=csv,noheader")
    bid = q.lower().replace("00000000:", "0000:")
    nn = sh(f"cat /sys/bus/pci/devices/{bid}/numa_node")
    print(f"gpu {g} bus {q} numa_node {nn} link", sh(f"nvidia-smi -i {g} --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv,noheader"))

NB = 1 << 30  # 1 GiB per copy
devs = [torch.empty(NB, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
streams = [torch.cuda.Stream(device=g) for g in range(ng)]


def pinned_on(node):
    r = set_mempolicy(MPOL_BIND, node) if node is not None else set_mempolicy(MPOL_DEFAULT)
    try:
        t = torch.empty(NB, dtype=torch.uint8).pin_memory()
    except Exception as e:  # noqa
        t = None
        print("   alloc failed on node", node, e)
    set_mempolicy(MPOL_DEFAULT)
    return t, r


def rate(pairs, reps=4):
    """pairs: list of (gpu, pinned tensor); all copies run concurrently; returns GB/s per pair and aggregate"""
    for g, h in pairs:
        with torch.cuda.stream(streams[g]):
            h.copy_(devs[g], non_blocking=True)
    for g in range(ng):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        for g, h in pairs:
            with torch.cuda.stream(streams[g]):
                h.copy_(devs[g], non_blocking=True)
    for g in range(ng):
        torch.cuda.synchronize(g)
    dt = time.perf_counter() - t0
    return reps * NB * len(pairs) / dt / 1e9


bufs = {}
for node in [None] + nodes:
    for g in range(ng):
        t, r = pinned_on(node)
        bufs[(node, g)] = t
        print(f"alloc node={node} gpu={g} set_mempolicy={r}")
for node in [None] + nodes:
    for g in range(ng):
        if bufs[(node, g)] is not None:
            print(f"D2H alone  gpu {g} <- node {node}: {rate([(g, bufs[(node, g)])]):.1f} GB/s")
for node in [None] + nodes:
    pairs = [(g, bufs[(node, g)]) for g in range(ng) if bufs[(node, g)] is not None]
    if len(pairs) == ng:
        print(f"D2H all {ng} gpus concurrently <- node {node}: {rate(pairs):.1f} GB/s aggregate")
if len(nodes) > 1:
    for shift in range(len(nodes)):
        pairs = [(g, bufs[(nodes[(g * len(nodes) // ng + shift) % len(nodes)], g)]) for g in range(ng)]
        if all(p[1] is not None for p in pairs):
            print(f"D2H all gpus, gpu g <- node (g*{len(nodes)}//{ng}+{shift})%{len(nodes)}: {rate(pairs):.1f} GB/s aggregate")
