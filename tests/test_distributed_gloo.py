"""world_size-2 gloo test of the multi-GPU host logic (sharding + the final gather of saveat blocks).
No GPU here, so each rank's shard is produced by the CPU oracle; the logic under test is b200dde.distributed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(n):
    i = np.arange(n)
    p = np.stack([0.1 + 0.2 * (i % 7) / 6, 0.2 + 0.2 * ((i // 7) % 5) / 4, 0.5 + 1.0 * (i // 35) / max(1, (n - 1) // 35)], axis=1)
    return np.ones((n, 3)), p


def _worker(rank, world, n, port, q):
    sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from b200dde.distributed import gather_saveat, shard_range
    from b200dde.api import saveat_grid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u0, p = _inputs(n)
    lo, hi = shard_range(n, rank, world)
    cfg = O.default_config(model=O.MODEL_BC)
    r = O.solve(cfg, u0[lo:hi], p[lo:hi], [1.0], 0.0, 10.0, saveat=saveat_grid(1.0, (0.0, 10.0)))
    full = gather_saveat(torch.from_numpy(r["u"]), n, dst=0)
    # the library's own gather needs a GPU, but not the way its communicator is set up: rank 0's ncclUniqueId must reach
    # every rank unchanged, and the block sizes handed to b2d_gather_saveat must tile the ensemble in rank order
    from b200dde.distributed import init_library_comm, gather_saveat_library

    class Recorder:
        def comm_init(self, world_, rank_, uid):
            self.args = (world_, rank_, bytes(uid))

        def gather_saveat(self, d_local, counts, root, d_root, stream=0):
            self.counts, self.root, self.has_root_buffer = list(counts), root, d_root is not None
    rec = Recorder()
    init_library_comm(rec)
    ids = [None] * world
    dist.all_gather_object(ids, rec.args[2])
    assert rec.args[:2] == (world, rank) and len(rec.args[2]) == 128 and all(i == ids[0] for i in ids)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "delaydiffeq.jl_b200"))
    from b200dde.distributed import shard_range
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1


def test_two_rank_gather_equals_single_process():
    import oracle_lib as O
    from b200dde.api import saveat_grid
    n = 71  # odd: shards of 36 and 35 exercise the padding
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, n, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    u0, p = _inputs(n)
    ref = O.solve(O.default_config(model=O.MODEL_BC), u0, p, [1.0], 0.0, 10.0, saveat=saveat_grid(1.0, (0.0, 10.0)))
    assert np.array_equal(got, ref["u"])  # sharded result == single-process result, bitwise
