"""Multi-GPU host logic: one process per GPU, trajectories sharded in contiguous ranges.

There is no collective inside the solve: trajectories are independent.  The only collective of the path is the
optional final gather of the saveat block to one rank; it works on whatever backend the process group uses
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_traj, rank, world):
    """Contiguous slice [lo, hi) of trajectories for `rank` (keeps neighbouring sweep points in one warp tile)."""
    base, rem = divmod(int(n_traj), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_saveat(local_out, n_traj, dst=0, group=None):
    """Gather per-rank saveat blocks [n_local, n_out, n_state] to `dst`; returns [n_traj, n_out, n_state] on dst,
    None elsewhere.  Shards may differ in size by one trajectory, so the blocks are padded to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_traj, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = local_out
    if local_out.shape[0] < nmax:
        pad = torch.empty((nmax,) + tuple(local_out.shape[1:]), dtype=local_out.dtype, device=local_out.device)
        pad[: local_out.shape[0]] = local_out
    # the blocks land directly in their final place: one [world * nmax, ...] allocation whose chunks are the receive buffers
    full = torch.empty((world * nmax,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device) if rank == dst else None
    bufs = list(full.chunk(world, dim=0)) if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    if all(hi - lo == nmax for lo, hi in sizes):
        return full
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def init_library_comm(solver, group=None):
    """Give `solver`'s handle the NCCL communicator of the library's own gather (b2d_comm_init): rank 0 creates the unique
    id, the process group (any backend) carries its 128 bytes to the other ranks."""
    from .api import comm_unique_id
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    solver.comm_init(world, rank, box[0])


def gather_saveat_library(solver, local_out, n_traj, dst=0, group=None, out=None):
    """The library's gather (b2d_gather_saveat: grouped ncclSend / ncclRecv over NVLink) of per-rank saveat blocks
    [n_local, n_out, n_state] to rank `dst`; returns the [n_traj, n_out, n_state] block there (in `out` when given), None
    elsewhere.  Asynchronous on torch's current stream."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = int(local_out[0].numel()) if local_out.shape[0] else 0
    counts = [(hi - lo) * per for lo, hi in (shard_range(n_traj, r, world) for r in range(world))]
    assert counts[rank] == local_out.numel() and local_out.is_contiguous()
    full = None
    if rank == dst:
        full = out if out is not None else torch.empty((n_traj,) + tuple(local_out.shape[1:]), dtype=local_out.dtype, device=local_out.device)
    stream = torch.cuda.current_stream(local_out.device).cuda_stream
    solver.gather_saveat(local_out.data_ptr(), counts, dst, full.data_ptr() if full is not None else None, stream)
    return full
