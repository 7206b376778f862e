"""Process-group plumbing for the batch-sharded forward (one process per GPU, torch.distributed).

The forward itself needs no collective: every rank runs the same weights on
its own shard of the batch. The only communication is the start/stop barrier and the max-over-ranks
reduction of the measured time, plus an optional gather of per-rank outputs for evaluation.
Backend `nccl` on GPUs, `gloo` for CPU tests.
"""
import os

import torch
import torch.distributed as dist


def setup(backend=None):
    """Initialise from torchrun's environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_*). Returns
    (rank, world, local_rank); a no-op for single-process runs."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def barrier(world):
    if world > 1:
        dist.barrier()


def max_over_ranks(x, world, device="cpu"):
    """max of a python float over all ranks (device time is always reported as the slowest rank's)."""
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch(global_batch, rank, world):
    """[start, stop) of this rank's slice of a global batch (contiguous, remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_outputs(out, world, dst=0):
    """Concatenate per-rank output dicts {task: [b_rank, ...]} along the batch on rank `dst` (evaluation).
    Ranks may hold DIFFERENT batch sizes (shard_batch gives the remainder to the low ranks): the sizes are
    all-gathered first, every rank pads its shard to the largest one for the collective, and `dst` trims."""
    if world == 1:
        return out
    rank = dist.get_rank()
    first = out[sorted(out)[0]]
    n = torch.tensor([first.shape[0]], dtype=torch.int64, device=first.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    res = {}
    for k in sorted(out):
        t = out[k].contiguous()
        if t.shape[0] != sizes[rank]:
            raise ValueError(f"gather_outputs: {k!r} has batch {t.shape[0]}, expected {sizes[rank]}")
        if t.shape[0] < mx:
            t = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
        parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, parts, dst=dst)
        if parts is not None:
            res[k] = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return res


def teardown(world):
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()
