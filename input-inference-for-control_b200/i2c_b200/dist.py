"""Multi-GPU plumbing: the problem batch is sharded contiguously over ranks (independent problems, no
data-path collective); the only exchange is the final gather of controllers / costs / status.
Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_problems, world, rank):
    """Contiguous, balanced split: the first (n % world) ranks get one extra problem."""
    base, extra = divmod(n_problems, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_problem_axis(local, n_problems, group=None):
    """all_gather of a per-problem tensor [B_local, ...] -> [B, ...] on every rank (uneven shards are padded to
    the largest shard for the collective and trimmed afterwards)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(n_problems, world, r) for r in range(world)]
    max_b = max(e - s for s, e in sizes)
    if all(e - s == max_b for s, e in sizes) and hasattr(dist, "all_gather_into_tensor"):
        # even shards (the usual case): ONE collective straight into the final [B, ...] tensor -- no padding, no
        # per-rank receive buffers, no concatenation copy
        out = torch.empty((n_problems,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((max_b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: e - s] for o, (s, e) in zip(outs, sizes)], dim=0)


def gather_controllers(K, k, sigK, n_problems, extra=(), group=None):
    """Final gather of K[B_local,T,du,dx], k, sigK (+ any per-problem extras such as costs, alpha, status)."""
    return tuple(gather_problem_axis(t, n_problems, group) for t in (K, k, sigK) + tuple(extra))
