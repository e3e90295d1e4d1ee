"""Multi-GPU plumbing: the problem batch is sharded contiguously over ranks (independent problems, no
data-path collective); the only exchange is the final gather of controllers / costs / status.
Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_problems, world, rank):
    """Contiguous, balanced split: the first (n % world) ranks get one extra problem."""
    base, extra = divmod(n_problems, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_problem_axis(local, n_problems, group=None, dst=None):
    """Gather of a per-problem tensor [B_local, ...] -> [B, ...].  dst=None: all_gather, every rank gets the result (uneven
    shards are padded to the largest shard for the collective and trimmed afterwards).  dst=r: gather onto rank r only,
    every shard received straight into its slice of the final tensor (the other ranks return None): each rank sends its
    shard once instead of receiving everybody's -- 1/world of the all_gather's traffic."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(n_problems, world, r) for r in range(world)]
    max_b = max(e - s for s, e in sizes)
    if dst is not None:
        even = all(e - s == max_b for s, e in sizes)
        if not even:  # the collective needs equal shapes: pad to the largest shard, trim on the receiver
            pad = torch.zeros((max_b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            pad[: local.shape[0]] = local
            local = pad
        local = local.contiguous()
        if dist.get_rank(group) != dst:
            dist.gather(local, None, dst=dst, group=group)
            return None
        if even:
            out = torch.empty((n_problems,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            dist.gather(local, [out[s:e] for s, e in sizes], dst=dst, group=group)
            return out
        outs = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, outs, dst=dst, group=group)
        return torch.cat([o[: e - s] for o, (s, e) in zip(outs, sizes)], dim=0)
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


def gather_controllers(K, k, sigK, n_problems, extra=(), group=None, dst=None):
    """Final gather of K[B_local,T,du,dx], k, sigK (+ any per-problem extras such as costs, alpha, status); dst as in
    gather_problem_axis."""
    return tuple(gather_problem_axis(t, n_problems, group, dst) for t in (K, k, sigK) + tuple(extra))
