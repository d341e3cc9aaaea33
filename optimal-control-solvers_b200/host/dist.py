"""Multi-GPU plumbing: the trajectory batch shards trivially (independent trajectories,
SURVEY.md section 8e) -- one process per GPU, contiguous slices, no data-path collective.
The only exchange is the gather of the small per-trajectory results J and x(T) with
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(Bt_total: int, rank: int, world: int):
    """Contiguous slice [b0, b1) of a batch of Bt_total trajectories owned by `rank`;
    the first Bt_total % world ranks take one extra trajectory."""
    base, rem = divmod(int(Bt_total), int(world))
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def gather_results(J, xT, world: int, group=None):
    """All-gathers J [Bt_local] and xT [n, Bt_local] (torch tensors on the rank's device) into
    J_all [world*Bt_local] and xT_all [world, n, Bt_local].  Equal shard sizes (weak scaling)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return J, xT.unsqueeze(0)
    J_all = torch.empty(world * J.numel(), dtype=J.dtype, device=J.device)
    xT_all = torch.empty((world,) + tuple(xT.shape), dtype=xT.dtype, device=xT.device)
    dist.all_gather_into_tensor(J_all, J.contiguous(), group=group)
    dist.all_gather_into_tensor(xT_all.view(-1), xT.contiguous().view(-1), group=group)
    return J_all, xT_all


def gather_ragged(J, world: int, rank: int, Bt_total: int, group=None):
    """All-gather of J when Bt_total does not divide evenly: pads every shard to the largest
    one, gathers, and strips the padding.  Returns the full [Bt_total] tensor on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return J
    sizes = [shard_range(Bt_total, r, world) for r in range(world)]
    big = max(b1 - b0 for b0, b1 in sizes)
    pad = torch.zeros(big, dtype=J.dtype, device=J.device)
    pad[: J.numel()] = J
    out = torch.empty(world * big, dtype=J.dtype, device=J.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * big: r * big + (b1 - b0)] for r, (b0, b1) in enumerate(sizes)])
