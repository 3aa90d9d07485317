"""Crop-sharded data parallelism for the hot path: every crop is independent, so
rank r of R takes the contiguous slice [r*n/R, (r+1)*n/R) -- contiguous so that rank 0 can paste
results back in vehicle order (trajectory_inference.py:150-152) -- and the only collective is the
all-gather of the completed uint8 crops after the path.  One process per GPU (torch.distributed,
NCCL on the device; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous [begin, end) of `n` crops owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_crops(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of completed crops: `local` is this rank's (n_local,H,W,3) uint8 slice (shard_range
    order); returns the (n_total,H,W,3) tensor in global crop order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local slice does not match shard_range"
    if n_total % world == 0:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged: pad to the largest shard, gather, strip
    nmax = max(e - b for b, e in sizes)
    padded = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], 0)
