"""Multi-GPU plumbing: the path shards by patch.

Every height map carries its own 1-texel border (main.cpp:135-141) and depends only on its
104-byte Quad, so ranks take contiguous leaf ranges and never exchange anything while computing.
The only collective is the optional gather of finished patches into one buffer; it runs on
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous, balanced [lo, hi) of `n_units` patches for `rank`; the shards tile [0, n) in
    rank order so the gathered buffer is in the reference's emission order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_patches(local, n_units=None):
    """All ranks contribute their shard (first dimension = patches, possibly ragged across ranks)
    and receive the concatenation in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = [int(c.item()) for c in all_counts]
    if n_units is not None and sum(all_counts) != n_units:
        raise RuntimeError(f"shards hold {sum(all_counts)} patches, expected {n_units}")
    if len(set(all_counts)) == 1:                              # equal shards: one all-gather in place
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged shards (they differ by at most one patch): pad to the largest, gather, trim
    cmax = max(all_counts)
    padded = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(all_counts)])
