"""Frame sharding across GPUs: frames and stereo pairs are independent units, so the
only multi-GPU mode is data parallel -- contiguous shards, no data-path collective, host gather of
counts. The reference itself is single-GPU (initDevice, /root/reference/cuda_utils.h:41-67)."""


def shard_range(n_units, world_size, rank):
    """Contiguous shard [lo, hi) of n_units for `rank`: unit u belongs to rank floor(u*world/n)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    lo = (n_units * rank + world_size - 1) // world_size
    hi = (n_units * (rank + 1) + world_size - 1) // world_size
    return lo, hi


def gather_counts(local_counts, n_units, world_size, rank, dist=None):
    """All ranks' per-frame keypoint counts in frame order (host gather over torch.distributed)."""
    import torch
    if dist is None or world_size == 1:
        return list(local_counts)
    out = [None] * world_size
    dist.all_gather_object(out, list(local_counts))
    flat = []
    for r in range(world_size):
        lo, hi = shard_range(n_units, world_size, r)
        assert len(out[r]) == hi - lo
        flat.extend(out[r])
    return flat
