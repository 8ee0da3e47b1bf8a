"""Multi-GPU sharding of the anchoring pass: one process per GPU, read pairs dealt to ranks,
no data-path collective; the only exchange is the gather of the small hit-record lists
(the reference has no collective at all -- SURVEY.md 2, 8e).  Works with backend "nccl"
(device tensors, NVLink) and "gloo" (CPU tensors, tests)."""
import numpy as np

from ._lib import HIT_DTYPE


def shard_range(n_pairs, rank, world):
    """Contiguous, tile-aligned share of [0, n_pairs) for `rank`."""
    tiles = (n_pairs + 31) // 32
    lo = (tiles * rank // world) * 32
    hi = (tiles * (rank + 1) // world) * 32
    return min(lo, n_pairs), min(hi, n_pairs)


def gather_hits_tensor(hits, counts, cap, group=None):
    """One all-gather of (count, first `cap` records) per rank.  hits: [>=cap, 4] int32 tensor on
    this rank's device, counts: int32 tensor with the hit count at index 1 (AF_CNT_HITS).
    Returns (all_counts [world] int32, all_hits [world, cap, 4] int32) tensors, asynchronously on
    the current stream for NCCL."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    payload = torch.empty((cap + 1, 4), dtype=torch.int32, device=hits.device)
    payload[0].zero_()
    payload[0, 0] = counts[1]
    payload[1:] = hits[:cap]
    out = torch.empty((world, cap + 1, 4), dtype=torch.int32, device=hits.device)
    dist.all_gather_into_tensor(out.view(-1, 4), payload, group=group)
    return out[:, 0, 0], out[:, 1:]


def merge_gathered(all_counts, all_hits, pair_offsets):
    """Host side: concatenate the ranks' records with read_ids made global
    (read_id += 2 * first pair of the rank's shard); raises if a rank overflowed the gather cap."""
    counts = all_counts.cpu().numpy()
    recs = all_hits.cpu().numpy()
    cap = recs.shape[1]
    parts = []
    for r, c in enumerate(counts):
        if c > cap:
            raise RuntimeError("rank %d holds %d hits, gather cap is %d" % (r, c, cap))
        h = np.ascontiguousarray(recs[r, :c]).view(np.uint8).reshape(-1).view(HIT_DTYPE).copy()
        h["read_id"] += np.uint32(2 * pair_offsets[r])
        parts.append(h)
    return np.concatenate(parts) if parts else np.zeros(0, HIT_DTYPE)
