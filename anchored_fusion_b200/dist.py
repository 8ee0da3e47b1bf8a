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


def gather_hits_tensor(counts_and_hits, cap, group=None):
    """One all-gather of (counters, first `cap` records) per rank.  counts_and_hits: the engine's
    [2 + capacity, 4] int32 tensor (Anchorer.counts_and_hits): rows 0..1 hold the 8 counters (hit count
    at flat index 1, AF_CNT_HITS), rows 2.. the records -- contiguous, so nothing is packed first.
    Returns (all_counts [world] int32, all_hits [world, cap, 4] int32) tensors, asynchronously on
    the current stream for NCCL."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = counts_and_hits[: 2 + cap]
    out = torch.empty((world, 2 + cap, 4), dtype=torch.int32, device=src.device)
    dist.all_gather_into_tensor(out.view(-1, 4), src, group=group)
    return out[:, 0, 1], out[:, 2:]


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
