"""Multi-GPU host logic on CPU: world_size-2 gloo run of the shard assignment and the hit-list
gather (the only collective of the path)."""
import os
import socket

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist
    from anchored_fusion_b200 import dist as afdist
    from anchored_fusion_b200._lib import HIT_DTYPE
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = afdist.shard_range(n_pairs, rank, world)
    # fake per-rank hit list: every 7th read of the shard, shard-local read ids
    ids = np.arange(0, 2 * (hi - lo), 7, dtype=np.uint32)
    hits = np.zeros(len(ids), dtype=HIT_DTYPE)
    hits["read_id"], hits["pos"], hits["m_len"] = ids, (2 * lo + ids) % 1000 + 1, 50
    cap = 4096
    buf = torch.zeros((2 + cap, 4), dtype=torch.int32)        # engine layout: 2 counter rows, then records
    buf[2: 2 + len(hits)] = torch.from_numpy(hits.view(np.int32).reshape(-1, 4))
    buf[0, 1] = len(hits)
    all_counts, all_hits = afdist.gather_hits_tensor(buf, cap)
    offsets = [afdist.shard_range(n_pairs, r, world)[0] for r in range(world)]
    merged = afdist.merge_gathered(all_counts, all_hits, offsets)
    q.put((rank, lo, hi, merged.tobytes()))
    dist.destroy_process_group()


def test_shard_range_partitions_tiles():
    from anchored_fusion_b200.dist import shard_range
    for n, world in [(10_000_000, 8), (11258, 2), (33, 4), (0, 2), (31, 8)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b and b % 32 == 0 or b == n


def test_two_rank_gather_over_gloo():
    import torch.multiprocessing as mp
    from anchored_fusion_b200._lib import HIT_DTYPE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_pairs, world = _free_port(), 10_000, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][3] == out[1][3]                      # every rank holds the same merged list
    merged = np.frombuffer(out[0][3], dtype=HIT_DTYPE)
    want = np.concatenate([2 * lo + np.arange(0, 2 * (hi - lo), 7) for _, lo, hi, _ in out])
    assert np.array_equal(merged["read_id"], want.astype(np.uint32))
    assert np.array_equal(merged["pos"], (want % 1000 + 1).astype(np.int32))
