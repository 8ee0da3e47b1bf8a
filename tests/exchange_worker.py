"""torchrun worker of tests/test_gpu_exchange.py: N ranks, one GPU each.  Every rank anchors its own
synthetic batches with the hit exchange on, then checks that its copy of EVERY rank's log equals the
records an NCCL all-gather of the ranks' hit lists delivers, and that rank 0's merged result equals
the CPU oracle on the union of the shards.  Prints "EXCHANGE_OK <rank>" on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import anchored_fusion_b200 as af
    from anchored_fusion_b200 import dist as afdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, n_batches, n_slots = 40_000, 5, 2
    spec = af.synth_spec(seed=11, ref_len=300_000, anchor_start=100_000, anchor_len=3000, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=15_000, fusion_ppm=50_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, local)
    ex = afdist.HitExchange(rank, world, n_slots, 60_000, dev)
    local_hits = []
    for round_ in range(2):                      # second round after a reset: logs start over
        if round_:
            ex.reset()
            local_hits = []
        for b in range(n_batches):
            first = (b * world + rank) * n       # job-wide index of the batch's first pair
            batch = af.synth_pairs_device(spec, first, n, index.pad_byte, local)
            torch.cuda.synchronize()             # generated on the current stream; slot 1 runs on its own stream
            sl = b % n_slots
            hits, counts = eng.enqueue(batch, slot=sl, exchange=ex, pair_base=first)
            st = eng.slot_stream(sl) if sl else torch.cuda.current_stream(dev)
            st.synchronize()                     # the slot's hits tensor is reused by batch b + n_slots
            c = counts.cpu().numpy().view(np.uint32)
            assert c[2] == 0, "status %d" % c[2]
            local_hits.append((first, hits[: int(c[1])].cpu().numpy().view(np.uint8).reshape(-1).view(af.HIT_DTYPE).copy()))
        got = ex.collect()
        assert len(got) == world * n_batches, (len(got), world, n_batches)
        # reference delivery: all-gather each batch's list with NCCL
        for b in range(n_batches):
            first, h = local_hits[b]
            cap = 8192
            assert len(h) <= cap
            t = torch.zeros((2 + cap, 4), dtype=torch.int32, device=dev)
            t[0, 1] = len(h)
            if len(h):
                t[2: 2 + len(h)] = torch.from_numpy(h.view(np.int32).reshape(-1, 4)).to(dev)
            ac, ah = afdist.gather_hits_tensor(t, cap)
            ac, ah = ac.cpu().numpy(), ah.cpu().numpy()
            for r in range(world):
                base = (b * world + r) * n
                mine = [hh for (src, pb, hh) in got if src == r and pb == base]
                assert len(mine) == 1, "rank %d: %d log batches of rank %d at base %d" % (rank, len(mine), r, base)
                ref = np.ascontiguousarray(ah[r, : ac[r]]).view(np.uint8).reshape(-1).view(af.HIT_DTYPE)
                assert mine[0].tobytes() == ref.tobytes(), "rank %d: batch %d of rank %d differs" % (rank, b, r)
                assert len(ref) > 0
    if rank == 0:
        # the merged job result equals the oracle over all pairs of all ranks
        from oracle import oracle
        merged = afdist.globalise([(pb, h) for (_, pb, h) in got])
        total = n * n_batches * world
        m1, m2 = af.synth_pairs_host(spec, 0, total)
        reads = np.empty((2 * total, spec.read_len), dtype=np.uint8)
        reads[0::2], reads[1::2] = m1, m2
        want = oracle.anchor_reads(oracle.encode(af.synth_anchor(spec)), reads, threads=8)
        assert len(merged) == len(want), (len(merged), len(want))
        for name in af.HIT_DTYPE.names:
            assert (merged[name].astype(np.int64) == want[name].astype(np.int64)).all(), name
    dist.barrier()
    ex.close()
    print("EXCHANGE_OK %d" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
