"""Full-size runs (BASELINE.json configs 2-4 shapes): every record of the whole run equals the oracle's
(chunk by chunk on all host cores), and the size-independent properties hold -- the oracle re-derives
every reported record from the hit reads alone, a random sample of unreported reads stays unreported,
results do not depend on how the job is cut into batches, planted anchor fragments are recovered."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import hits_equal

pytestmark = pytest.mark.gpu


def _oracle_on_reads(af, oracle, spec, anchor_codes, read_ids):
    """Oracle records for the given global read ids (the generator is a pure function of the pair index)."""
    pairs = np.unique(read_ids >> 1)
    L = spec.read_len
    reads = np.empty((2 * len(pairs), L), dtype=np.uint8)
    # generate only the needed pairs, in runs of consecutive indices
    runs = np.split(pairs, np.nonzero(np.diff(pairs) != 1)[0] + 1) if len(pairs) else []
    k = 0
    for run in runs:
        m1, m2 = af.synth_pairs_host(spec, int(run[0]), len(run))
        reads[2 * k: 2 * (k + len(run)): 2], reads[2 * k + 1: 2 * (k + len(run)): 2] = m1, m2
        k += len(run)
    h = oracle.anchor_reads(anchor_codes, reads, threads=8)
    h["read_id"] = (2 * pairs[h["read_id"] >> 1] + (h["read_id"] & 1)).astype(np.uint32)
    return h, set((2 * pairs[:, None] + np.arange(2)[None, :]).reshape(-1).tolist())


def _oracle_on_range(af, oracle, spec, anchor_codes, lo, hi):
    """Oracle records (read ids relative to `lo`) of pairs [lo, hi), generated and anchored on all host cores."""
    n, threads = hi - lo, os.cpu_count() or 1
    reads = np.empty((2 * n, spec.read_len), dtype=np.uint8)
    step = (n + threads - 1) // threads

    def fill(a):
        b = min(a + step, n)
        m1, m2 = af.synth_pairs_host(spec, lo + a, b - a)
        reads[2 * a: 2 * b: 2], reads[2 * a + 1: 2 * b: 2] = m1, m2

    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(fill, range(0, n, step)))
    return oracle.anchor_reads(anchor_codes, reads, threads=threads)


@pytest.mark.parametrize("name,n,anchor_len,sub_ppm,fusion_ppm", [
    ("config2_10M_pairs", 10_000_000, 6783, 10_000, 0),
    ("config4_long_anchor_fusions", 5_000_000, 10_000, 15_000, 10_000),
])
def test_full_size_properties(name, n, anchor_len, sub_ppm, fusion_ppm):
    import torch
    import anchored_fusion_b200 as af
    from oracle import oracle
    spec = af.synth_spec(seed=21, ref_len=10_000_000, anchor_start=3_000_000, anchor_len=anchor_len, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=sub_ppm, fusion_ppm=fusion_ppm)
    anchor = af.synth_anchor(spec)
    acodes = oracle.encode(anchor)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    batch = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    hits, stats = eng.anchor(batch, cand_cap=n // 2, hits_cap=n // 8)
    assert stats["flagged"] < n // 2
    # (1) sorted, unique read ids; fields in range
    assert np.all(np.diff(hits["read_id"].astype(np.int64)) > 0)
    assert np.all(hits["clip_l"].astype(int) + hits["m_len"] + hits["clip_r"] == 150)
    assert np.all((hits["score_strand"] >> 1) >= 30) and np.all(hits["pos"] >= 1) and np.all(hits["pos"] + hits["m_len"].astype(int) - 1 <= anchor_len)
    # (1b) the whole run, record for record, against the oracle (2.5 M pairs at a time)
    for lo in range(0, n, 2_500_000):
        hi = min(lo + 2_500_000, n)
        want_part = _oracle_on_range(af, oracle, spec, acodes, lo, hi)
        want_part["read_id"] += np.uint32(2 * lo)
        part = hits[(hits["read_id"] >= 2 * lo) & (hits["read_id"] < 2 * hi)]
        assert hits_equal(part, want_part), (name, lo)
    # (2) the oracle, given only the reported pairs, reproduces every record bit for bit (both
    #     mates of those pairs: an unreported mate of a reported pair must stay unreported)
    want, _ = _oracle_on_reads(af, oracle, spec, acodes, hits["read_id"])
    assert hits_equal(hits, want)
    # (3) a random sample of pairs: reported iff the oracle reports them
    rng = np.random.default_rng(3)
    sample = np.sort(rng.choice(n, 200_000, replace=False))
    runs_want, ids = _oracle_on_reads(af, oracle, spec, acodes, (2 * sample).astype(np.uint32))
    got = hits[np.isin(hits["read_id"] >> 1, sample)]
    assert hits_equal(got, runs_want)
    # (4) planted truth: pairs whose fragment lies inside the anchor are anchored on both mates
    #     (at <= 1.5 % substitutions a 150 bp read keeps a 19-mer and scores >= 30)
    expected_inside = n * (anchor_len - 300) / (10_000_000 - 300)
    both = np.sum(np.diff(hits["read_id"].astype(np.int64)) == 1) 
    assert both > 0.85 * expected_inside
    if fusion_ppm:
        clipped = np.sum((hits["clip_l"] >= 15) | (hits["clip_r"] >= 15))
        assert clipped > 0.3 * n * fusion_ppm / 1e6          # junction reads come back soft-clipped
    # (5) batch-cut invariance: the same pairs as 7 uneven device batches give the same records
    cuts = [0, 32 * 1000, 32 * 50_001, n // 3 // 32 * 32, n // 2 // 32 * 32, (n - 77) // 32 * 32, n - 32, n]
    parts = []
    for lo, hi in zip(cuts, cuts[1:]):
        b = af.synth_pairs_device(spec, lo, hi - lo, index.pad_byte, 0)
        h, _ = eng.anchor(b, cand_cap=max(hi - lo, 1024), hits_cap=max(hi - lo, 1024))
        h["read_id"] += np.uint32(2 * lo)
        parts.append(h)
    assert hits_equal(np.concatenate(parts), hits)
    # (6) and through the host-buffer pipeline (pinned tiles, chunks of 1M pairs)
    sub = 3_000_000
    host = af.PackedBatch(batch.packed[: af.layout(150, sub).packed_bytes // 4].cpu().numpy().view(np.uint32), sub, 150, 150)
    h2, _ = eng.anchor_host(host)
    assert hits_equal(h2, hits[hits["read_id"] < 2 * sub])
    torch.cuda.synchronize()
