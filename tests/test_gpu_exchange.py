"""The multi-GPU hit exchange: the hit-compaction kernel appends each batch's records
to a log on every rank over NVLink peer memory.  World 1 in-process (the log on the own GPU), and --
when the box has two GPUs -- two ranks under torchrun against an NCCL all-gather and the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, hits_equal

pytestmark = pytest.mark.gpu


def _setup(n_slots=2, log_cap=50_000):
    import torch
    import anchored_fusion_b200 as af
    from anchored_fusion_b200 import dist as afdist
    spec = af.synth_spec(seed=21, ref_len=300_000, anchor_start=100_000, anchor_len=3000, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=15_000, fusion_ppm=50_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, 0)
    ex = afdist.HitExchange(0, 1, n_slots, log_cap, torch.device("cuda", 0))
    return torch, af, afdist, spec, index, eng, ex


def test_log_on_one_gpu_holds_every_batch_in_order():
    torch, af, afdist, spec, index, eng, ex = _setup()
    n, want = 30_000, []
    sizes = [n, n, 0, n, 31, n]                         # an empty batch leaves no marker
    first = 0
    for b, nb in enumerate(sizes):
        batch = af.synth_pairs_device(spec, first, nb, index.pad_byte, 0)
        ref, _ = eng.anchor(batch)                      # the plain path (slot 0), synchronous
        if nb:
            want.append((b % 2, first, ref))
        eng.enqueue(batch, slot=b % 2, exchange=ex, pair_base=first)
        torch.cuda.synchronize()
        first += nb
    got = ex.collect()                                  # slot 0's batches, then slot 1's
    want.sort(key=lambda t: t[0])
    assert [(pb, len(h)) for (_, pb, h) in got] == [(pb, len(h)) for (_, pb, h) in want]
    for (_, _, h), (_, _, ref) in zip(got, want):
        assert hits_equal(h, ref)
    assert sum(len(h) for _, _, h in got) > 500
    # job-wide read_ids
    merged = afdist.globalise([(pb, h) for (_, pb, h) in got])
    every, _ = eng.anchor(af.synth_pairs_device(spec, 0, first, index.pad_byte, 0))
    assert np.array_equal(merged["read_id"], every["read_id"].astype(np.int64))
    assert np.array_equal(merged["pos"], every["pos"]) and np.array_equal(merged["score_strand"], every["score_strand"])
    # reset empties the logs
    ex.reset()
    assert ex.collect() == []
    recs, status, nb = ex.read(0, 0)
    assert len(recs) == 0 and status == 0 and nb == 0
    ex.close()


def test_log_overflow_is_flagged_not_silent():
    torch, af, afdist, spec, index, eng, ex = _setup(n_slots=1, log_cap=64)
    batch = af.synth_pairs_device(spec, 0, 30_000, index.pad_byte, 0)
    hits, counts = eng.enqueue(batch, exchange=ex, pair_base=0)
    torch.cuda.synchronize()
    c = counts.cpu().numpy().view(np.uint32)
    assert c[1] > 64 and c[2] & afdist.STATUS_LOG_OVERFLOW
    recs, status, nb = ex.read(0, 0)
    assert status & afdist.STATUS_LOG_OVERFLOW and len(recs) == 64 and nb == 1
    with pytest.raises(RuntimeError):
        ex.collect()
    # the batch's own hit list is unaffected
    ref, _ = eng.anchor(batch)
    assert hits_equal(hits[: c[1]].cpu().numpy().view(np.uint8).reshape(-1).view(af.HIT_DTYPE), ref)
    ex.close()


def test_bad_arguments_fail_loudly():
    torch, af, afdist, spec, index, eng, ex = _setup(n_slots=1, log_cap=64)
    batch = af.synth_pairs_device(spec, 0, 64, index.pad_byte, 0)
    with pytest.raises(af.AnchoredFusionError):
        eng.enqueue(batch, slot=0, exchange=ex, pair_base=-1)
    with pytest.raises(af.AnchoredFusionError):
        afdist.HitExchange(0, 17, 1, 64, torch.device("cuda", 0))
    ex.close()


def test_two_ranks_over_nvlink_equal_nccl_all_gather_and_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "exchange_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "EXCHANGE_OK 0" in r.stdout and "EXCHANGE_OK 1" in r.stdout
