"""Parity at the north star's own sizes, driver-visible: BASELINE.json configs[2] (100 M synthetic 2x150 bp
pairs, single anchor) and configs[3] (50 M pairs, ~10 kb anchor, 1 % fusion fragments, sequencing errors).
Every 16-byte record of the CUDA path, chunk by chunk, equals the CPU oracle's on the same pairs.  The pairs
are a pure function of (seed, pair index): the device generator writes packed tiles in HBM, the oracle side
gets base codes from oracle/af_synth.cpp (same generator definition; equality of the two is a test of its
own in test_gpu_parity.py and test_oracle.py).  A JSON record of each run is left in gpurun_out/."""
import json
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, pairs, chunk, anchor_len, sub_ppm, fusion_ppm):
    import anchored_fusion_b200 as af
    from oracle import oracle
    threads = os.cpu_count() or 1
    spec = af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=anchor_len, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=sub_ppm, fusion_ppm=fusion_ppm)
    anchor = af.synth_anchor(spec)
    assert anchor == oracle.synth_anchor(spec)
    acodes = oracle.encode(anchor)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    reads = np.empty((2 * chunk, 150), dtype=np.uint8)
    t0 = time.time()
    done = hits_total = flagged_total = 0
    t_gen = t_cpu = t_gpu = 0.0
    while done < pairs:
        n = min(chunk, pairs - done)
        t = time.time()
        oracle.synth_reads(spec, done, n, threads=threads, out=reads)
        t_gen += time.time() - t
        t = time.time()
        want = oracle.anchor_reads(acodes, reads[: 2 * n], threads=threads)
        t_cpu += time.time() - t
        t = time.time()
        got, stats = eng.anchor(af.synth_pairs_device(spec, done, n, index.pad_byte, 0), cand_cap=n, hits_cap=n // 4)
        t_gpu += time.time() - t
        assert len(got) == len(want), (name, done, len(got), len(want))
        assert got.tobytes() == want.tobytes(), (name, "chunk starting at pair %d differs" % done)
        done += n
        hits_total += len(got)
        flagged_total += stats["flagged"]
    res = {"parity": "bit-exact", "config": name, "pairs": done, "read_len": 150, "anchor_len": anchor_len, "sub_ppm": sub_ppm,
           "fusion_ppm": fusion_ppm, "chunk_pairs": chunk, "anchored_reads": hits_total, "flagged_reads": flagged_total,
           "seconds": {"total": time.time() - t0, "host_generation": t_gen, "cpu_oracle": t_cpu,
                       "gpu_incl_device_generation_and_d2h": t_gpu},
           "host_threads": threads,
           "compared": "16-byte records (read_id, pos, clip_l, m_len, clip_r, score*2+strand), byte for byte, per chunk"}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_%s.json" % name), "w") as fh:
        json.dump(res, fh, indent=1)
    return res


def test_config3_100m_pairs_every_record_equals_the_oracle():
    res = _run("config3_100m", 100_000_000, 10_000_000, 6783, 10_000, 0)
    assert res["anchored_reads"] > 100_000


def test_config4_50m_pairs_long_anchor_fusions_errors_every_record_equals_the_oracle():
    res = _run("config4_50m", 50_000_000, 10_000_000, 10_000, 15_000, 10_000)
    assert res["anchored_reads"] > 500_000
