"""Bit-exact parity at the north star's size: N synthetic 2x150 bp pairs (default 100 M, BASELINE.json
configs[2]) through the CUDA path, chunk by chunk, against the CPU oracle on the same pairs.
The pairs are a pure function of (seed, pair index): the device generator writes packed tiles in HBM, the
host generator hands the oracle base codes; tests/test_gpu_parity.py proves the two generators equal.

  python tools/parity_fullsize.py --pairs 100000000 --chunk 5000000 --out profiles/r01_parity_100m.json
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import anchored_fusion_b200 as af  # noqa: E402
from oracle import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=100_000_000)
ap.add_argument("--chunk", type=int, default=5_000_000)
ap.add_argument("--anchor-len", type=int, default=6783)
ap.add_argument("--sub-ppm", type=int, default=10_000)
ap.add_argument("--fusion-ppm", type=int, default=0)
ap.add_argument("--out", type=str, default="")
args = ap.parse_args()
threads = os.cpu_count() or 1
spec = af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=args.anchor_len, read_len=150,
                     frag_mean=300, frag_sd=30, sub_ppm=args.sub_ppm, fusion_ppm=args.fusion_ppm)
anchor = af.synth_anchor(spec)
acodes = oracle.encode(anchor)
index = af.AnchorIndex(anchor)
eng = af.Anchorer(index, 0)
reads = np.empty((2 * args.chunk, 150), dtype=np.uint8)


def fill(job):
    first, lo, hi = job
    m1, m2 = af.synth_pairs_host(spec, first + lo, hi - lo)      # ctypes call: releases the GIL
    reads[2 * lo: 2 * hi: 2], reads[2 * lo + 1: 2 * hi: 2] = m1, m2


t0 = time.time()
done = hits_total = flagged_total = 0
t_gen = t_cpu = t_gpu = 0.0
with ThreadPoolExecutor(max_workers=threads) as pool:
    while done < args.pairs:
        n = min(args.chunk, args.pairs - done)
        t = time.time()
        step = (n + threads - 1) // threads
        list(pool.map(fill, [(done, lo, min(lo + step, n)) for lo in range(0, n, step)]))
        t_gen += time.time() - t
        t = time.time()
        want = oracle.anchor_reads(acodes, reads[: 2 * n], threads=threads)
        t_cpu += time.time() - t
        t = time.time()
        got, stats = eng.anchor(af.synth_pairs_device(spec, done, n, index.pad_byte, 0))
        t_gpu += time.time() - t
        if got.tobytes() != want.tobytes():
            bad = next(i for i in range(min(len(got), len(want))) if got[i].tobytes() != want[i].tobytes()) if len(got) == len(want) else -1
            print(json.dumps({"parity": "FAILED", "chunk_first_pair": done, "gpu_hits": len(got), "oracle_hits": len(want), "first_diff": bad}))
            sys.exit(1)
        done += n
        hits_total += len(got)
        flagged_total += stats["flagged"]
        print("[parity] %d / %d pairs, %d anchored reads so far, all records equal" % (done, args.pairs, hits_total), file=sys.stderr, flush=True)
res = {"parity": "bit-exact", "pairs": done, "read_len": 150, "anchor_len": args.anchor_len, "sub_ppm": args.sub_ppm,
       "fusion_ppm": args.fusion_ppm, "chunk_pairs": args.chunk, "anchored_reads": hits_total, "flagged_reads": flagged_total,
       "seconds": {"total": time.time() - t0, "host_generation": t_gen, "cpu_oracle": t_cpu, "gpu_incl_device_generation_and_d2h": t_gpu},
       "host_threads": threads, "compared": "16-byte records (read_id, pos, clip_l, m_len, clip_r, score*2+strand), byte for byte, per chunk"}
print(json.dumps(res))
if args.out:
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
