#!/usr/bin/env python
"""Diagnostic: GPU path vs oracle on a synthetic spec; prints the records only one side holds."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--anchor-len", type=int, default=3000)
    ap.add_argument("--ref-len", type=int, default=300_000)
    ap.add_argument("--pairs", type=int, default=100_000)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--sub-ppm", type=int, default=15_000)
    ap.add_argument("--fusion-ppm", type=int, default=50_000)
    ap.add_argument("--modes", type=str, default="9,10")
    args = ap.parse_args()
    import anchored_fusion_b200 as af
    from anchored_fusion_b200._lib import check, lib
    from oracle import oracle
    spec = af.synth_spec(seed=args.seed, ref_len=args.ref_len, anchor_start=100_000, anchor_len=args.anchor_len, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=args.sub_ppm, fusion_ppm=args.fusion_ppm)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    n = args.pairs
    m1, m2 = af.synth_pairs_host(spec, args.first, n)
    codes = np.empty((2 * n, 150), np.uint8)
    codes[0::2], codes[1::2] = m1, m2
    want = oracle.anchor_reads(oracle.encode(anchor), codes, threads=16)
    for mode in [int(x) for x in args.modes.split(",")]:
        check(lib().af_seed_scan_config(0, mode))
        got, stats = eng.anchor(af.synth_pairs_device(spec, args.first, n, index.pad_byte, 0))
        w = {int(r["read_id"]): r for r in want}
        g = {int(r["read_id"]): r for r in got}
        only_w = sorted(set(w) - set(g))
        only_g = sorted(set(g) - set(w))
        diff = [k for k in set(w) & set(g) if w[k].tobytes() != g[k].tobytes()]
        print("mode", mode, "oracle", len(want), "gpu", len(got), stats, "oracle-only", len(only_w), "gpu-only", len(only_g), "differ", len(diff))
        for k in only_w[:12]:
            print("  oracle-only", w[k], "seq", af.codes_to_ascii(codes[k]).decode()[:60])
        for k in only_g[:5]:
            print("  gpu-only", g[k])
        for k in diff[:5]:
            print("  differ", w[k], g[k])


if __name__ == "__main__":
    main()
