"""Time the seed-scan kernel alone (CUDA events, warm, input > L2) over its tuning knobs."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import anchored_fusion_b200 as af  # noqa: E402
from anchored_fusion_b200._lib import check, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--configs", default="12:3:768,12:3:640,12:3:512,13:3:768")
args = ap.parse_args()
peak = 6554.2
spec = af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=6783, read_len=args.read_len,
                     frag_mean=2 * args.read_len, sub_ppm=10_000)
anchor = af.synth_anchor(spec)
alg = 2 * ((2 * args.read_len + 7) // 8)
batch = None
for cfg in args.configs.split(","):
    kp, mode, threads = (int(x) for x in cfg.split(":"))
    index = af.AnchorIndex(anchor, kp=kp)
    eng = af.Anchorer(index, 0)
    if batch is None:
        batch = af.synth_pairs_device(spec, 0, args.pairs, index.pad_byte, 0)
    check(lib().af_seed_scan_config(threads, mode))
    for _ in range(3):
        flags = eng.seed_scan(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        flags = eng.seed_scan(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    f = flags.cpu().numpy().view("uint32")
    nflag = int(sum(bin(int(x)).count("1") for x in f.reshape(-1)[:20000]))   # sample
    gbs = alg * args.pairs / (ms * 1e-3) / 1e9
    print(json.dumps({"kp": kp, "mode": mode, "threads": threads, "ms": round(ms, 4), "alg_GBs": round(gbs, 1),
                      "frac_of_measured_peak": round(gbs / peak, 3), "flag_rate_sample": nflag / (10000 * 64)}))
check(lib().af_seed_scan_config(0, 0))
