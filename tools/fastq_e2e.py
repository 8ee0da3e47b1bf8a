"""Third GPU rate: FASTQ.gz files on disk -> anchored records on the host, through
the C++ reader (zlib, 2-bit packing into pinned tiles) and the streaming GPU pipeline.  The gz
decode is the wall here, not the GPU; the number is reported next to the resident and the
host-buffer rates of bench.py."""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import anchored_fusion_b200 as af  # noqa: E402
from anchored_fusion_b200.stage import scan_fastq_pair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--read-len", type=int, default=150)
args = ap.parse_args()
spec = af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=6783, read_len=args.read_len,
                     frag_mean=2 * args.read_len, sub_ppm=10_000)
anchor = af.synth_anchor(spec)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
d = tempfile.mkdtemp()
p1, p2 = os.path.join(d, "s_1.fastq.gz"), os.path.join(d, "s_2.fastq.gz")
qual = "I" * args.read_len
t0 = time.time()
with gzip.open(p1, "wt", compresslevel=1) as f1, gzip.open(p2, "wt", compresslevel=1) as f2:
    for lo in range(0, args.pairs, 100_000):
        m1, m2 = af.synth_pairs_host(spec, lo, min(100_000, args.pairs - lo))
        a1, a2 = lut[m1], lut[m2]
        f1.write("".join("@frag%d/1\n%s\n+\n%s\n" % (lo + i, a1[i].tobytes().decode(), qual) for i in range(len(a1))))
        f2.write("".join("@frag%d/2\n%s\n+\n%s\n" % (lo + i, a2[i].tobytes().decode(), qual) for i in range(len(a2))))
gen_s = time.time() - t0
index = af.AnchorIndex(anchor)
eng = af.Anchorer(index, 0)
scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 20)          # warm-up (staging buffers, page cache)
t0 = time.time()
anchored, mates, stats = scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 20)
dt = time.time() - t0
print(json.dumps({"metric": "fastq_gz_end_to_end_pairs_per_s", "value": args.pairs / dt, "pairs": args.pairs,
                  "seconds": dt, "anchored_reads": len(anchored), "half_anchored_pairs": len(mates),
                  "gz_bytes": os.path.getsize(p1) + os.path.getsize(p2), "host_cores": os.cpu_count(),
                  "decode_threads": "2 inflate + 2 parse/pack", "fixture_generation_s": gen_s}))
