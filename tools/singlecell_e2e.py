"""Config-5 shape end to end: many small per-cell FASTQ.gz pairs -> the single-cell driver.
Reports cells/s and pairs/s of this process; under torchrun every rank takes its share of the cells
(no exchange) and rank 0 prints the sum.

  python tools/singlecell_e2e.py --cells 400 --pairs-per-cell 5000
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/singlecell_e2e.py ...
"""
import argparse
import gzip
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import anchored_fusion_b200 as af  # noqa: E402
from anchored_fusion_b200.cli import main_singlecell  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=400)
ap.add_argument("--pairs-per-cell", type=int, default=5000)
ap.add_argument("--genes", type=int, default=1)
ap.add_argument("--dir", type=str, default="")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
root = args.dir or os.path.join(tempfile.gettempdir(), "af_sc_e2e")
cells_dir, out_dir = os.path.join(root, "cells"), os.path.join(root, "out")
spec = af.synth_spec(seed=5, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=6783, read_len=150,
                     frag_mean=300, sub_ppm=10_000, fusion_ppm=2_000)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(cells_dir)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    qual = "F" * 150
    n = args.pairs_per_cell
    for c in range(args.cells):
        m1, m2 = af.synth_pairs_host(spec, c * n, n)
        a1, a2 = lut[m1], lut[m2]
        for mate, a in ((1, a1), (2, a2)):
            with gzip.open(os.path.join(cells_dir, "cell%05d_%d.fastq.gz" % (c, mate)), "wt", compresslevel=1) as fh:
                fh.write("".join("@c%d_%d/%d\n%s\n+\n%s\n" % (c, i, mate, a[i].tobytes().decode(), qual) for i in range(n)))
    anchor = af.synth_anchor(spec).decode()
    with open(os.path.join(root, "genes.fa"), "w") as fh:
        for g in range(args.genes):
            seq = anchor if g == 0 else anchor[g * 300:] + anchor[: g * 300]
            fh.write(">NM_%d.1 GENE%d [organism=synthetic]\n%s\n" % (g, g, seq))
if world > 1:
    dist.barrier()
t0 = time.time()
main_singlecell(["--file_anchored_cds", os.path.join(root, "genes.fa"), "--fastq_dir", cells_dir, "--out_folder", out_dir])
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    total = args.cells * args.pairs_per_cell
    print(json.dumps({"metric": "singlecell_fastq_gz_end_to_end", "cells": args.cells, "pairs_per_cell": args.pairs_per_cell,
                      "genes": args.genes, "ranks": world, "seconds": dt, "cells_per_s": args.cells / dt, "pairs_per_s": total / dt}))
if world > 1:
    dist.destroy_process_group()
