"""Host ingest alone (no GPU): FASTQ files -> 2-bit packed tiles through the task-parallel reader, per input
format and worker count.  The files come from oracle/af_synth.cpp (Illumina-style names, binned qualities).

  python tools/ingest_bench.py --pairs 2000000 --threads 1,2,4,8,16 --out profiles/r02_ingest.json
"""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import anchored_fusion_b200 as af  # noqa: E402
from anchored_fusion_b200._lib import check, lib  # noqa: E402
from oracle import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2_000_000)
ap.add_argument("--threads", type=str, default="1,2,4,8,16")
ap.add_argument("--level", type=int, default=6)
ap.add_argument("--batch", type=int, default=1 << 19)
ap.add_argument("--out", type=str, default="")
ap.add_argument("--real-gzip", type=int, default=0, help="also time a single member written by zlib itself at this level (one continuous stream, back-references across every chunk)")
args = ap.parse_args()
L = lib()
spec = oracle.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=6783, read_len=150, frag_mean=300, sub_ppm=10_000)
d = tempfile.mkdtemp(prefix="af_ingest_")
res = {"pairs": args.pairs, "host_cores": os.cpu_count(), "deflate_level": args.level, "batch_pairs": args.batch, "formats": {}}


def run(p1, p2, threads):
    h = ctypes.c_void_p()
    check(L.af_fastq_open_threads(p1.encode(), p2.encode(), threads, ctypes.byref(h)))
    lay = af.layout(160, args.batch)
    packed = np.zeros(lay.packed_bytes // 4, dtype=np.uint32)
    lens = np.zeros(2 * args.batch, np.uint16)
    nids = np.zeros(2 * args.batch, np.uint32)
    nmask = np.zeros((2 * args.batch, 8), np.uint32)
    nn, ul, n = ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_int64(0)
    t0 = time.perf_counter()
    tot = 0
    while True:
        check(L.af_fastq_next(h, args.batch, 160, 0xE4, packed.ctypes.data, lens.ctypes.data, nids.ctypes.data, nmask.ctypes.data,
                              len(nids), ctypes.byref(nn), ctypes.byref(ul), ctypes.byref(n)))
        if n.value == 0:
            break
        tot += n.value
    dt = time.perf_counter() - t0
    L.af_fastq_close(h)
    assert tot == args.pairs
    return dt


try:
    for key, fmt, ext in (("bgzf", oracle.FASTQ_BGZF, ".fastq.gz"), ("gzip", oracle.FASTQ_GZIP, ".fastq.gz"), ("plain", oracle.FASTQ_PLAIN, ".fastq")):
        p1, p2 = os.path.join(d, key + "_1" + ext), os.path.join(d, key + "_2" + ext)
        oracle.synth_fastq(spec, 0, args.pairs, [p1], [p2], fmt, args.level, threads=2)
        run(p1, p2, 0)                                       # page cache, buffer cache
        rows = {}
        for t in [int(x) for x in args.threads.split(",")]:
            best = min(run(p1, p2, t) for _ in range(3))
            rows[str(t)] = {"seconds": best, "pairs_per_s": args.pairs / best}
            print("%-6s threads=%-2d %.3f s  %.2f M pairs/s" % (key, t, best, args.pairs / best / 1e6), file=sys.stderr, flush=True)
        res["formats"][key] = {"file_bytes": os.path.getsize(p1) + os.path.getsize(p2), "by_threads": rows}
        os.remove(p1)
        os.remove(p2)
    if args.real_gzip:
        import zlib
        p1, p2 = os.path.join(d, "z_1.fastq.gz"), os.path.join(d, "z_2.fastq.gz")
        q1, q2 = os.path.join(d, "z_1.fastq"), os.path.join(d, "z_2.fastq")
        oracle.synth_fastq(spec, 0, args.pairs, [q1], [q2], oracle.FASTQ_PLAIN, 0, threads=8)
        for src, dst in ((q1, p1), (q2, p2)):
            co = zlib.compressobj(args.real_gzip, zlib.DEFLATED, 31)
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                while True:
                    blk = fi.read(1 << 24)
                    if not blk:
                        break
                    fo.write(co.compress(blk))
                fo.write(co.flush())
            os.remove(src)
        run(p1, p2, 0)
        rows = {}
        for label, env in (("parallel", None), ("serial", "1")):
            if env:
                os.environ["AF_GZIP_SERIAL"] = env
            else:
                os.environ.pop("AF_GZIP_SERIAL", None)
            for t in [int(x) for x in args.threads.split(",")]:
                best = min(run(p1, p2, t) for _ in range(3))
                rows["%s_%d" % (label, t)] = {"seconds": best, "pairs_per_s": args.pairs / best}
                print("zlib-%d %-8s threads=%-2d %.3f s  %.2f M pairs/s" % (args.real_gzip, label, t, best, args.pairs / best / 1e6), file=sys.stderr, flush=True)
        os.environ.pop("AF_GZIP_SERIAL", None)
        res["formats"]["zlib_single_member"] = {"level": args.real_gzip, "file_bytes": os.path.getsize(p1) + os.path.getsize(p2), "by_mode_threads": rows}
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(res))
if args.out:
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
