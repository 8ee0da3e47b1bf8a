#!/usr/bin/env python
"""bench.py -- read pairs/s anchored on B200, seed-scan HBM roofline, CPU baseline.

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the reference arm on the host cores
  python bench.py --workload config3 ...                     100 M pairs read-sharded over the ranks (strong scaling)
  python bench.py --workload singlecell ...                  configs[4]: 20 M pairs in thousands of per-cell FASTQ.gz pairs

A step = one pass of the hot path (seed scan -> compaction -> verify -> extend -> compaction)
over one batch of synthetic 2x150 bp pairs resident in HBM (BASELINE.json configs[1]:
10M pairs against a 6 783 bp anchored CDS cut from a random 10 Mbp reference).  Under torchrun
every rank owns its own 10M-pair shard (weak scaling) and every rank receives all ranks' hit
lists: by default the hit-compaction kernel stores them into a log on every GPU over NVLink
peer memory (--exchange p2p), alternatively one NCCL all-gather per step (--exchange nccl).
See DESIGN.md "Measurement".
"""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read_pairs_per_s_anchored"


def alg_bytes(read_len):
    return 2 * ((2 * read_len + 7) // 8)        # 76 B for 2x150, SURVEY.md 8d


def spec_kwargs(args):
    return dict(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=args.anchor_len, read_len=args.read_len,
                frag_mean=2 * args.read_len, frag_sd=30, sub_ppm=args.sub_ppm, fusion_ppm=args.fusion_ppm)


def pairs_per_gpu(args, n_gpus):
    if args.workload == "config3":
        return (args.total_pairs // n_gpus + 31) // 32 * 32
    return args.pairs


def config_dict(args, n_gpus):
    """The workload, identical for both arms (`--impl reference` included): nothing about HOW it is run."""
    n = pairs_per_gpu(args, n_gpus)
    if args.workload == "config3":
        what = "configs[2]: %d synthetic 2x%d bp pairs, single anchor, read-sharded over %d GPU(s) (%d pairs each)" % (
            args.total_pairs, args.read_len, n_gpus, n)
    elif args.workload == "singlecell":
        what = "configs[4]: single-cell layout, %d cells x %d synthetic 2x%d bp pairs as per-cell FASTQ.gz pairs, cells dealt to %d GPU(s)" % (
            args.cells, args.pairs_per_cell, args.read_len, n_gpus)
    else:
        what = "configs[1]: %d synthetic 2x%d bp pairs per GPU" % (n, args.read_len)
    return {"workload": what + ", one %d bp anchored CDS on a random 10 Mbp reference (seeded generator, %d ppm substitutions, "
                               "%d ppm fusion fragments)" % (args.anchor_len, args.sub_ppm, args.fusion_ppm),
            "pairs_per_gpu": n, "read_len": args.read_len, "anchor_len": args.anchor_len,
            "sharding": "reads x%d, hit lists delivered to every rank" % n_gpus,
            "l2_policy": "input per step (%.0f MB) exceeds the 126 MB L2; no flush needed" % (n * 80 / 1e6)}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (pynvml == nvidia-smi's source)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the CPU arm: oracle/ only (its own generator, af_synth.cpp: no product library is loaded) ---------------
def cpu_sample(args, n_pairs, first_pair=0):
    """Pairs [first_pair, first_pair + n_pairs) of the workload as base codes (generated by all cores)."""
    from oracle import oracle
    spec = oracle.synth_spec(**spec_kwargs(args))
    anchor = oracle.encode(oracle.synth_anchor(spec))
    threads = os.cpu_count() or 1
    reads = oracle.synth_reads(spec, first_pair, n_pairs, threads=threads)
    oracle.anchor_reads(anchor, reads[:2000], threads=threads)   # warm the library
    return anchor, reads


def cpu_reference_rate(n_pairs, threads, sample):
    """oracle/af_oracle.c (a port: the reference's own path is bwa/samtools, absent here and on the GPU box)."""
    from oracle import oracle
    anchor, reads = sample
    t0 = time.perf_counter()
    hits = oracle.anchor_reads(anchor, reads, threads=threads)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, hits


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_pairs
    data = cpu_sample(args, sample)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, dt, _ = cpu_reference_rate(sample, threads, data)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = sample * len(rates) / sum(dt for _, dt in rates)
    desc = "%d of the workload's pairs per step, oracle/af_oracle.c with %d OpenMP threads" % (sample, threads)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(dt for _, dt in rates) / len(rates),
                      "higher_is_better": True, "scaling": "strong" if args.workload == "config3" else "weak",
                      "vs_baseline": None, "dtype": "u8",
                      "data": "synthetic", "config": config_dict(args, args.gpus),
                      "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc},
                      "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---- FASTQ files on disk -> records (the third rate) --------------------------------------------
def fastq_rates(args, index, eng):
    """A bounded sample of the workload written as FASTQ files by oracle/af_synth.cpp (Illumina-style names,
    binned qualities), then read back through the whole host path (task-parallel reader, packer, pinned staging,
    GPU pipeline, record retrieval) once per format."""
    from oracle import oracle
    from anchored_fusion_b200.stage import scan_fastq_pair
    n = args.fastq_pairs
    spec = oracle.synth_spec(**spec_kwargs(args))
    d = tempfile.mkdtemp(prefix="af_bench_fq_")
    out = {}
    try:
        for key, fmt, ext, level in (("bgzf", oracle.FASTQ_BGZF, ".fastq.gz", 1), ("gzip", oracle.FASTQ_GZIP, ".fastq.gz", 1),
                                     ("plain", oracle.FASTQ_PLAIN, ".fastq", 0)):
            p1, p2 = os.path.join(d, key + "_1" + ext), os.path.join(d, key + "_2" + ext)
            oracle.synth_fastq(spec, 0, n, [p1], [p2], fmt, level, threads=max(2, min(8, os.cpu_count() or 2)))
            scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 19, threads=args.threads)      # warm-up: staging, page cache
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                anchored, mates, stats = scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 19, threads=args.threads)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[key] = {"value": n / best, "unit": "pairs/s", "pairs": n, "seconds": best, "anchored_reads": len(anchored),
                        "file_bytes": os.path.getsize(p1) + os.path.getsize(p2)}
            os.remove(p1)
            os.remove(p2)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    threads = args.threads or (os.cpu_count() or 1)
    res = dict(out["bgzf"])
    res.update({"format": "BGZF-compressed .fastq.gz (bgzip): independent <= 64 KB blocks, inflated in parallel",
                "threads": threads, "host_cores": os.cpu_count(),
                "path": "two FASTQ files -> %d-worker reader (own DEFLATE decoder, newline index, CRC, SIMD 2-bit pack) -> pinned "
                        "tiles -> af_pipeline_run -> records" % threads,
                "single_member_gzip": dict(out["gzip"], format="one gzip member per file (gzip / bcl2fastq style): the serial bit "
                                           "stream is cut at block starts found by bit search and decoded by several workers "
                                           "(af_inflate_par.h; AF_GZIP_SERIAL=1 for one inflate thread per file)"),
                "plain_text": out["plain"],
                "input": "oracle/af_synth.cpp: Illumina-style read names, binned qualities (F : , #), deflate level 1"})
    return res


# ---- genome pass of the contiguity filter (the stage after the path; csrc/af_genome.cu) --------------------------
def genome_pass_rate(args, device, hbm_peak):
    """`--genome-reads` reads drawn from a synthetic genome of `--genome-bases` bases (1 % substitutions, both strands)
    aligned back to it through Genome.align (wall clock, best of 3, host text -> records on the host); every read must
    come back where it was drawn.  On a 10 Mbp genome the records are also compared with the CPU oracle."""
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.genome import Genome
    from oracle import oracle
    rng = np.random.default_rng(17)
    rc = str.maketrans("ACGT", "TGCA")
    L = args.read_len

    def draw(n_bases, n_reads, seed):
        reads, where = [], []
        for _ in range(n_reads):
            at = int(rng.integers(0, n_bases - L))
            s = af.synth_anchor(af.synth_spec(seed=seed, ref_len=n_bases, anchor_start=at, anchor_len=L))
            b = bytearray(s if isinstance(s, bytes) else s.encode())
            for i in np.flatnonzero(rng.random(L) < 0.01):
                b[i] = b"ACGT"[(b"ACGT".index(b[i]) + 1 + int(rng.integers(0, 3))) % 4]
            s = b.decode()
            strand = int(rng.integers(0, 2))
            reads.append(s.translate(rc)[::-1] if strand else s)
            where.append((at, strand))
        return reads, where

    g = Genome.synthetic(23, args.genome_bases, device)
    reads, where = draw(args.genome_bases, args.genome_reads, 23)
    g.align(reads[:8])
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        hits = g.align(reads)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, dict(g.last_stats))
    by = {int(h["read_id"]): h for h in hits}
    placed = sum(1 for i, (at, strand) in enumerate(where) if i in by and (int(by[i]["score_strand"]) & 1) == strand
                 and int(by[i]["pos"]) - int(by[i]["clip_l"]) == at + 256 + 1)
    g.close()
    # parity on a genome the oracle can index in a few seconds
    nb = 10_000_000
    g2 = Genome.synthetic(29, nb, device)
    reads2, _ = draw(nb, 400, 29)
    h2 = g2.align(reads2)
    g2.close()
    ref = af.synth_anchor(af.synth_spec(seed=29, ref_len=nb, anchor_start=0, anchor_len=nb))
    concat = oracle.encode(b"N" * 256 + (ref if isinstance(ref, bytes) else ref.encode()) + b"N" * 256)
    o = oracle.anchor_reads(concat, np.stack([oracle.encode(r) for r in reads2]), threads=os.cpu_count() or 1)
    equal = len(o) == len(h2) and all(np.array_equal(o[f].astype(np.int64), h2[f].astype(np.int64))
                                      for f in ("read_id", "pos", "clip_l", "m_len", "clip_r", "score_strand"))
    st = best[1]
    scan_bytes = args.genome_bases / 4.0 * st["n_passes"]
    return {"what": "Genome.align: reads vs a synthetic genome resident in HBM (2 bit/base), no index; replaces `bwa mem` on the genome "
                    "inside del_too_many_reads", "genome_bases": args.genome_bases, "reads": args.genome_reads, "read_len": L,
            "value": args.genome_reads / best[0], "unit": "reads/s", "ms_per_call": best[0] * 1e3, "passes": st["n_passes"],
            "scan_ms_per_pass": st["scan_ms"] / max(st["n_passes"], 1),
            "scan_gb_per_s": scan_bytes / (st["scan_ms"] * 1e-3) / 1e9 if st["scan_ms"] else None,
            "scan_frac_of_hbm_peak": scan_bytes / (st["scan_ms"] * 1e-3) / 1e9 / hbm_peak if st["scan_ms"] else None,
            "reads_back_where_drawn": placed,
            "parity": {"genome_bases": nb, "reads": len(reads2), "records": int(len(o)), "equal": bool(equal),
                       "compared": "records of 400 reads on a 10 Mbp genome vs oracle/af_oracle.c with the concatenated genome as its anchor"}}


# ---- configs[4]: the single-cell layout --------------------------------------------------------------------------
def run_singlecell(args):
    """Thousands of per-cell FASTQ.gz pairs through the single-cell stage (all of a rank's cells through one
    reader, shared GPU batches, per-cell files).  Timed: decode -> pack -> H2D -> kernels -> D2H -> files."""
    import torch
    from anchored_fusion_b200.stage import GeneAnchorer, anchor_cells
    from oracle import oracle
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    spec = oracle.synth_spec(**spec_kwargs(args))
    root = tempfile.mkdtemp(prefix="af_bench_sc_%d_" % rank)
    try:
        cells_dir, out_dir = os.path.join(root, "cells"), os.path.join(root, "out")
        os.makedirs(cells_dir)
        mine = list(range(rank, args.cells, world))
        f1 = [os.path.join(cells_dir, "cell%06d_1.fastq.gz" % c) for c in mine]
        f2 = [os.path.join(cells_dir, "cell%06d_2.fastq.gz" % c) for c in mine]
        t0 = time.perf_counter()
        cores = max(1, (os.cpu_count() or 1) // world)
        # cell c holds pairs [c * ppc, (c + 1) * ppc) of the generator; one generator call writes a run of
        # consecutive cells (all of them when there is one rank), calls run side by side otherwise
        ppc = args.pairs_per_cell
        if world == 1:
            for k in range(0, len(mine), 512):
                oracle.synth_fastq(spec, mine[k] * ppc, ppc, f1[k: k + 512], f2[k: k + 512], oracle.FASTQ_GZIP, 1, threads=cores)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=max(1, cores // 2)) as pool:
                list(pool.map(lambda k: oracle.synth_fastq(spec, mine[k] * ppc, ppc, [f1[k]], [f2[k]], oracle.FASTQ_GZIP, 1, threads=2),
                              range(len(mine))))
        gen_s = time.perf_counter() - t0
        fa = os.path.join(root, "anchor.fa")
        with open(fa, "w") as fh:
            fh.write(">GENE0\n" + oracle.synth_anchor(spec).decode() + "\n")
        ga = GeneAnchorer(fa, str(local), "GENE0")

        def prefix_of(gene, cell):
            dd = os.path.join(out_dir, gene, "work_dir", cell)
            os.makedirs(dd, exist_ok=True)
            return os.path.join(dd, gene + "_fusion")

        cell_files = [("cell%06d" % c, a, b) for c, a, b in zip(mine, f1, f2)]
        anchor_cells([ga], cell_files[: min(len(cell_files), 64)], prefix_of, thread=str(args.threads))     # warm-up: buffers, page cache
        shutil.rmtree(out_dir, ignore_errors=True)
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local)
        sampler.start()
        t0 = time.perf_counter()
        res = anchor_cells([ga], cell_files, prefix_of, thread=str(args.threads))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        totals = [res["pairs"], res["anchored"], dt, res["seconds_scan"], res["seconds_write"]]
        if world > 1:
            box = [None] * world
            dist.all_gather_object(box, totals)
            totals = [sum(b[0] for b in box), sum(b[1] for b in box), max(b[2] for b in box), max(b[3] for b in box), max(b[4] for b in box)]
        if rank == 0:
            from anchored_fusion_b200._lib import lib
            pairs, anchored, dt, t_scan, t_write = totals
            value = pairs / dt
            return ({
                "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_dict(args, world),
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": int(pairs * 80), "d2h_bytes_per_step": int(anchored * 16),
                        "path": "per-cell FASTQ.gz pairs -> one reader per rank (cells decoded concurrently, packed back to back) -> "
                                "shared GPU batches -> hit lists split per cell -> per-cell BAM / FASTQ / SAM files"},
                "singlecell": {"cells": args.cells, "pairs_per_cell": args.pairs_per_cell, "anchored_reads": anchored,
                               "seconds_decode_and_gpu": t_scan, "seconds_per_cell_files": t_write,
                               "ms_per_cell_files": 1e3 * t_write / max(len(mine), 1), "cells_per_s": args.cells / dt,
                               "reader_threads_per_rank": res["threads"], "fixture_generation_s": gen_s},
                "roofline": None, "cpu_baseline": None, "gpu_launches": int(lib().af_kernel_launches()), "clocks": clocks})
        return None
    finally:
        shutil.rmtree(root, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config1", choices=["config1", "config3", "singlecell"],
                    help="config1: BASELINE configs[1], 10 M pairs per GPU per step (weak scaling, default); config3: configs[2], "
                         "--total-pairs read-sharded over the ranks (strong scaling); singlecell: configs[4] through the single-cell stage")
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step (config1)")
    ap.add_argument("--total-pairs", type=int, default=100_000_000, help="pairs of the whole job (config3)")
    ap.add_argument("--cells", type=int, default=4000)
    ap.add_argument("--pairs-per-cell", type=int, default=5000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--anchor-len", type=int, default=6783)
    ap.add_argument("--sub-ppm", type=int, default=10_000)
    ap.add_argument("--fusion-ppm", type=int, default=0)
    ap.add_argument("--kp", type=int, default=0)
    ap.add_argument("--scan-threads", type=int, default=0)
    ap.add_argument("--scan-mode", type=str, default="0", help="af_seed_scan_config mode(s), comma separated")
    ap.add_argument("--slots", type=int, default=3, help="workspace slots / streams consecutive steps alternate between")
    ap.add_argument("--graphs", type=int, default=0, help="1: replay one CUDA graph per slot in the throughput region")
    ap.add_argument("--cand-cap", type=int, default=0, help="candidate capacity per batch (default 2 x pairs: every read)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="iterations of the host-buffer end-to-end region (0: --steps, at most 50)")
    ap.add_argument("--e2e-format", choices=["wire", "tiles"], default="wire",
                    help="host buffers of the end-to-end region: the wire format (76 B per 2x150 pair, expanded on the GPU) or whole tiles (80 B)")
    ap.add_argument("--cpu-pairs", type=int, default=10_000_000, help="bounded sample for cpu_baseline")
    ap.add_argument("--ref-pairs", type=int, default=10_000_000, help="pairs per step of the reference arm")
    ap.add_argument("--parity-pairs", type=int, default=1_000_000, help="pairs per rank compared with the oracle after the clock (0: skip)")
    ap.add_argument("--threads", type=int, default=0, help="FASTQ reader workers (0: one per host core)")
    ap.add_argument("--gather-cap", type=int, default=0, help="hit records per rank in the per-step all-gather "
                    "(0: sized from a probe pass, 1.25 x the largest per-rank hit count, rounded up to 4096)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the ranks' hit lists reach every rank -- p2p: the hit-compaction kernel stores them "
                         "into every GPU's log over NVLink peer memory; nccl: one all-gather per step")
    ap.add_argument("--fastq-pairs", type=int, default=4_000_000, help="pairs of the FASTQ end-to-end samples (0: skip)")
    ap.add_argument("--genome-bases", type=int, default=3_100_000_000, help="synthetic genome of the genome-pass leg (0: skip)")
    ap.add_argument("--genome-reads", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version
    # there) are sent to stderr while we run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.workload == "singlecell":
        obj = run_singlecell(args)
        if obj is not None:
            emit(obj)
        return

    import torch
    import anchored_fusion_b200 as af
    from anchored_fusion_b200 import dist as afdist
    from anchored_fusion_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = None
    if world > 1:
        import torch.distributed as dist
        cpus = afdist.bind_near_gpu(local)      # host staging buffers on the GPU's own NUMA node
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    modes = [int(m) for m in str(args.scan_mode).split(",") if m and int(m)]
    if args.scan_threads or modes:
        from anchored_fusion_b200._lib import check
        check(L.af_seed_scan_config(args.scan_threads, 0))
        for m in modes:
            check(L.af_seed_scan_config(0, m))

    spec = af.synth_spec(**spec_kwargs(args))
    index = af.AnchorIndex(af.synth_anchor(spec), kp=args.kp)
    eng = af.Anchorer(index, local)
    n = pairs_per_gpu(args, world)
    first_pair = rank * n
    batch = af.synth_pairs_device(spec, first_pair, n, index.pad_byte, local)   # this rank's shard
    torch.cuda.synchronize()
    n_slots = max(1, args.slots)
    cand_cap = args.cand_cap or 2 * n
    if args.workload == "config3":
        cand_cap = args.cand_cap or max(n // 4, 1 << 20)       # 2 % of the reads are flagged; 12.5 % is ample and keeps 3 slots small
    gather_cap = args.gather_cap
    if world > 1 and gather_cap <= 0:
        _, probe = eng.anchor(batch, cand_cap=cand_cap)
        t = torch.tensor([probe["hits"]], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_cap = (int(t.item()) * 5 // 4 + 4095) // 4096 * 4096

    profiling = [False]
    exchange = None
    if world > 1 and args.exchange == "p2p" and not args.graphs:
        # a log region takes the batches of one slot of one rank for a whole timed region
        exchange = afdist.HitExchange(rank, world, n_slots, (max(args.warmup, args.steps) + 1) * (gather_cap + 1), dev)
    step_base = [0]
    streams = []
    for sl in range(n_slots):               # slot 0 runs on the current stream, the others own side streams
        eng.enqueue(batch, cand_cap=cand_cap, slot=sl)
        streams.append(eng.slot_stream(sl) if sl else torch.cuda.current_stream(dev))
    torch.cuda.synchronize()

    graphs = {}
    if args.graphs:
        # one CUDA graph per workspace slot: memsets + kernels of af_anchor_batch, replayed each step
        for sl in range(n_slots):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[sl] if sl else None):
                st = torch.cuda.current_stream(dev)
                eng.enqueue(batch, cand_cap=cand_cap, slot=sl, stream=st)
            graphs[sl] = g
        torch.cuda.synchronize()

    def step(i):
        """One pass of the hot path over this rank's batch.  Consecutive steps are independent
        batches, so they alternate between workspace slots / streams and may overlap, the way a
        run over many batches is pipelined; everything is complete before the clock stops."""
        sl = i % n_slots
        if graphs and not profiling[0]:
            with torch.cuda.stream(streams[sl]):
                graphs[sl].replay()
            hits, counts = eng.counts_and_hits(sl)[2:], eng.counts_and_hits(sl)[:2].view(-1)
        elif exchange is not None:
            # the records reach every rank from inside the last kernel of the path; nothing else to launch
            return eng.enqueue(batch, cand_cap=cand_cap, slot=sl, exchange=exchange,
                               pair_base=((step_base[0] + i) * world + rank) * n)
        else:
            hits, counts = eng.enqueue(batch, cand_cap=cand_cap, slot=sl)
        if world > 1:
            with torch.cuda.stream(streams[sl]):
                return afdist.gather_hits_tensor(eng.counts_and_hits(sl), gather_cap)
        return counts, hits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k):
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for st in streams[1:]:
            st.wait_event(fork)
        for i in range(k):
            out = step(i)
        for st in streams[1:]:
            join = torch.cuda.Event()
            join.record(st)
            cur.wait_event(join)
        return out

    run_steps(args.warmup)
    barrier()
    if exchange is not None:
        exchange.reset()
    stats_counts = eng.counts_and_hits(0)[:2].reshape(-1).cpu().numpy().view(np.uint32)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.af_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.af_kernel_launches() - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    exchange_info = None
    if world > 1:
        exchange_info = {"kind": "nccl all_gather_into_tensor per step", "records_cap_per_rank": gather_cap}
    if exchange is not None:
        # outside the clock: every rank checks that its logs hold the K timed batches of every rank and
        # that each equals what one NCCL all-gather of the ranks' hit lists delivers
        got = exchange.collect()
        with torch.cuda.stream(streams[0]):
            eng.enqueue(batch, cand_cap=cand_cap, slot=0)
            ac, ah = afdist.gather_hits_tensor(eng.counts_and_hits(0), gather_cap)
        torch.cuda.synchronize()
        ac, ah = ac.cpu().numpy(), ah.cpu().numpy()
        for r in range(world):
            mine = [(b, h) for (src, b, h) in got if src == r]
            want_bases = sorted((i * world + r) * n for i in range(args.steps))
            assert sorted(b for b, _ in mine) == want_bases, "rank %d: log of rank %d holds batches %s" % (rank, r, [b for b, _ in mine][:4])
            ref = np.ascontiguousarray(ah[r, : ac[r]]).view(np.uint8).reshape(-1).view(af.HIT_DTYPE)
            for _, h in mine:
                assert len(h) == len(ref) and (h.view(np.uint8) == ref.view(np.uint8)).all(), "rank %d: log of rank %d differs from the all-gather" % (rank, r)
        exchange_info = {"kind": "p2p: k_hit_scatter stores each record into every rank's log (NVLink peer memory, CUDA IPC)",
                         "validated": "each rank's logs hold the %d timed batches of all %d ranks, byte-equal to an NCCL all-gather" % (args.steps, world),
                         "records_per_rank_per_step": int(ac[rank]), "log_records_per_region": exchange.log_cap}
        step_base[0] = args.steps
        exchange.reset()

    # ---- parity, outside the clock: what the ranks DELIVERED for the first --parity-pairs pairs of every shard
    # (through the same exchange the timed steps used) against the CPU oracle on the same pairs; rank 0 judges
    parity = None
    if args.parity_pairs > 0:
        pp = min(args.parity_pairs, n) // 32 * 32
        lay_p = af.layout(args.read_len, pp)
        sub = af.PackedBatch(batch.packed[: lay_p.packed_bytes // 4], pp, args.read_len, args.read_len)
        delivered = {}
        if exchange is not None:
            with torch.cuda.stream(streams[0]):
                eng.enqueue(sub, cand_cap=2 * pp, slot=0, exchange=exchange, pair_base=first_pair)
            for src, base, h in exchange.collect():
                delivered[src] = h.copy()
            exchange.reset()
        elif world > 1:
            with torch.cuda.stream(streams[0]):
                eng.enqueue(sub, cand_cap=2 * pp, slot=0)
                ac, ah = afdist.gather_hits_tensor(eng.counts_and_hits(0), gather_cap)
            torch.cuda.synchronize()
            ac, ah = ac.cpu().numpy(), ah.cpu().numpy()
            for r in range(world):
                delivered[r] = np.ascontiguousarray(ah[r, : ac[r]]).view(np.uint8).reshape(-1).view(af.HIT_DTYPE).copy()
        else:
            delivered[0], _ = eng.anchor(sub)
        if rank == 0:
            from oracle import oracle
            ospec = oracle.synth_spec(**spec_kwargs(args))
            acodes = oracle.encode(oracle.synth_anchor(ospec))
            equal, n_rec = True, 0
            for r in range(world):
                reads = oracle.synth_reads(ospec, r * n, pp, threads=os.cpu_count() or 1)
                want = oracle.anchor_reads(acodes, reads, threads=os.cpu_count() or 1)
                got_r = delivered.get(r)
                ok = got_r is not None and len(got_r) == len(want) and got_r.tobytes() == want.tobytes()
                equal = equal and ok
                n_rec += len(want)
            parity = {"pairs": pp * world, "records": n_rec, "equal": bool(equal),
                      "compared": "16-byte records each rank delivered (%s) for the first %d pairs of its shard vs oracle/af_oracle.c on the same pairs"
                                  % ("p2p hit log as rank 0 holds it" if exchange is not None else ("NCCL all-gather" if world > 1 else "af_anchor_batch"), pp)}
        flag = torch.tensor([1 if (parity is None or parity["equal"]) else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(flag, 0)
        if int(flag.item()) == 0:
            if rank == 0:
                print("bench.py: PARITY FAILURE -- delivered records differ from the oracle: %s" % json.dumps(parity), file=sys.stderr)
            sys.exit(3)

    # Second timed region, the same K steps on ONE stream, with CUDA events recorded on that stream
    # around every stage (af_profile_*): clean per-kernel durations for the roofline.  (In the
    # region above consecutive steps overlap across streams, which is right for throughput but
    # lets another step's small kernels share the SMs with the scan.)
    n_slots_saved, n_slots = n_slots, 1
    streams_saved, streams = streams, streams[:1]
    profiling[0] = True
    L.af_profile_begin()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    run_steps(args.steps)
    s1.record()
    barrier()
    serial_ms_per_step = s0.elapsed_time(s1) / args.steps
    stage_ms = (ctypes.c_double * 5)()
    stage_calls = (ctypes.c_int64 * 5)()
    L.af_profile_end(stage_ms, stage_calls)
    n_slots, streams = n_slots_saved, streams_saved
    profiling[0] = False

    # roofline of the dominant kernel (seed scan): algorithmic bytes / its mean launch duration,
    # CUDA events on the launching stream, inside the timed region above
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    scan_ms = stage_ms[0] / max(stage_calls[0], 1)
    achieved = alg_bytes(args.read_len) * n / (scan_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "seed_scan_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("pairs") == n and tj.get("read_len") == args.read_len:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_seed_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_pair": alg_bytes(args.read_len), "ms_per_launch": scan_ms,
                "measured": "CUDA events on the launching stream around each of the %d launches of a second timed "
                            "region (same steps, one stream)" % stage_calls[0],
                "serial_ms_per_step": serial_ms_per_step,
                "whole_step_frac": alg_bytes(args.read_len) * n / (ms_per_step * 1e-3) / 1e9 / peak,
                "stage_ms_per_step": ({"seed_scan": stage_ms[0] / args.steps, "flag_compaction": stage_ms[1] / args.steps,
                                       "verify": stage_ms[2] / args.steps, "extend": stage_ms[3] / args.steps,
                                       "hit_compaction": stage_ms[4] / args.steps} if stage_calls[3] else
                                      {"seed_scan_emit": stage_ms[0] / args.steps,
                                       "tail_verify_extend_place": stage_ms[2] / args.steps}),
                "kernels_per_step": 6 if stage_calls[3] else 2}

    # end to end through the host-buffer C-ABI call: pinned host batch -> H2D -> kernels -> D2H hits
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 50)
        lay = af.layout(args.read_len, n)
        hptr = L.af_host_alloc(lay.packed_bytes)
        host_packed = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint32)), (lay.packed_bytes // 4,))
        host_packed[:] = batch.packed.cpu().numpy().view(np.uint32)
        wire = args.e2e_format == "wire"
        h2d_bytes = int(lay.packed_bytes)
        wptr = None
        if wire:
            # the library's wire format: 4 L bits per pair (76 bytes for 2 x 150 bp) instead of whole tiles (80); the
            # conversion is input preparation, outside the clock -- a packer would write this format directly
            h2d_bytes = int(af.wire_bytes(args.read_len, n))
            wptr = L.af_host_alloc(h2d_bytes)
            host_wire = np.ctypeslib.as_array(ctypes.cast(wptr, ctypes.POINTER(ctypes.c_uint32)), (h2d_bytes // 4,))
            af.wire_from_packed(host_packed, args.read_len, n, out=host_wire)
            src_words = host_wire
        else:
            src_words = host_packed
        hb = af.PackedBatch(src_words, n, args.read_len, args.read_len)
        hits_out = np.zeros(max(1 << 20, n // 16), dtype=af.HIT_DTYPE)
        for _ in range(3):
            eng.anchor_host(hb, hits_out=hits_out, wire=wire)              # warm-up (allocates the slots)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h, st = eng.anchor_host(hb, hits_out=hits_out, wire=wire)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        if args.parity_pairs and wire:                          # the wire path delivers what the resident path delivers
            ref_hits, _ = eng.anchor(batch)
            if not (len(ref_hits) == len(h) and ref_hits.tobytes() == np.ascontiguousarray(h).tobytes()):
                print("bench: the wire-format pipeline and the resident path disagree", file=sys.stderr)
                sys.exit(3)
        # the ceiling the fabric in front of the GPUs sets for this: the same tiles through cudaMemcpyAsync alone,
        # all ranks copying at once
        h2d_stream = torch.cuda.Stream(device=dev)
        pinned_view = torch.empty(src_words.shape, dtype=torch.int32, pin_memory=True)
        pinned_view.numpy()[:] = src_words.view(np.int32)
        dst = torch.empty(src_words.shape, dtype=torch.int32, device=dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(h2d_stream):
            dst.copy_(pinned_view, non_blocking=True)
        barrier()
        with torch.cuda.stream(h2d_stream):
            c0.record()
            for _ in range(max(3, e2e_steps // 4)):
                dst.copy_(pinned_view, non_blocking=True)
            c1.record()
        barrier()
        copy_ms = c0.elapsed_time(c1) / max(3, e2e_steps // 4)
        if world > 1:
            t = torch.tensor([dt, copy_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, copy_ms = float(t[0].item()), float(t[1].item())
        e2e = {"value": world * n / dt, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": int(len(h) * 16 + 32 * ((n + (1 << 20) - 1) >> 20)), "ms_per_step": dt * 1e3,
               "steps": e2e_steps,
               "path": ("af_pipeline_run_wire: pinned host batch in wire format (4 L bits per pair) -> cudaMemcpyAsync -> expansion "
                        "into tiles on the GPU -> kernels -> hit list on host") if wire else
                       "af_pipeline_run: pinned host tiles -> cudaMemcpyAsync -> kernels -> hit list on host",
               "format": args.e2e_format,
               "h2d_only_ceiling": {"ms_per_step": copy_ms, "pairs_per_s": world * n / (copy_ms * 1e-3),
                                    "gb_per_s_all_ranks": world * h2d_bytes / (copy_ms * 1e-3) / 1e9,
                                    "what": "the same %d MB per rank through cudaMemcpyAsync alone, all %d rank(s) at once, max over ranks"
                                            % (h2d_bytes // 1_000_000, world)},
               "host_cpus_rank0": ("%d CPUs local to the GPU (NVML affinity)" % len(cpus)) if cpus else "unbound"}
        del dst, pinned_view
        eng.close_pipeline()
        L.af_host_free(hptr)
        if wptr:
            L.af_host_free(wptr)

    # third rate: FASTQ files on disk -> reader -> packer -> pipeline -> records
    fastq = None
    if rank == 0 and world == 1 and args.fastq_pairs > 0 and not args.no_e2e and args.workload == "config1":
        fastq = fastq_rates(args, index, eng)

    # the next stage: genome pass of the contiguity filter -- 2-op reads against a genome resident in HBM
    genome = None
    if rank == 0 and world == 1 and args.genome_bases > 0 and not args.no_e2e and args.workload == "config1":
        genome = genome_pass_rate(args, local, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        data = cpu_sample(args, args.cpu_pairs)
        rate, dt, _ = cpu_reference_rate(args.cpu_pairs, threads, data)
        n1 = min(args.cpu_pairs, 2_000_000)                      # and one core alone, on the first 2 M pairs
        rate1, dt1, _ = cpu_reference_rate(n1, 1, (data[0], data[1][: 2 * n1]))
        cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "first %d pairs of the workload, oracle/af_oracle.c, %d OpenMP threads, %.1f s"
                         % (args.cpu_pairs, threads, dt),
               "value_1_core": rate1, "sample_1_core": "first %d pairs, 1 thread, %.1f s" % (n1, dt1)}

    if rank == 0:
        nh = int(stats_counts[1])
        emit({"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
              "scaling": "strong" if args.workload == "config3" else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
              "config": config_dict(args, world),
              "run": {"streams": n_slots, "cuda_graphs": bool(args.graphs), "exchange": exchange_info},
              "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "fastq_gz": fastq, "genome_pass": genome, "parity": parity,
              "gpu_launches": int(launches), "clocks": clocks,
              "per_step": {"flagged_reads": int(stats_counts[0]), "seeded_reads": int(stats_counts[3]),
                           "anchored_reads": nh, "kp": index.info.kp, "stride": index.info.stride}})
    if exchange is not None:
        barrier()
        exchange.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
