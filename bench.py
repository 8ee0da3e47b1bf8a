#!/usr/bin/env python
"""bench.py -- read pairs/s anchored on B200, seed-scan HBM roofline, CPU baseline.

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the reference arm on the host cores

A step = one pass of the hot path (seed scan -> compaction -> verify/extend -> compaction)
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
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_PAIR = {150: 76, 101: 52}   # 2 * ceil(2L/8), SURVEY.md 8d
METRIC = "read_pairs_per_s_anchored"


def alg_bytes(read_len):
    return 2 * ((2 * read_len + 7) // 8)


def workload(args):
    import anchored_fusion_b200 as af
    return af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=args.anchor_len,
                         read_len=args.read_len, frag_mean=2 * args.read_len, frag_sd=30, sub_ppm=args.sub_ppm,
                         fusion_ppm=args.fusion_ppm)


def config_dict(args, n_gpus):
    return {"workload": "configs[1]: %d synthetic 2x%d bp pairs per GPU, one %d bp anchored CDS on a random 10 Mbp "
                        "reference (seeded generator, %d ppm substitutions, %d ppm fusion fragments)"
                        % (args.pairs, args.read_len, args.anchor_len, args.sub_ppm, args.fusion_ppm),
            "pairs_per_gpu": args.pairs, "read_len": args.read_len, "anchor_len": args.anchor_len,
            "sharding": "reads x%d, hit lists delivered to every rank" % n_gpus,
            "l2_policy": "input per step (%.0f MB) exceeds the 126 MB L2; no flush needed" % (args.pairs * 80 / 1e6)}


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


def cpu_sample(args, n_pairs):
    """The first n_pairs pairs of the workload as base codes for the CPU arm (generated once, by all cores)."""
    from concurrent.futures import ThreadPoolExecutor
    import anchored_fusion_b200 as af
    from oracle import oracle
    spec = workload(args)
    anchor = oracle.encode(af.synth_anchor(spec))
    reads = np.empty((2 * n_pairs, spec.read_len), dtype=np.uint8)
    threads = os.cpu_count() or 1
    step = (n_pairs + threads - 1) // threads

    def fill(lo):
        hi = min(lo + step, n_pairs)
        m1, m2 = af.synth_pairs_host(spec, lo, hi - lo)          # a ctypes call: runs without the GIL
        reads[2 * lo: 2 * hi: 2], reads[2 * lo + 1: 2 * hi: 2] = m1, m2

    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(fill, range(0, n_pairs, step)))
    oracle.anchor_reads(anchor, reads[:2000], threads=threads)   # warm the library
    return anchor, reads


def cpu_reference_rate(args, n_pairs, threads, sample=None):
    """The CPU arm: oracle/af_oracle.c (a port: the reference's own path is bwa/samtools, absent)."""
    from oracle import oracle
    anchor, reads = sample if sample is not None else cpu_sample(args, n_pairs)
    t0 = time.perf_counter()
    hits = oracle.anchor_reads(anchor, reads, threads=threads)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, len(hits)


def fastq_gz_rate(args, spec, index, eng):
    """A bounded sample of the workload written as two FASTQ.gz files, then read back through the whole
    host path (inflate threads, parser, packer, pinned staging, GPU pipeline, record retrieval)."""
    import gzip
    import shutil
    import tempfile
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import scan_fastq_pair
    n = args.fastq_pairs
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    d = tempfile.mkdtemp(prefix="af_bench_fq_")
    try:
        p1, p2 = os.path.join(d, "s_1.fastq.gz"), os.path.join(d, "s_2.fastq.gz")
        qual = "I" * args.read_len
        with gzip.open(p1, "wt", compresslevel=1) as f1, gzip.open(p2, "wt", compresslevel=1) as f2:
            for lo in range(0, n, 100_000):
                m1, m2 = af.synth_pairs_host(spec, lo, min(100_000, n - lo))
                a1, a2 = lut[m1], lut[m2]
                f1.write("".join("@frag%d/1\n%s\n+\n%s\n" % (lo + i, a1[i].tobytes().decode(), qual) for i in range(len(a1))))
                f2.write("".join("@frag%d/2\n%s\n+\n%s\n" % (lo + i, a2[i].tobytes().decode(), qual) for i in range(len(a2))))
        scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 20)          # warm-up: staging buffers, page cache
        t0 = time.perf_counter()
        anchored, mates, stats = scan_fastq_pair(index, p1, p2, engine=eng, batch_pairs=1 << 20)
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "pairs/s", "pairs": n, "seconds": dt, "anchored_reads": len(anchored),
                "gz_bytes": os.path.getsize(p1) + os.path.getsize(p2),
                "path": "two FASTQ.gz files -> 2 inflate + 2 parse/pack threads -> pinned tiles -> af_pipeline_run -> records",
                "bound": "zlib inflate of the two files"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_pairs
    data = cpu_sample(args, sample)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, dt, nh = cpu_reference_rate(args, sample, threads, data)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = sample * len(rates) / sum(dt for _, dt in rates)
    desc = "%d of the workload's pairs per step, oracle/af_oracle.c with %d OpenMP threads" % (sample, threads)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(dt for _, dt in rates) / len(rates),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                      "data": "synthetic", "config": config_dict(args, args.gpus),
                      "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc},
                      "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step")
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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-pairs", type=int, default=10_000_000, help="bounded sample for cpu_baseline")
    ap.add_argument("--ref-pairs", type=int, default=10_000_000, help="pairs per step of the reference arm")
    ap.add_argument("--gather-cap", type=int, default=0, help="hit records per rank in the per-step all-gather "
                    "(0: sized from a probe pass, 1.25 x the largest per-rank hit count, rounded up to 4096)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the ranks' hit lists reach every rank -- p2p: the hit-compaction kernel stores them "
                         "into every GPU's log over NVLink peer memory; nccl: one all-gather per step")
    ap.add_argument("--fastq-pairs", type=int, default=500_000, help="pairs of the FASTQ.gz end-to-end sample (0: skip)")
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

    spec = workload(args)
    index = af.AnchorIndex(af.synth_anchor(spec), kp=args.kp)
    eng = af.Anchorer(index, local)
    n = args.pairs
    batch = af.synth_pairs_device(spec, rank * n, n, index.pad_byte, local)   # this rank's shard
    torch.cuda.synchronize()
    gather_cap = args.gather_cap
    if world > 1 and gather_cap <= 0:
        _, probe = eng.anchor(batch)
        t = torch.tensor([probe["hits"]], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_cap = (int(t.item()) * 5 // 4 + 4095) // 4096 * 4096

    profiling = [False]
    n_slots = max(1, args.slots)
    cand_cap = args.cand_cap or 2 * n
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
        # (removes the host launch gaps between the eight small operations of a step)
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

    out = run_steps(args.warmup)
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
    out = run_steps(args.steps)
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
    out = run_steps(args.steps)
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
                "stage_ms_per_step": {"seed_scan": stage_ms[0] / args.steps, "flag_compaction": stage_ms[1] / args.steps,
                                      "verify": stage_ms[2] / args.steps, "extend": stage_ms[3] / args.steps,
                                      "hit_compaction": stage_ms[4] / args.steps}}

    # end to end through the host-buffer C-ABI call: pinned host batch -> H2D -> kernels -> D2H hits
    e2e = None
    if not args.no_e2e:
        lay = af.layout(args.read_len, n)
        hptr = L.af_host_alloc(lay.packed_bytes)
        host_packed = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint32)), (lay.packed_bytes // 4,))
        host_packed[:] = batch.packed.cpu().numpy().view(np.uint32)
        hb = af.PackedBatch(host_packed, n, args.read_len, args.read_len)
        hits_out = np.zeros(1 << 20, dtype=af.HIT_DTYPE)
        eng.anchor_host(hb, hits_out=hits_out)                  # warm-up (allocates the slots)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            h, st = eng.anchor_host(hb, hits_out=hits_out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(lay.packed_bytes),
               "d2h_bytes_per_step": int(len(h) * 16 + 32 * ((n + (1 << 20) - 1) >> 20)), "ms_per_step": dt * 1e3,
               "path": "af_pipeline_run: pinned host tiles -> cudaMemcpyAsync -> kernels -> hit list on host",
               "host_cpus_rank0": ("%d CPUs local to the GPU (NVML affinity)" % len(cpus)) if cpus else "unbound"}
        eng.close_pipeline()
        L.af_host_free(hptr)

    # third rate: FASTQ.gz files on disk -> zlib reader -> packer -> pipeline -> records
    fastq = None
    if rank == 0 and world == 1 and args.fastq_pairs > 0 and not args.no_e2e:
        fastq = fastq_gz_rate(args, spec, index, eng)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        data = cpu_sample(args, args.cpu_pairs)
        rate, dt, nh = cpu_reference_rate(args, args.cpu_pairs, threads, data)
        n1 = min(args.cpu_pairs, 2_000_000)                      # and one core alone, on the first 2 M pairs
        rate1, dt1, _ = cpu_reference_rate(args, n1, 1, (data[0], data[1][: 2 * n1]))
        cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "first %d pairs of the workload, oracle/af_oracle.c, %d OpenMP threads, %.1f s"
                         % (args.cpu_pairs, threads, dt),
               "value_1_core": rate1, "sample_1_core": "first %d pairs, 1 thread, %.1f s" % (n1, dt1)}

    if rank == 0:
        nh = int(stats_counts[1])
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps({"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": dict(config_dict(args, world), streams=n_slots, cuda_graphs=bool(args.graphs), exchange=exchange_info), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "fastq_gz": fastq,
                          "gpu_launches": int(launches), "clocks": clocks,
                          "per_step": {"flagged_reads": int(stats_counts[0]), "seeded_reads": int(stats_counts[3]),
                                       "anchored_reads": nh,
                                       "kp": index.info.kp, "stride": index.info.stride}}) + "\n").encode())
    if exchange is not None:
        barrier()
        exchange.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
