"""The anchoring stage as the driver scripts see it: per-gene anchor FASTA + two FASTQ(.gz)
files in, the stage's output files out.  Replaces Anchored_Fusion.py:167-194 (and
Anchored_Fusion_singlecell.py:185-231):

    bwa index <anchor.fa>
    bwa mem -M -t T <anchor.fa> fq1 fq2 | samtools view -bSu - | samtools sort - -o <w>_realign_reads.bam
    samtools view -u -f 8 -F 260 <w>_realign_reads.bam | samtools fastq - -o <w>_tmp_1.fastq
    samtools view -u -f 4 -F 264 <w>_realign_reads.bam | samtools fastq - -o <w>_tmp_2.fastq
    samtools view -u -F 772 -h <w>_realign_reads.bam | samtools sort - -o <w>_anchored_reads.bam

The FASTQ pair is streamed once through the C++ reader (zlib, 2-bit packing into pinned
buffers) and the GPU pipeline; only the anchored reads and the mates of half-anchored pairs are
ever materialised as text.
"""
import ctypes
import os
import sys

import numpy as np

from . import _lib
from ._lib import HIT_DTYPE, AnchoredFusionError, check, lib
from .anchoring import Anchorer, AnchorIndex, PackedBatch, layout
from .bam import BamWriter
from .records import cigar_string, revcomp, sort_hits


def read_single_fasta(path):
    """(name, sequence) of a one-record FASTA such as <work>_anchored_gene_sequence.fa."""
    name, parts = None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    break
                name = line[1:].strip()
            else:
                parts.append(line.strip())
    if name is None:
        raise AnchoredFusionError("%s holds no FASTA record" % path)
    return name, "".join(parts)


class FastqPairReader:
    """Paired FASTQ / FASTQ.gz / BGZF -> packed host batches (C++: af_fastq_open_* / af_fastq_next).
    path1 / path2 may be lists of equal length: the file pairs are then read back to back as one stream of
    pairs (single-cell layout: one pair of files per cell) and `file_starts()` tells where each begins.
    threads = worker threads of the reader (`--thread`; 0 = one per host core)."""

    def __init__(self, path1, path2, max_read_len, pad_byte, batch_pairs, buffers=None, threads=0):
        self.max_read_len, self.pad_byte, self.batch_pairs = max_read_len, pad_byte, batch_pairs
        h = ctypes.c_void_p()
        if isinstance(path1, (list, tuple)):
            if len(path1) != len(path2) or not path1:
                raise AnchoredFusionError("the two mates' file lists differ in length (or are empty)")
            arr = ctypes.c_char_p * len(path1)
            a1, a2 = arr(*[p.encode() for p in path1]), arr(*[p.encode() for p in path2])
            check(lib().af_fastq_open_multi(a1, a2, len(path1), int(threads), ctypes.byref(h)))
            self.n_files = len(path1)
        else:
            check(lib().af_fastq_open_threads(path1.encode(), path2.encode(), int(threads), ctypes.byref(h)))
            self.n_files = 1
        self._h = h
        self.threads = lib().af_fastq_threads(h)
        self._bufs = buffers if buffers is not None else HostBuffers(max_read_len, batch_pairs)
        self._own = buffers is None
        self.packed, self.lens, self.nids, self.nmask = self._bufs.packed, self._bufs.lens, self._bufs.nids, self._bufs.nmask

    def next_batch(self):
        """PackedBatch (host) or None at EOF; record text stays valid until the next call."""
        nn, ulen, n = ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_int64(0)
        check(lib().af_fastq_next(self._h, self.batch_pairs, self.max_read_len, self.pad_byte, self.packed.ctypes.data,
                                  self.lens.ctypes.data, self.nids.ctypes.data, self.nmask.ctypes.data, len(self.nids),
                                  ctypes.byref(nn), ctypes.byref(ulen), ctypes.byref(n)))
        if n.value == 0:
            return None
        k = nn.value
        return PackedBatch(self.packed, n.value, self.max_read_len, ulen.value, self.lens[: 2 * n.value],
                           self.nids[:k] if k else None, self.nmask[:k] if k else None)

    def skip_batch(self):
        """Step over the next batch without packing it (another rank's share); pairs skipped, 0 at EOF."""
        n = ctypes.c_int64(0)
        check(lib().af_fastq_skip(self._h, self.batch_pairs, ctypes.byref(n)))
        return n.value

    @property
    def first_pair(self):
        """Index, over the whole run, of the current batch's first pair."""
        return lib().af_fastq_batch_first_pair(self._h)

    def file_starts(self):
        """int64 array: index (over the whole run) of the first pair of every input file started so far."""
        out = np.zeros(self.n_files, dtype=np.int64)
        k = lib().af_fastq_file_starts(self._h, out.ctypes.data, self.n_files)
        return out[:k]

    def record(self, read_id):
        """(name, seq, qual) strings of a read of the current batch."""
        name, seq, qual = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        nl, ln = ctypes.c_int32(0), ctypes.c_int32(0)
        check(lib().af_fastq_record(self._h, read_id, ctypes.byref(name), ctypes.byref(nl), ctypes.byref(seq),
                                    ctypes.byref(qual), ctypes.byref(ln)))
        return (ctypes.string_at(name.value, nl.value).decode(), ctypes.string_at(seq.value, ln.value).decode(),
                ctypes.string_at(qual.value, ln.value).decode())

    def records(self, read_ids):
        """[(name, seq, qual)] of many reads of the current batch: one library call, one text buffer."""
        ids = np.ascontiguousarray(read_ids, dtype=np.int64)
        n = len(ids)
        if n == 0:
            return []
        offs = np.zeros(4 * n, dtype=np.int64)
        used = ctypes.c_int64(0)
        cap = n * (3 * self.max_read_len + 64)
        buf = ctypes.create_string_buffer(cap)
        rc = lib().af_fastq_records(self._h, ids.ctypes.data, n, buf, cap, offs.ctypes.data, ctypes.byref(used))
        if rc != 0 and used.value > cap:                      # very long read names: size the buffer exactly
            cap = used.value
            buf = ctypes.create_string_buffer(cap)
            rc = lib().af_fastq_records(self._h, ids.ctypes.data, n, buf, cap, offs.ctypes.data, ctypes.byref(used))
        check(rc)
        text = buf.raw[: used.value].decode("latin-1")
        o = offs.tolist()
        return [(text[o[4 * i]: o[4 * i + 1]], text[o[4 * i + 1]: o[4 * i + 2]], text[o[4 * i + 2]: o[4 * i + 3]]) for i in range(n)]

    def close(self):
        if self._h:
            lib().af_fastq_close(self._h)
            self._h = None
        if self._own and self._bufs is not None:
            self._bufs.free()
        self._bufs = None

    def __del__(self):
        try:
            if _lib._lib is not None:
                self.close()
        except Exception:
            pass


class HostBuffers:
    """Pinned staging for one packed batch (+ lens and the N side list); reusable across files."""

    def __init__(self, max_read_len, batch_pairs):
        self.key = (max_read_len, batch_pairs)
        lay = layout(max_read_len, batch_pairs)
        self._pinned = lib().af_host_alloc(lay.packed_bytes)     # pageable if pinning fails
        if self._pinned:
            self.packed = np.ctypeslib.as_array(ctypes.cast(self._pinned, ctypes.POINTER(ctypes.c_uint32)),
                                                (lay.packed_bytes // 4,))
        else:
            self.packed = np.zeros(lay.packed_bytes // 4, dtype=np.uint32)
        self.lens = np.zeros(2 * batch_pairs, dtype=np.uint16)
        self.nids = np.zeros(2 * batch_pairs, dtype=np.uint32)
        self.nmask = np.zeros((2 * batch_pairs, _lib.NMASK_WORDS), dtype=np.uint32)

    def free(self):
        if self._pinned and _lib._lib is not None:
            lib().af_host_free(self._pinned)
        self._pinned = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def peek_max_read_len(path, n_records=2000):
    """Longest read among the first records, rounded up to a packed word (16 bases)."""
    longest = ctypes.c_int32(0)
    check(lib().af_fastq_peek(path.encode(), n_records, ctypes.byref(longest)))
    return min(_lib.MAX_READ_LEN, (max(longest.value, 1) + 15) // 16 * 16)


def resolve_device(gpu_number):
    """--gpu_number is a string in the reference ('-1' = CPU for its filter model, Model.py:14-19).
    The anchoring path has no CPU fallback, so '-1' means 'the first visible GPU' here -- or, in a
    process started by torchrun, the GPU of its LOCAL_RANK."""
    import os
    try:
        g = int(gpu_number)
    except (TypeError, ValueError):
        g = -1
    if g < 0:
        g = int(os.environ.get("LOCAL_RANK", "0"))
        try:
            import torch
            n = torch.cuda.device_count()
            if n > 0:
                g %= n           # more ranks than GPUs (tests on a one-GPU box): ranks share devices
        except Exception:
            pass
    return max(g, 0)


def host_threads(thread):
    """`--thread` is a string in the reference (Anchored_Fusion.py:29, forwarded to `bwa mem -t`).  Here it
    is the reader's worker count; anything that is not a positive integer means one worker per host core."""
    try:
        t = int(thread)
    except (TypeError, ValueError):
        t = 0
    if t > 0:
        return t
    # one process per GPU (torchrun): the ranks of a node share its cores instead of each starting one worker per core
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or "1")
    if local_world > 1:
        return max(2, (os.cpu_count() or 1) // local_world)
    return 0


class AnchoredRead:
    """One anchored read: its record fields, its read id over the whole run (2 * pair + mate) and its text."""
    __slots__ = ("gid", "pos", "clip_l", "m_len", "clip_r", "score_strand", "name", "seq", "qual")

    def __init__(self, gid, pos, clip_l, m_len, clip_r, score_strand, name, seq, qual):
        self.gid, self.pos, self.clip_l, self.m_len, self.clip_r, self.score_strand = gid, pos, clip_l, m_len, clip_r, score_strand
        self.name, self.seq, self.qual = name, seq, qual

    @property
    def rev(self):
        return self.score_strand & 1

    def as_tuple(self):
        return (self.gid, self.pos, self.clip_l, self.m_len, self.clip_r, self.score_strand, self.name, self.seq, self.qual)


def _host_group():
    """(rank, world, gloo group or None).  Under torchrun the ranks of a job exchange their small host-side
    results (records + read text) through a gloo group; the GPUs need no collective for this path."""
    import os
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, None
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo")
        return rank, world, None                 # the default group is gloo
    if dist.get_backend() == "gloo":
        return rank, world, None
    g = getattr(_host_group, "_g", None)
    if g is None:
        g = _host_group._g = dist.new_group(backend="gloo")
    return rank, world, g


def anchor_host_multi(engines, batch, slot_pairs=1 << 18, n_slots=3):
    """One pass of a host batch through the GPU for several anchor engines (same device): each chunk is
    copied to the GPU once and scanned for every gene while resident (af_pipeline_run_multi).
    Returns [(hits, stats)] per engine."""
    eng0 = engines[0]
    if len(engines) == 1:
        return [eng0.anchor_host(batch, slot_pairs=slot_pairs, n_slots=n_slots)]
    pipe = eng0.pipeline(batch.max_read_len, slot_pairs, n_slots)
    g = len(engines)
    cap = 2 * max(batch.n_pairs, 1)
    outs = [np.zeros(cap, dtype=HIT_DTYPE) for _ in engines]
    idx = (ctypes.c_void_p * g)(*[e.dindex._h for e in engines])
    hp = (ctypes.c_void_p * g)(*[o.ctypes.data for o in outs])
    caps = (ctypes.c_int64 * g)(*([cap] * g))
    nh, nf = (ctypes.c_int64 * g)(), (ctypes.c_int64 * g)()
    cb = batch.c_struct()
    check(lib().af_pipeline_run_multi(pipe, g, idx, ctypes.byref(cb), hp, caps, nh, nf))
    return [(outs[k][: nh[k]], {"flagged": nf[k], "hits": nh[k]}) for k in range(g)]


def scan_fastq_pair_multi(gene_engines, fastq1, fastq2, batch_pairs=1 << 20, max_read_len=None, batch_filter=None,
                          threads=0, rank=0, world=1):
    """Stream a FASTQ pair (or a list of pairs: cells) ONCE through the GPU for several anchored genes
    (the reference re-reads both FASTQs once per gene, Anchored_Fusion.py:126,182).
    gene_engines is a list of (AnchorIndex, Anchorer) on one device; every decoded / packed batch is copied
    to the GPU once and scanned for each gene while resident.
    Returns (results, info): results = one (anchored, mates, stats) per gene -- anchored is a list of
    AnchoredRead (read ids over the whole run), mates maps the id of every UNanchored mate of a
    half-anchored pair to its (name, seq, qual); info = {"file_starts": first pair of every input file,
    "pairs": total}.  With world > 1 (or batch_filter) this process anchors only its share of the batches
    (batch i belongs to rank i % world; all ranks decode the stream, nobody packs or copies what it skips).
    The pad pattern of the first gene's index is used for all of them (the pad only influences false
    positives of the filter, never results)."""
    index0, eng0 = gene_engines[0]
    engines = [e for _, e in gene_engines]
    multi = isinstance(fastq1, (list, tuple))
    first1, first2 = (fastq1[0], fastq2[0]) if multi else (fastq1, fastq2)
    if max_read_len:
        mrl = max_read_len
    elif multi:
        mrl = max(peek_max_read_len(p) for p in list(fastq1[:8]) + list(fastq2[:8]))
    else:
        mrl = max(peek_max_read_len(first1), peek_max_read_len(first2))
    if batch_filter is None and world > 1:
        def batch_filter(i):
            return i % world == rank
    while True:
        bufs = getattr(eng0, "_host_buffers", None)
        if bufs is None or bufs.key != (mrl, batch_pairs):
            if bufs is not None:
                bufs.free()
            bufs = eng0._host_buffers = HostBuffers(mrl, batch_pairs)     # lives with the engine: reused across files
        reader = FastqPairReader(fastq1, fastq2, mrl, index0.pad_byte, batch_pairs, buffers=bufs, threads=threads)
        out = [([], {}, {"pairs": 0, "flagged": 0, "anchored": 0}) for _ in gene_engines]
        total = 0
        try:
            i = 0
            while True:
                mine = batch_filter is None or batch_filter(i)
                if mine:
                    batch = reader.next_batch()
                    if batch is None:
                        break
                    n = batch.n_pairs
                    base = reader.first_pair
                    for (anchored, mates, stats), (hits, st) in zip(out, anchor_host_multi(engines, batch)):
                        if len(hits):
                            rids = hits["read_id"].astype(np.int64)
                            lone = rids[~np.isin(rids ^ 1, rids)] ^ 1          # unanchored mates of half-anchored pairs
                            text = reader.records(np.concatenate([rids, lone]))
                            k = len(rids)
                            g0 = 2 * base
                            for h, (name, seq, qual) in zip(hits.tolist(), text[:k]):
                                anchored.append(AnchoredRead(g0 + h[0], h[1], h[2], h[3], h[4], h[5], name, seq, qual))
                            for rid, t in zip(lone.tolist(), text[k:]):
                                mates[g0 + rid] = t
                        stats["flagged"] += st["flagged"]
                        stats["anchored"] += len(hits)
                else:
                    n = reader.skip_batch()
                    if n == 0:
                        break
                total += n
                i += 1
            info = {"file_starts": reader.file_starts(), "pairs": total, "threads": reader.threads}
            for _, _, stats in out:
                stats["pairs"] = total
            reader.close()
            return out, info
        except AnchoredFusionError as e:
            reader.close()
            if "max_read_len" in str(e) and mrl < _lib.MAX_READ_LEN:
                mrl = 256 if mrl < 256 else _lib.MAX_READ_LEN      # a later read was longer than the peeked ones: start over
                continue
            if "max_read_len" in str(e):
                raise AnchoredFusionError(
                    "%s -- this path packs reads of at most %d bases (AF_MAX_READ_LEN); longer reads (merged pairs, "
                    "long-read platforms) are not supported: trim them or run that sample through the reference's bwa stage"
                    % (e, _lib.MAX_READ_LEN))
            raise


def gather_results(results, group=None, rank=0, world=1):
    """Multi-GPU runs: every rank hands its share of the records (with their text) to rank 0 -- a few
    hundred bytes per anchored read, host to host.  Rank 0 gets the merged per-gene results, the others None."""
    if world <= 1:
        return results
    import torch.distributed as dist
    payload = [([a.as_tuple() for a in anchored], mates, stats) for anchored, mates, stats in results]
    box = [None] * world if rank == 0 else None
    dist.gather_object(payload, box, dst=0, group=group)
    if rank != 0:
        return None
    merged = []
    for g in range(len(results)):
        anchored, mates = [], {}
        stats = dict(box[0][g][2])
        stats["flagged"] = stats["anchored"] = 0
        for r in range(world):
            a, m, st = box[r][g]
            anchored += [AnchoredRead(*t) for t in a]
            mates.update(m)
            stats["flagged"] += st["flagged"]
            stats["anchored"] += st["anchored"]
        merged.append((anchored, mates, stats))
    return merged


def scan_fastq_pair(index, fastq1, fastq2, device=0, batch_pairs=1 << 20, max_read_len=None, engine=None,
                    batch_filter=None, threads=0):
    """One gene: see scan_fastq_pair_multi.  Returns (anchored, mates, stats)."""
    eng = engine or Anchorer(index, device)
    res, _ = scan_fastq_pair_multi([(index, eng)], fastq1, fastq2, batch_pairs, max_read_len, batch_filter, threads)
    return res[0]


def _sorted_anchored(anchored):
    """(POS, strand, read id over the run) -- see records.py for why ties are pinned."""
    return sorted(anchored, key=lambda a: (a.pos, a.score_strand & 1, a.gid))


def _flag(a, mate):
    f = 0x1 | (0x80 if a.gid & 1 else 0x40)
    if a.score_strand & 1:
        f |= 0x10
    if mate is None:
        f |= 0x8
    elif mate.score_strand & 1:
        f |= 0x20
    return f


def _cigar_ops(a):
    ops = []
    if a.clip_l:
        ops.append((a.clip_l, "S"))
    ops.append((a.m_len, "M"))
    if a.clip_r:
        ops.append((a.clip_r, "S"))
    return ops


def _cigar_text(a):
    return ("%dS" % a.clip_l if a.clip_l else "") + "%dM" % a.m_len + ("%dS" % a.clip_r if a.clip_r else "")


def write_stage_outputs(prefix, gene, anchor_len, anchored, mates):
    """Write the files Anchored_Fusion.py:181-194 would have left behind:
       <prefix>_anchored_reads.bam   mapped primary reads, coordinate order   (samtools view -F 772 | sort)
       <prefix>_tmp_1.fastq / _tmp_2.fastq   half-anchored pairs: anchored mate / its unanchored mate,
                                     original orientation, names suffixed /1 /2   (samtools fastq)
       <prefix>_realign_reads.bam    the anchored reads plus those unanchored mates -- everything the
                                     reference's three samtools filters can select; pairs with no
                                     anchored mate are not written (no later stage reads them)
       <prefix>_anchored_reads.raw.sam   the same records as 11-column SAM text, for samtools-less hosts
    Every file is written under a temporary name and renamed when complete, the two BAMs last: the
    reference's existence guards (Anchored_Fusion.py:181,193) then never see a half-written stage.
    """
    import os
    srt = _sorted_anchored(anchored)
    by_gid = {a.gid: a for a in srt}
    paths = {"anchored_bam": prefix + "_anchored_reads.bam", "realign_bam": prefix + "_realign_reads.bam",
             "tmp1": prefix + "_tmp_1.fastq", "tmp2": prefix + "_tmp_2.fastq",
             "raw_sam": prefix + "_anchored_reads.raw.sam"}
    tmp = {k: v + ".partial" for k, v in paths.items()}
    with BamWriter(tmp["anchored_bam"], gene, anchor_len) as ab, BamWriter(tmp["realign_bam"], gene, anchor_len) as rb, \
            open(tmp["tmp1"], "w") as t1, open(tmp["tmp2"], "w") as t2, open(tmp["raw_sam"], "w") as sam:
        for a in srt:
            rev = bool(a.score_strand & 1)
            seq = revcomp(a.seq) if rev else a.seq
            qual = a.qual[::-1] if rev else a.qual
            mate = by_gid.get(a.gid ^ 1)
            flag = _flag(a, mate)
            pnext = mate.pos if mate else a.pos
            ops = _cigar_ops(a)
            for w in (ab, rb):
                w.write(a.name, flag, a.pos, 60, ops, seq, qual, next_pos=pnext)
            sam.write("\t".join([a.name, str(flag), gene, str(a.pos), "60", _cigar_text(a), "=", str(pnext),
                                 "0", seq, qual]) + "\n")
            if mate is None:
                mname, mseq, mqual = mates[a.gid ^ 1]
                mflag = 0x1 | 0x4 | (0x20 if rev else 0) | (0x80 if (a.gid ^ 1) & 1 else 0x40)
                rb.write(mname, mflag, a.pos, 0, [], mseq, mqual, next_pos=a.pos, mapped=False)
                t1.write("@%s/%d\n%s\n+\n%s\n" % (a.name, (a.gid & 1) + 1, a.seq, a.qual))
                t2.write("@%s/%d\n%s\n+\n%s\n" % (mname, ((a.gid ^ 1) & 1) + 1, mseq, mqual))
    for k in ("tmp1", "tmp2", "raw_sam", "realign_bam", "anchored_bam"):
        os.replace(tmp[k], paths[k])
    return paths


def hits_array(anchored, id_base=0):
    """HIT_DTYPE array of a list of AnchoredRead, read ids relative to id_base, in (POS, strand, id) order."""
    hits = np.zeros(len(anchored), dtype=HIT_DTYPE)
    for i, a in enumerate(anchored):
        hits[i] = ((a.gid - id_base) & 0xFFFFFFFF, a.pos, a.clip_l, a.m_len, a.clip_r, a.score_strand)
    return sort_hits(hits)


class GeneAnchorer:
    """Index + GPU engine of one anchored gene, reusable over many FASTQ pairs (single-cell runs
    call the stage once per cell: the index upload and the pinned / device staging buffers are
    paid once per gene, not once per cell)."""

    def __init__(self, file_anchored_seq, gpu_number="-1", gene_name=None, kp=0):
        name, self.seq = read_single_fasta(file_anchored_seq)
        self.gene = gene_name or name.split()[0]
        self.index = AnchorIndex(self.seq, kp=kp)
        info = self.index.info
        if self.index.bloom:
            # the 3-slot fingerprint buckets hold anchors up to ~12 kb; beyond that the same shared memory holds a
            # blocked Bloom filter: exact results, more reads than usual reach the verify stage
            print("[anchoring] %s: anchor of %d bp is beyond the fingerprint filter (%d of %d buckets would overflow); "
                  "using the Bloom filter, results are unaffected" % (self.gene, len(self.seq), info.n_overflow, info.n_buckets),
                  file=sys.stderr)
        self.engine = Anchorer(self.index, resolve_device(gpu_number))


def finish_stage(ga, out_prefix, anchored, mates, stats, id_base=0):
    """Write one gene's stage files from a scan result; returns the stats dict of anchor_stage."""
    stats = dict(stats)
    stats.update(write_stage_outputs(out_prefix, ga.gene, len(ga.seq), anchored, mates))
    stats["half_anchored_pairs"] = len(mates)
    stats["hits"] = hits_array(anchored, id_base)
    return stats


def default_batch_pairs():
    """Pairs per packed host batch (AF_BATCH_PAIRS overrides it: tests use small batches to exercise the
    multi-batch and multi-rank paths on small samples)."""
    import os
    return int(os.environ.get("AF_BATCH_PAIRS", str(1 << 20)))


def anchor_stage_multi(gene_anchorers, fastq1, fastq2, out_prefixes, batch_pairs=None, thread="0"):
    """The anchoring stage for several genes with ONE pass over the FASTQ pair.  Under torchrun
    (WORLD_SIZE > 1) the batches are dealt to the ranks, each rank anchors its share on its own GPU and
    rank 0 -- which gets every rank's records -- writes the one set of files; the other ranks return None."""
    rank, world, group = _host_group()
    results, _ = scan_fastq_pair_multi([(ga.index, ga.engine) for ga in gene_anchorers], fastq1, fastq2,
                                       batch_pairs or default_batch_pairs(), threads=host_threads(thread), rank=rank, world=world)
    results = gather_results(results, group, rank, world)
    if results is None:
        return None
    return [finish_stage(ga, prefix, *res) for ga, prefix, res in zip(gene_anchorers, out_prefixes, results)]


def anchor_stage(file_anchored_seq, fastq1, fastq2, out_prefix, thread="1", gpu_number="-1", gene_name=None,
                 batch_pairs=None, kp=0, gene_anchorer=None):
    """Drop-in for the anchoring stage.  file_anchored_seq is <work>_anchored_gene_sequence.fa; `thread` is the
    reader's worker count, as it is bwa's in the reference.  Returns a stats dict (None on ranks other than 0
    of a torchrun job); see write_stage_outputs for the files."""
    ga = gene_anchorer or GeneAnchorer(file_anchored_seq, gpu_number, gene_name, kp)
    res = anchor_stage_multi([ga], fastq1, fastq2, [out_prefix], batch_pairs, thread)
    return res[0] if res is not None else None


def empty_stage_files(prefix, gene, anchor_len, cache={}):
    """The stage's files for a sample without a single anchored read (most cells of a single-cell run):
    the two BAMs are the same few bytes every time, so they are compressed once per gene."""
    import os
    key = (gene, anchor_len)
    if key not in cache:
        import tempfile
        fd, tmp = tempfile.mkstemp(suffix=".bam")
        os.close(fd)
        BamWriter(tmp, gene, anchor_len).close()
        with open(tmp, "rb") as fh:
            cache[key] = fh.read()
        os.remove(tmp)
    blob = cache[key]
    for suffix in ("_tmp_1.fastq", "_tmp_2.fastq", "_anchored_reads.raw.sam"):
        open(prefix + suffix, "w").close()
    for suffix in ("_realign_reads.bam", "_anchored_reads.bam"):
        with open(prefix + suffix + ".partial", "wb") as fh:
            fh.write(blob)
        os.replace(prefix + suffix + ".partial", prefix + suffix)
    return {"anchored_bam": prefix + "_anchored_reads.bam", "realign_bam": prefix + "_realign_reads.bam",
            "tmp1": prefix + "_tmp_1.fastq", "tmp2": prefix + "_tmp_2.fastq", "raw_sam": prefix + "_anchored_reads.raw.sam"}


def anchor_cells(gene_anchorers, cell_files, prefix_of, batch_pairs=None, thread="0", on_cell=None):
    """Single-cell layout (Anchored_Fusion_singlecell.py:86-113,205-231: one FASTQ pair per cell, the stage run
    once per gene and cell).  All cells of `cell_files` ([(cell, fastq1, fastq2)]) are decoded concurrently and
    packed back to back into shared batches -- the GPU sees 1 M-pair launches, not one launch per 5 k-pair
    cell -- and the hit lists are split back per cell by pair range on the host.
    prefix_of(gene, cell) -> output prefix.  on_cell(cell, [stats per gene]) is called as each cell's files are
    written.  Returns {"cells", "pairs", "anchored", "seconds_scan", "seconds_write"}."""
    import time
    t0 = time.time()
    f1 = [c[1] for c in cell_files]
    f2 = [c[2] for c in cell_files]
    results, info = scan_fastq_pair_multi([(ga.index, ga.engine) for ga in gene_anchorers], f1, f2,
                                          batch_pairs or default_batch_pairs(), threads=host_threads(thread))
    t1 = time.time()
    starts = info["file_starts"]
    assert len(starts) == len(cell_files), "reader reported %d of %d cells" % (len(starts), len(cell_files))
    bounds = np.concatenate([starts, [info["pairs"]]]).astype(np.int64) * 2          # read-id bounds per cell
    per_gene = []
    for anchored, mates, _ in results:
        anchored.sort(key=lambda a: a.gid)
        gids = np.fromiter((a.gid for a in anchored), dtype=np.int64, count=len(anchored))
        cuts = np.searchsorted(gids, bounds)
        mg = np.sort(np.fromiter(mates.keys(), dtype=np.int64, count=len(mates)))
        mcuts = np.searchsorted(mg, bounds)
        per_gene.append((anchored, cuts, mates, mg, mcuts))
    n_anchored = 0
    for ci, (cell, _, _) in enumerate(cell_files):
        cell_stats = []
        for ga, (anchored, cuts, mates, mg, mcuts) in zip(gene_anchorers, per_gene):
            prefix = prefix_of(ga.gene, cell)
            part = anchored[cuts[ci]: cuts[ci + 1]]
            pairs = int((bounds[ci + 1] - bounds[ci]) // 2)
            if not part:
                st = {"pairs": pairs, "flagged": 0, "anchored": 0, "half_anchored_pairs": 0,
                      "hits": np.zeros(0, dtype=HIT_DTYPE)}
                st.update(empty_stage_files(prefix, ga.gene, len(ga.seq)))
            else:
                cm = {int(g): mates[int(g)] for g in mg[mcuts[ci]: mcuts[ci + 1]]}
                st = finish_stage(ga, prefix, part, cm, {"pairs": pairs, "flagged": 0, "anchored": len(part)},
                                  id_base=int(bounds[ci]))
            n_anchored += len(part)
            cell_stats.append(st)
        if on_cell:
            on_cell(cell, cell_stats)
    return {"cells": len(cell_files), "pairs": info["pairs"], "anchored": n_anchored, "seconds_scan": t1 - t0,
            "seconds_write": time.time() - t1, "threads": info["threads"]}
