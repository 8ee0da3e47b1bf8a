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
    """Paired FASTQ / FASTQ.gz -> packed host batches (C++: af_fastq_open / af_fastq_next)."""

    def __init__(self, path1, path2, max_read_len, pad_byte, batch_pairs, buffers=None):
        self.max_read_len, self.pad_byte, self.batch_pairs = max_read_len, pad_byte, batch_pairs
        h = ctypes.c_void_p()
        check(lib().af_fastq_open(path1.encode(), path2.encode(), ctypes.byref(h)))
        self._h = h
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

    def record(self, read_id):
        """(name, seq, qual) strings of a read of the current batch."""
        name, seq, qual = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        nl, ln = ctypes.c_int32(0), ctypes.c_int32(0)
        check(lib().af_fastq_record(self._h, read_id, ctypes.byref(name), ctypes.byref(nl), ctypes.byref(seq),
                                    ctypes.byref(qual), ctypes.byref(ln)))
        return (ctypes.string_at(name.value, nl.value).decode(), ctypes.string_at(seq.value, ln.value).decode(),
                ctypes.string_at(qual.value, ln.value).decode())

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
    return max(g, 0)


class AnchoredRead:
    __slots__ = ("hit", "gid", "name", "seq", "qual")

    def __init__(self, hit, gid, name, seq, qual):
        self.hit, self.gid, self.name, self.seq, self.qual = hit, gid, name, seq, qual


def scan_fastq_pair_multi(gene_engines, fastq1, fastq2, batch_pairs=1 << 20, max_read_len=None, batch_filter=None):
    """Stream a FASTQ pair ONCE through the GPU for several anchored genes (the
    reference re-reads both FASTQs once per gene, Anchored_Fusion.py:126,182).  gene_engines is a list
    of (AnchorIndex, Anchorer); every decoded / packed batch is handed to each engine in turn.
    Returns one (anchored, mates, stats) per gene: anchored is a list of AnchoredRead (global read
    ids), mates maps the global read id of every UNanchored mate of a half-anchored pair to its
    (name, seq, qual).  batch_filter(i) -> bool lets a rank of a multi-GPU job take only its share
    of the batches (all ranks still decode the stream).  The pad pattern of the first gene's index
    is used for all of them (the pad only influences false positives of the filter, never results)."""
    index0, eng0 = gene_engines[0]
    mrl = max_read_len or max(peek_max_read_len(fastq1), peek_max_read_len(fastq2))
    while True:
        bufs = getattr(eng0, "_host_buffers", None)
        if bufs is None or bufs.key != (mrl, batch_pairs):
            if bufs is not None:
                bufs.free()
            bufs = eng0._host_buffers = HostBuffers(mrl, batch_pairs)     # lives with the engine: reused across files
        reader = FastqPairReader(fastq1, fastq2, mrl, index0.pad_byte, batch_pairs, buffers=bufs)
        out = [([], {}, {"pairs": 0, "flagged": 0, "anchored": 0}) for _ in gene_engines]
        base = 0
        try:
            i = 0
            while True:
                batch = reader.next_batch()
                if batch is None:
                    break
                for (index, eng), (anchored, mates, stats) in zip(gene_engines, out):
                    if batch_filter is None or batch_filter(i):
                        hits, st = eng.anchor_host(batch, slot_pairs=1 << 18, n_slots=3)
                        have = set(int(r) for r in hits["read_id"])
                        for h in hits:
                            rid = int(h["read_id"])
                            name, seq, qual = reader.record(rid)
                            anchored.append(AnchoredRead(h.copy(), 2 * base + rid, name, seq, qual))
                            if (rid ^ 1) not in have:
                                mates[2 * base + (rid ^ 1)] = reader.record(rid ^ 1)
                        stats["flagged"] += st["flagged"]
                        stats["anchored"] += len(hits)
                    stats["pairs"] += batch.n_pairs
                base += batch.n_pairs
                i += 1
            reader.close()
            return out
        except AnchoredFusionError as e:
            reader.close()
            if "max_read_len" in str(e) and mrl < _lib.MAX_READ_LEN:
                mrl = _lib.MAX_READ_LEN      # a later read was longer than the peeked ones: start over
                continue
            raise


def scan_fastq_pair(index, fastq1, fastq2, device=0, batch_pairs=1 << 20, max_read_len=None, engine=None,
                    batch_filter=None):
    """One gene: see scan_fastq_pair_multi.  Returns (anchored, mates, stats)."""
    eng = engine or Anchorer(index, device)
    return scan_fastq_pair_multi([(index, eng)], fastq1, fastq2, batch_pairs, max_read_len, batch_filter)[0]


def _sorted_anchored(anchored):
    """(POS, strand, global read id) -- see records.py for why ties are pinned."""
    return sorted(anchored, key=lambda a: (int(a.hit["pos"]), int(a.hit["score_strand"]) & 1, a.gid))


def _flag(a, mate_hit):
    f = 0x1 | (0x80 if a.gid & 1 else 0x40)
    if a.hit["score_strand"] & 1:
        f |= 0x10
    if mate_hit is None:
        f |= 0x8
    elif mate_hit["score_strand"] & 1:
        f |= 0x20
    return f


def _cigar_ops(h):
    ops = []
    if h["clip_l"]:
        ops.append((int(h["clip_l"]), "S"))
    ops.append((int(h["m_len"]), "M"))
    if h["clip_r"]:
        ops.append((int(h["clip_r"]), "S"))
    return ops


def write_stage_outputs(prefix, gene, anchor_len, anchored, mates):
    """Write the files Anchored_Fusion.py:181-194 would have left behind:
       <prefix>_anchored_reads.bam   mapped primary reads, coordinate order   (samtools view -F 772 | sort)
       <prefix>_tmp_1.fastq / _tmp_2.fastq   half-anchored pairs: anchored mate / its unanchored mate,
                                     original orientation, names suffixed /1 /2   (samtools fastq)
       <prefix>_realign_reads.bam    the anchored reads plus those unanchored mates -- everything the
                                     reference's three samtools filters can select; pairs with no
                                     anchored mate are not written (no later stage reads them)
       <prefix>_anchored_reads.raw.sam   the same records as 11-column SAM text, for samtools-less hosts
    """
    srt = _sorted_anchored(anchored)
    by_gid = {a.gid: a for a in srt}
    paths = {"anchored_bam": prefix + "_anchored_reads.bam", "realign_bam": prefix + "_realign_reads.bam",
             "tmp1": prefix + "_tmp_1.fastq", "tmp2": prefix + "_tmp_2.fastq",
             "raw_sam": prefix + "_anchored_reads.raw.sam"}
    with BamWriter(paths["anchored_bam"], gene, anchor_len) as ab, BamWriter(paths["realign_bam"], gene, anchor_len) as rb, \
            open(paths["tmp1"], "w") as t1, open(paths["tmp2"], "w") as t2, open(paths["raw_sam"], "w") as sam:
        for a in srt:
            h = a.hit
            rev = bool(h["score_strand"] & 1)
            seq = revcomp(a.seq) if rev else a.seq
            qual = a.qual[::-1] if rev else a.qual
            mate = by_gid.get(a.gid ^ 1)
            flag = _flag(a, mate.hit if mate else None)
            pnext = int(mate.hit["pos"]) if mate else int(h["pos"])
            for w in (ab, rb):
                w.write(a.name, flag, int(h["pos"]), 60, _cigar_ops(h), seq, qual, next_pos=pnext)
            sam.write("\t".join([a.name, str(flag), gene, str(int(h["pos"])), "60", cigar_string(h), "=", str(pnext),
                                 "0", seq, qual]) + "\n")
            if mate is None:
                mname, mseq, mqual = mates[a.gid ^ 1]
                mflag = 0x1 | 0x4 | (0x20 if rev else 0) | (0x80 if (a.gid ^ 1) & 1 else 0x40)
                rb.write(mname, mflag, int(h["pos"]), 0, [], mseq, mqual, next_pos=int(h["pos"]), mapped=False)
                t1.write("@%s/%d\n%s\n+\n%s\n" % (a.name, (a.gid & 1) + 1, a.seq, a.qual))
                t2.write("@%s/%d\n%s\n+\n%s\n" % (mname, ((a.gid ^ 1) & 1) + 1, mseq, mqual))
    return paths


def hits_array(anchored):
    """HIT_DTYPE array of a list of AnchoredRead with GLOBAL read ids, in (POS, strand, id) order."""
    hits = np.zeros(len(anchored), dtype=HIT_DTYPE)
    for i, a in enumerate(anchored):
        hits[i] = a.hit
        hits[i]["read_id"] = a.gid & 0xFFFFFFFF
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
        if info.n_overflow * 200 > info.n_buckets:
            # the shared-memory filter holds ~3 k'-mers per bucket; past ~12 kb of anchor more and more buckets
            # overflow into "always hit" and the exact verify stage has to sort out the difference
            print("[anchoring] %s: anchor of %d bp fills the seed filter (%d of %d buckets overflow); results are "
                  "unaffected, the scan flags more reads than usual" % (self.gene, len(self.seq), info.n_overflow, info.n_buckets),
                  file=sys.stderr)
        self.engine = Anchorer(self.index, resolve_device(gpu_number))


def finish_stage(ga, out_prefix, anchored, mates, stats):
    """Write one gene's stage files from a scan result; returns the stats dict of anchor_stage."""
    stats = dict(stats)
    stats.update(write_stage_outputs(out_prefix, ga.gene, len(ga.seq), anchored, mates))
    stats["half_anchored_pairs"] = len(mates)
    stats["hits"] = hits_array(anchored)
    return stats


def anchor_stage_multi(gene_anchorers, fastq1, fastq2, out_prefixes, batch_pairs=1 << 20):
    """The anchoring stage for several genes with ONE pass over the FASTQ pair."""
    results = scan_fastq_pair_multi([(ga.index, ga.engine) for ga in gene_anchorers], fastq1, fastq2, batch_pairs)
    return [finish_stage(ga, prefix, *res) for ga, prefix, res in zip(gene_anchorers, out_prefixes, results)]


def anchor_stage(file_anchored_seq, fastq1, fastq2, out_prefix, thread="1", gpu_number="-1", gene_name=None,
                 batch_pairs=1 << 20, kp=0, gene_anchorer=None):
    """Drop-in for the anchoring stage.  file_anchored_seq is <work>_anchored_gene_sequence.fa;
    `thread` is accepted for signature compatibility (two inflate + two parse/pack threads and the GPU do
    the work).  Returns a stats dict; see write_stage_outputs for the files."""
    ga = gene_anchorer or GeneAnchorer(file_anchored_seq, gpu_number, gene_name, kp)
    anchored, mates, stats = scan_fastq_pair(ga.index, fastq1, fastq2, batch_pairs=batch_pairs, engine=ga.engine)
    return finish_stage(ga, out_prefix, anchored, mates, stats)
