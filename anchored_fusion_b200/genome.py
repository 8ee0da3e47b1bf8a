"""Genome pass of the contiguity filter on the GPU.

Replaces the `bwa mem -M -t T <genome> <w>_del_tmp.fa` call inside del_too_many_reads
(functions.py:716): the 2-op anchored reads are aligned to the whole genome under the anchoring
spec of the main path, with no genome index -- the genome sits 2 bit/base in HBM and is streamed past
a shared-memory filter built from the reads (csrc/af_genome.cu).  `Genome.sam_lines` prints the
records the way `bwa mem` prints them (one primary line per read; FLAG 0 / 16 / 4, CIGAR
`<clip>S<m>M<clip>S`, SEQ reverse-complemented on strand 1), which is what
functions.contiguity_filter reads.
"""
import bisect
import ctypes

import numpy as np

from . import _lib
from ._lib import GENOME_HIT_DTYPE, GENOME_SEP, GenomeStats, Params, check, lib

_RC = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def _revcomp(seq):
    """Reverse complement as bwa prints SEQ on FLAG 16 (letters outside ACGTN stay as they are)."""
    return seq.encode().translate(_RC)[::-1].decode()


class Genome:
    """A genome resident in one GPU's HBM (2 bit/base + N bitmap)."""

    def __init__(self, handle):
        self._h = handle
        L = lib()
        self.length = L.af_genome_length(handle)
        self.contigs = []                      # (name, start in the concatenation, length)
        for i in range(L.af_genome_n_contigs(handle)):
            name, start, ln = ctypes.c_char_p(), ctypes.c_int64(), ctypes.c_int64()
            check(L.af_genome_contig(handle, i, ctypes.byref(name), ctypes.byref(start), ctypes.byref(ln)))
            self.contigs.append((name.value.decode(), start.value, ln.value))
        self._starts = [c[1] for c in self.contigs]
        self.last_stats = None

    @classmethod
    def from_fasta(cls, path, device=0):
        h = ctypes.c_void_p()
        check(lib().af_genome_from_fasta(str(path).encode(), device, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_contigs(cls, contigs, device=0):
        """contigs: [(name, sequence)]"""
        n = len(contigs)
        names = (ctypes.c_char_p * n)(*[c[0].encode() for c in contigs])
        blobs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
        seqs = (ctypes.c_char_p * n)(*blobs)
        lens = (ctypes.c_int64 * n)(*[len(b) for b in blobs])
        h = ctypes.c_void_p()
        check(lib().af_genome_from_contigs(names, seqs, lens, n, device, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def synthetic(cls, seed, length, device=0):
        """One contig whose base x is the seeded generator's reference base (measurement input)."""
        h = ctypes.c_void_p()
        check(lib().af_genome_synth(seed, length, device, ctypes.byref(h)))
        return cls(h)

    def close(self):
        if self._h is not None:
            lib().af_genome_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align(self, seqs, params=None, reads_per_pass=0):
        """seqs: list of str / bytes (<= 256 bases each) -> records (GENOME_HIT_DTYPE), ordered by read_id;
        pos is 1-based in the concatenated genome (`locate` maps it to a contig)."""
        blobs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        offs = np.zeros(len(blobs) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in blobs], out=offs[1:])
        text = b"".join(blobs)
        hits = np.zeros(max(len(blobs), 1), dtype=GENOME_HIT_DTYPE)
        n, st = ctypes.c_int64(0), GenomeStats()
        p = params if params is not None else None
        check(lib().af_genome_align(self._h, text, offs.ctypes.data, len(blobs), ctypes.byref(p) if p is not None else None,
                                    reads_per_pass, hits.ctypes.data, ctypes.byref(n), ctypes.byref(st)))
        self.last_stats = {f: getattr(st, f) for f, _ in GenomeStats._fields_}
        return hits[: n.value].copy()

    def locate(self, pos):
        """1-based position in the concatenation -> (contig name, 1-based position in the contig)."""
        i = bisect.bisect_right(self._starts, pos - 1) - 1
        name, start, _ = self.contigs[max(i, 0)]
        return name, pos - start

    def sam_lines(self, names, seqs, hits):
        """The reads as `bwa mem` would print them: one line per read, in input order."""
        by_read = {int(h["read_id"]): h for h in hits}
        out = []
        for i, (name, seq) in enumerate(zip(names, seqs)):
            h = by_read.get(i)
            if h is None:
                out.append("\t".join([name, "4", "*", "0", "0", "*", "*", "0", "0", seq, "*"]) + "\n")
                continue
            strand = int(h["score_strand"]) & 1
            clip_l, m, clip_r = int(h["clip_l"]), int(h["m_len"]), int(h["clip_r"])
            # the contig is the one under the middle of the block; an end-to-end extension may run one mismatch
            # into the separator, which bwa cannot do (the contig ends there): those bases are soft-clipped
            i = max(bisect.bisect_right(self._starts, int(h["pos"]) - 1 + m // 2) - 1, 0)
            chrom, start, ln = self.contigs[i]
            pos = int(h["pos"]) - start
            if pos < 1:
                clip_l, m, pos = clip_l + (1 - pos), m - (1 - pos), 1
            if pos + m - 1 > ln:
                over = pos + m - 1 - ln
                clip_r, m = clip_r + over, m - over
            cigar = "".join("%d%s" % (n, op) for n, op in ((clip_l, "S"), (m, "M"), (clip_r, "S")) if n)
            out.append("\t".join([name, "16" if strand else "0", chrom, str(pos), "60", cigar, "*", "0", "0",
                                  _revcomp(seq) if strand else seq, "*", "AS:i:%d" % (int(h["score_strand"]) >> 1)]) + "\n")
        return out


_cache = {}


def genome_for(path, device=0):
    """One resident genome per (FASTA path, device) and process: del_too_many_reads runs once per gene
    (and per cell), the genome is loaded once."""
    key = (str(path), device)
    if key not in _cache:
        _cache[key] = Genome.from_fasta(path, device)
    return _cache[key]


def genome_sam(f_ref, fasta_records, device=0):
    """fasta_records: [(name, seq)] as written to <w>_del_tmp.fa -> SAM lines of their genome alignment."""
    g = genome_for(f_ref, device)
    names = [r[0] for r in fasta_records]
    seqs = [r[1] for r in fasta_records]
    usable = [len(s) <= _lib.GENOME_MAX_READ_LEN for s in seqs]
    hits = g.align([s if ok else "" for s, ok in zip(seqs, usable)])
    return g.sam_lines(names, seqs, hits)
