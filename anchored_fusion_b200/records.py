"""From hit records to the files the rest of the pipeline reads.

The anchoring stage of the reference ends in three files (Anchored_Fusion.py:182-194):
  <work>_anchored_reads.bam  = samtools view -F 772: every mapped primary read, coordinate sorted
  <work>_tmp_1.fastq         = samtools view -f 8 -F 260 | samtools fastq: anchored mates whose mate is not
  <work>_tmp_2.fastq         = samtools view -f 4 -F 264 | samtools fastq: those unanchored mates
This module derives the same three sets from the GPU's hit list.  Order: (POS, strand, read_id)
-- samtools sorts by (tid, pos, reverse flag) and leaves ties unspecified; pinning them makes
the output independent of batch size and GPU count.
"""
import numpy as np

_RC = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def revcomp(seq):
    if isinstance(seq, str):
        return seq.encode().translate(_RC)[::-1].decode()
    return bytes(seq).translate(_RC)[::-1]


def read_names(fastq_names):
    """QNAMEs as bwa prints them: up to the first blank, trailing /1 or /2 dropped."""
    out = []
    for n in fastq_names:
        n = n.split()[0] if n else n
        out.append(n[:-2] if n.endswith(("/1", "/2")) else n)
    return out


def sort_hits(hits):
    """(POS, strand, read_id) order."""
    order = np.lexsort((hits["read_id"], hits["score_strand"] & 1, hits["pos"]))
    return hits[order]


def cigar_string(h):
    s = "%dS" % h["clip_l"] if h["clip_l"] else ""
    s += "%dM" % h["m_len"]
    if h["clip_r"]:
        s += "%dS" % h["clip_r"]
    return s


def sam_flag(h, mate_hit):
    """Paired-end FLAG as bwa would set it, minus 0x2 (proper pair needs bwa's insert-size model)."""
    rid = int(h["read_id"])
    f = 0x1 | (0x80 if rid & 1 else 0x40)
    if h["score_strand"] & 1:
        f |= 0x10
    if mate_hit is None:
        f |= 0x8
    elif mate_hit["score_strand"] & 1:
        f |= 0x20
    return f


def pseudo_sam_lines(hits, gene, names, seqs1, seqs2):
    """The 11-column text contact_reads / Find_fine_block parse (format of functions.py:735):
    QNAME 0 RNAME POS 60 CIGAR = 1111 0 SEQ A, SEQ in anchor-forward orientation."""
    lines = []
    for h in sort_hits(hits):
        rid = int(h["read_id"])
        seq = (seqs1, seqs2)[rid & 1][rid >> 1]
        if h["score_strand"] & 1:
            seq = revcomp(seq)
        lines.append("\t".join([names[rid >> 1], "0", gene, str(int(h["pos"])), "60", cigar_string(h), "=", "1111",
                                "0", seq, "A"]) + "\n")
    return lines


def split_sets(hits):
    """(anchored, half_pairs): anchored = all hits sorted; half_pairs = [(anchored_hit, read_id of the
    unanchored mate)] for pairs with exactly one anchored mate, in the anchored mate's order."""
    srt = sort_hits(hits)
    have = set(int(r) for r in hits["read_id"])
    half = [(h, int(h["read_id"]) ^ 1) for h in srt if (int(h["read_id"]) ^ 1) not in have]
    return srt, half
