"""Generate the committed golden vectors.  Runs ONLY in the build container (it reads
/root/reference); the GPU box never runs it.

  bundled_c1.npz       the reference's bundled sample (test/target_gene.fasta,
                       test/test_sample_{1,2}.fastq.gz) re-encoded: anchor, read names, 2-bit reads;
                       plus the frozen oracle's hit records for it.
  ref_functions.json   known answers of the REFERENCE's own record interpretation
                       (functions.py:656 deal_cigar, :498 reverse, :892 contact_reads), obtained
                       by importing /root/reference/functions.py through oracle/ref_bridge.py.

usage: python tests/golden/make_goldens.py
"""
import gzip
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from oracle.ref_bridge import REFERENCE_ROOT, load_reference_functions  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def read_fastq(path):
    with gzip.open(path, "rt") as fh:
        lines = fh.read().split("\n")
    names = [l[1:] for l in lines[0::4] if l]
    seqs = [l for l in lines[1::4] if l]
    quals = [l for l in lines[3::4] if l]
    return names, seqs, quals


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def pseudo_sam(name, gene, pos, cigar, seq):
    # the 11-column line del_too_many_reads writes (functions.py:735) and contact_reads reads
    return "\t".join([name, "0", gene, str(pos), "60", cigar, "=", "1111", "0", seq, "A"]) + "\n"


def cigar_of(h):
    c = ""
    if h["clip_l"]:
        c += "%dS" % h["clip_l"]
    c += "%dM" % h["m_len"]
    if h["clip_r"]:
        c += "%dS" % h["clip_r"]
    return c


def dump_split(b):
    return {"chrom": b.chrom, "breakpoint": int(b.breakpoint), "type": b.type_, "cnt": int(b.cnt),
            "reads": list(b.reads), "seq_left": b.seq_left, "seq_right": b.seq_right}


def run_contact_reads(ref, lines):
    with tempfile.NamedTemporaryFile("w", suffix=".sam", delete=False) as fh:
        fh.writelines(lines)
        path = fh.name
    try:
        return [dump_split(b) for b in ref.contact_reads(path, "", "", "1")]
    finally:
        os.remove(path)


def main():
    ref = load_reference_functions()
    # ---- bundled sample -------------------------------------------------------------------
    fa = open(os.path.join(REFERENCE_ROOT, "test/target_gene.fasta")).read().split("\n")
    header, anchor = fa[0], "".join(fa[1:])
    n1, s1, q1 = read_fastq(os.path.join(REFERENCE_ROOT, "test/test_sample_1.fastq.gz"))
    n2, s2, q2 = read_fastq(os.path.join(REFERENCE_ROOT, "test/test_sample_2.fastq.gz"))
    assert len(n1) == len(n2) == 11258
    assert all(set(s) <= set("ACGT") and len(s) == 101 for s in s1 + s2)
    assert all(q == "2" * 101 for q in q1 + q2)
    codes = np.stack([oracle.encode(s) for pair in zip(s1, s2) for s in pair])  # read_id = 2*pair + mate
    packed = np.zeros((codes.shape[0], 26), dtype=np.uint8)                     # 4 bases per byte
    for i in range(101):
        packed[:, i // 4] |= codes[:, i] << (2 * (i % 4))
    hits = oracle.anchor_reads(oracle.encode(anchor), codes, threads=8)
    np.savez_compressed(os.path.join(HERE, "bundled_c1.npz"), header=np.array(header), anchor=np.array(anchor),
                        names1="\n".join(n1), names2="\n".join(n2), reads_2bit=packed, read_len=np.int32(101),
                        qual_char=np.array("2"), oracle_hits=hits)
    print("bundled: %d pairs, %d oracle hits, %d pairs with a hit" % (len(n1), len(hits), len(set(hits["read_id"] >> 1))))

    # ---- reference record interpretation ---------------------------------------------------
    out = {"deal_cigar": [], "reverse": [], "contact_reads": []}
    rng = random.Random(7)
    cigars = ["40S61M", "61M40S", "10S91M", "101M", "5S90M6S", "20H81M", "50M2D41M10S", "30M1I60M10S",
              "25S30M500N46M", "15S86M", "14S87M", "86M15S", "1S100M", "100M1S", "30M2I20M3D49M", "33S40M28S"]
    for c in cigars:
        seq = "".join(rng.choice("ACGT") for _ in range(101))
        res, seq2 = ref.deal_cigar(c, seq)
        out["deal_cigar"].append({"cigar": c, "seq": seq, "ops": [[int(a), int(b), o] for a, b, o in res], "seq_out": seq2})
    for s in ["ACGTN", "", "AAAAC", "NHACGT", "GATTACA"]:
        out["reverse"].append({"seq": s, "out": ref.reverse(s)})

    # (a) hand-made records incl. the SURVEY appendix cases
    base = "".join(rng.choice("ACGT") for _ in range(101))
    hand = [pseudo_sam("r1", "BCR", 100, "40S61M", base), pseudo_sam("r2", "BCR", 100, "61M40S", base),
            pseudo_sam("r3", "BCR", 100, "10S91M", base), pseudo_sam("r4", "BCR", 120, "50M2D41M10S", base),
            pseudo_sam("r5", "BCR", 100, "40S61M", base), pseudo_sam("r6", "BCR", 100, "40S61M", "T" + base[1:]),
            pseudo_sam("r7", "BCR", 101, "41S60M", base), pseudo_sam("r8", "BCR", 5, "101M", base)]
    out["contact_reads"].append({"name": "hand", "lines": hand, "out": run_contact_reads(ref, hand)})

    # (b) the bundled sample's oracle records, in the order the product emits them:
    #     (POS, strand, read_id)
    names = [n.split()[0] for n in n1]
    names = [n[:-2] if n.endswith(("/1", "/2")) else n for n in names]
    order = np.lexsort((hits["read_id"], hits["score_strand"] & 1, hits["pos"]))
    lines = []
    for h in hits[order]:
        rid = int(h["read_id"])
        seq = (s1, s2)[rid & 1][rid >> 1]
        if h["score_strand"] & 1:
            seq = revcomp(seq)
        lines.append(pseudo_sam(names[rid >> 1], "BCR", int(h["pos"]), cigar_of(h), seq))
    out["contact_reads"].append({"name": "bundled_c1", "lines": None, "n_lines": len(lines),
                                 "out": run_contact_reads(ref, lines)})

    # (c) synthetic split reads around a few junctions, with mismatches, to exercise
    #     combine_split_reads' merging (functions.py:771-889)
    lines = []
    acds = anchor
    k = 0
    for bp in (500, 502, 503, 900, 2568, 3235, 3236, 3240):
        for rep in range(rng.randint(2, 9)):
            typ = rng.choice(["SM", "MS"])
            m = rng.randint(30, 80)
            s = 101 - m
            partner = "".join(random.Random(bp * 7 + (typ == "SM")).choice("ACGT") for _ in range(101))
            if typ == "SM":
                seq = partner[-s:] + acds[bp - 1: bp - 1 + m]
                cg, pos = "%dS%dM" % (s, m), bp
            else:
                seq = acds[bp - m: bp] + partner[:s]
                cg, pos = "%dM%dS" % (m, s), bp - m + 1
            seq = list(seq)
            for _ in range(rng.randint(0, 2)):
                seq[rng.randrange(101)] = rng.choice("ACGT")
            lines.append((pos, pseudo_sam("syn%d" % k, "BCR", pos, cg, "".join(seq))))
            k += 1
    lines = [l for _, l in sorted(lines, key=lambda t: t[0])]
    out["contact_reads"].append({"name": "synthetic_junctions", "lines": lines, "out": run_contact_reads(ref, lines)})

    with open(os.path.join(HERE, "ref_functions.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    for c in out["contact_reads"]:
        print("contact_reads[%s]: %d groups" % (c["name"], len(c["out"])))


if __name__ == "__main__":
    main()
