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


def run_find_fine_block(ref, lines):
    """Reference Find_fine_block (functions.py:506) with blat stubbed: captures the FASTA its first
    loop writes and the spanning_anchored objects it builds."""
    made, captured = [], {}
    real_cls, real_system = ref.spanning_anchored, ref.os.system

    class Spy(real_cls):
        def __init__(self, type_, left_length, right_length, read_name):
            super().__init__(type_, left_length, right_length, read_name)
            made.append([type_, int(left_length), int(right_length), read_name])

    def fake_system(cmd):
        assert cmd.startswith("blat "), cmd
        parts = cmd.split()
        captured["fasta"] = open(parts[-2]).read()
        open(parts[-1], "w").close()
        return 0

    tmp = tempfile.mkdtemp()
    f_read = os.path.join(tmp, "reads.sam")
    with open(f_read, "w") as fh:
        fh.writelines(lines)
    ref.spanning_anchored, ref.os.system = Spy, fake_system
    try:
        ref.Find_fine_block(f_read, "genome.fa", os.path.join(tmp, "w"), None, [], {})
    finally:
        ref.spanning_anchored, ref.os.system = real_cls, real_system
    return {"lines": lines, "candidates": made, "fasta": captured["fasta"]}


def random_cigar(rng, total, kinds):
    """A CIGAR over `total` query bases built from the op kinds given (M always present)."""
    kind = rng.choice(kinds)
    if kind == "M":
        return "%dM" % total
    if kind in ("SM", "HM"):
        c = rng.randint(1, total - 20)
        return "%d%s%dM" % (c, kind[0], total - c)
    if kind in ("MS", "MH"):
        c = rng.randint(1, total - 20)
        return "%dM%d%s" % (total - c, c, kind[1])
    if kind == "SMS":
        a, b = rng.randint(1, 30), rng.randint(1, 30)
        return "%dS%dM%dS" % (a, total - a - b, b)
    if kind == "HMH":
        a, b = rng.randint(1, 30), rng.randint(1, 30)
        return "%dH%dM%dH" % (a, total - a - b, b)
    if kind == "MDM":
        a = rng.randint(10, total - 10)
        return "%dM%dD%dM" % (a, rng.randint(1, 5), total - a)
    if kind == "MIM":
        a, i = rng.randint(10, total - 20), rng.randint(1, 4)
        return "%dM%dI%dM" % (a, i, total - a - i)
    if kind == "MNM":
        a = rng.randint(10, total - 10)
        return "%dM%dN%dM" % (a, rng.randint(50, 5000), total - a)
    if kind == "SMDMS":
        a, b, m = rng.randint(1, 20), rng.randint(1, 20), rng.randint(10, 30)
        return "%dS%dM2D%dM%dS" % (a, m, total - a - b - m, b)
    raise ValueError(kind)


def run_del_too_many(ref, rng, case):
    """Reference del_too_many_reads (functions.py:705) with `samtools view` replaced by prepared
    anchored records and the genome `bwa mem` by a prepared SAM text: pins the 2-op selection
    (the FASTA handed to bwa) and the contiguity decision (out_sam) for that text."""
    L = 101
    anchored = []
    for k in range(60):
        seq = "".join(rng.choice("ACGT") for _ in range(L))
        cg = random_cigar(rng, L, ["SM", "MS", "SM", "MS", "M", "SMS", "HM", "MDM"])
        flag = rng.choice([0, 16, 83, 99, 147, 163])
        anchored.append("\t".join(["q%d_%d" % (case, k), str(flag), "BCR", str(rng.randint(1, 6000)), "60", cg, "=",
                                   str(rng.randint(1, 6000)), "0", seq, "2" * L]))
    captured = {}
    genome = []
    real_system, real_popen = ref.os.system, ref.os.popen

    class FakePipe:
        def __init__(self, text):
            self.text = text

        def read(self):
            return self.text

    def fake_popen(cmd):
        assert cmd.startswith("samtools view "), cmd
        return FakePipe("\n".join(anchored) + "\n")

    def fake_system(cmd):
        assert cmd.startswith("bwa mem "), cmd
        left, tmp_sam = cmd.split(" > ")
        tmp_fa = left.split()[-1]
        captured["fasta"] = open(tmp_fa).read()
        fa = captured["fasta"].split("\n")
        if case != 5:
            genome.append("@SQ\tSN:chr9\tLN:141213431\n")
            genome.append("@PG\tID:bwa\tPN:bwa\n")
        for tag, seq in zip(fa[0::2], fa[1::2]):
            tag = tag[1:]
            n_lines = rng.choice([1, 1, 2, 3])
            for j in range(n_lines):
                flag = rng.choice([0, 16, 0, 16, 2048, 2064, 256, 272])
                if j == 0 and rng.random() < 0.15:
                    genome.append("\t".join([tag, "4", "*", "0", "0", "*", "*", "0", "0", seq, "*"]) + "\n")
                    break
                cg = random_cigar(rng, len(seq), ["M", "SM", "MS", "SM", "MS", "HM", "MH", "SMS", "HMH", "MDM", "MIM", "MNM", "SMDMS"])
                s = revcomp(seq) if flag & 16 else seq
                if "H" in cg:      # hard clips are not part of SEQ
                    ops = [(int(n), o) for n, o in __import__("re").findall(r"(\d+)([A-Z])", cg)]
                    lead = ops[0][0] if ops[0][1] == "H" else 0
                    tail = ops[-1][0] if ops[-1][1] == "H" else 0
                    s = s[lead: len(s) - tail]
                genome.append("\t".join([tag, str(flag), "chr9", str(rng.randint(1, 10 ** 8)), str(rng.choice([0, 60])), cg,
                                         "*", "0", "0", s, "*"]) + "\n")
        if case == 4:
            genome.append("@CO\ta trailing header line: the last read is then never flushed\n")
        open(tmp_sam, "w").writelines(genome)
        return 0

    tmp = tempfile.mkdtemp()
    out_sam = os.path.join(tmp, "out.sam")
    ref.os.system, ref.os.popen = fake_system, fake_popen
    try:
        ref.del_too_many_reads("anchored.bam", out_sam, os.path.join(tmp, "w"), "genome.fa", "1")
    finally:
        ref.os.system, ref.os.popen = real_system, real_popen
    return {"anchored": anchored, "fasta": captured["fasta"], "genome_sam": genome, "out_sam": open(out_sam).read()}


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

    # ---- rows 8(f)-2 / 8(f)-3: the loops right behind the path, run in the REFERENCE with its two
    #      shell-outs replaced (blat -> empty .psl, genome bwa mem -> a prepared SAM text) -------------
    out["find_fine_block_first_loop"] = run_find_fine_block(ref, out["contact_reads"][2]["lines"] + hand)
    out["del_too_many_reads"] = [run_del_too_many(ref, rng, case) for case in range(6)]

    with open(os.path.join(HERE, "ref_functions.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    for c in out["contact_reads"]:
        print("contact_reads[%s]: %d groups" % (c["name"], len(c["out"])))


if __name__ == "__main__":
    main()
