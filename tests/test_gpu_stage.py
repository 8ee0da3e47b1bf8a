"""The drop-in boundary end to end on the GPU: the reference's CLI flags in, the anchoring
stage's files out, on the bundled sample (config 1) and in single-cell layout (config 5 shape)."""
import gzip
import json
import os

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _write_bundled_fastqs(bundled, d, stem="test_sample", sel=None):
    sel = range(len(bundled["seqs1"])) if sel is None else sel
    q = bundled["qual_char"] * bundled["read_len"]
    p1, p2 = os.path.join(d, stem + "_1.fastq.gz"), os.path.join(d, stem + "_2.fastq.gz")
    with gzip.open(p1, "wt") as f1, gzip.open(p2, "wt") as f2:
        for i in sel:
            f1.write("@%s\n%s\n+\n%s\n" % (bundled["names1"][i], bundled["seqs1"][i], q))
            f2.write("@%s\n%s\n+\n%s\n" % (bundled["names2"][i], bundled["seqs2"][i], q))
    return p1, p2


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def test_bulk_cli_on_the_bundled_sample(bundled, tmp_path):
    from anchored_fusion_b200.bam import read_bam, sam_line
    from anchored_fusion_b200.cli import main_bulk
    from anchored_fusion_b200.functions import contact_reads
    from anchored_fusion_b200.records import cigar_string, read_names, sort_hits
    d = str(tmp_path)
    fa = os.path.join(d, "target_gene.fasta")
    with open(fa, "w") as fh:
        fh.write(bundled["header"] + "\n")
        for i in range(0, len(bundled["anchor"]), 70):
            fh.write(bundled["anchor"][i:i + 70] + "\n")
    p1, p2 = _write_bundled_fastqs(bundled, d)
    out = os.path.join(d, "out")
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out,
                      "--not_filter_false_positive", "--thread", "4"]) == 0
    w = os.path.join(out, "BCR_fusion", "work_dir", "BCR_fusion")
    assert open(w + "_anchored_gene_sequence.fa").read().startswith(">BCR\n")
    want = sort_hits(bundled["oracle_hits"])
    names = read_names(bundled["names1"])
    # <w>_anchored_reads.bam == samtools view -F 772 of the oracle's records, coordinate order
    _, refs, recs = read_bam(w + "_anchored_reads.bam")
    assert refs == [("BCR", 6783)] and len(recs) == len(want)
    have = set(int(r) for r in want["read_id"])
    for r, h in zip(recs, want):
        rid = int(h["read_id"])
        seq = (bundled["seqs1"], bundled["seqs2"])[rid & 1][rid >> 1]
        rev = int(h["score_strand"]) & 1
        assert (r["qname"], r["pos"], r["cigar"], r["seq"]) == (names[rid >> 1], int(h["pos"]), cigar_string(h),
                                                                _rc(seq) if rev else seq)
        assert r["flag"] & 0x904 == 0 and bool(r["flag"] & 0x10) == bool(rev)
        assert bool(r["flag"] & 0x8) == ((rid ^ 1) not in have) and bool(r["flag"] & 0x80) == bool(rid & 1)
    # tmp_1 / tmp_2: pairs with exactly one anchored mate, original orientation, /1 /2 names
    t1 = open(w + "_tmp_1.fastq").read().split("\n")
    t2 = open(w + "_tmp_2.fastq").read().split("\n")
    half = [int(h["read_id"]) for h in want if (int(h["read_id"]) ^ 1) not in have]
    assert len(half) == 33 and len(t1) == len(t2) == 4 * len(half) + 1
    for k, rid in enumerate(half):
        assert t1[4 * k] == "@%s/%d" % (names[rid >> 1], (rid & 1) + 1)
        assert t1[4 * k + 1] == (bundled["seqs1"], bundled["seqs2"])[rid & 1][rid >> 1]
        assert t2[4 * k] == "@%s/%d" % (names[rid >> 1], ((rid ^ 1) & 1) + 1)
        assert t2[4 * k + 1] == (bundled["seqs1"], bundled["seqs2"])[(rid ^ 1) & 1][rid >> 1]
    # realign.bam: the reference's three samtools filters select exactly those sets
    _, _, rl = read_bam(w + "_realign_reads.bam")
    sel_a = [r for r in rl if r["flag"] & 8 and not r["flag"] & 260]       # -f 8 -F 260
    sel_b = [r for r in rl if r["flag"] & 4 and not r["flag"] & 264]       # -f 4 -F 264
    sel_c = [r for r in rl if not r["flag"] & 772]                         # -F 772
    assert len(sel_a) == len(sel_b) == len(half) and [sam_line(r) for r in sel_c] == [sam_line(r) for r in recs]
    assert [r["qname"] for r in sel_a] == [r["qname"] for r in sel_b]
    # split points through contact_reads == the REFERENCE's contact_reads on the oracle's records
    sam11 = os.path.join(d, "anchored_reads.sam")
    with open(sam11, "w") as o:
        for r in recs:
            o.write("\t".join([r["qname"], "0", "BCR", str(r["pos"]), "60", r["cigar"], "=", "1111", "0", r["seq"], "A"]) + "\n")
    got = [{"chrom": g.chrom, "breakpoint": int(g.breakpoint), "type": g.type_, "cnt": int(g.cnt), "reads": list(g.reads),
            "seq_left": g.seq_left, "seq_right": g.seq_right} for g in contact_reads(sam11, "", "", "1")]
    golden = json.load(open(os.path.join(GOLDEN, "ref_functions.json")))
    assert got == next(c for c in golden["contact_reads"] if c["name"] == "bundled_c1")["out"]
    sp = open(w + "_split_points.txt").read().split("\n")
    assert sp[1].split("\t")[:4] == ["BCR", "3235", "MS", "49"]
    # second run: outputs exist -> skipped, like the reference's existence guards
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out]) == 0


def test_singlecell_cli_matches_bulk_union(bundled, tmp_path):
    """Per-cell FASTQ pairs (the reference's single-cell layout): the union of the cells' anchored
    reads equals the bulk result on the concatenation."""
    from anchored_fusion_b200.bam import read_bam
    from anchored_fusion_b200.cli import main_singlecell
    from anchored_fusion_b200.records import read_names
    d = str(tmp_path)
    fa = os.path.join(d, "t.fa")
    open(fa, "w").write(bundled["header"] + "\n" + bundled["anchor"] + "\n")
    cells = os.path.join(d, "cells")
    os.mkdir(cells)
    n = len(bundled["seqs1"])
    parts = {"cellA": range(0, n, 3), "cellB": range(1, n, 3), "cellC": range(2, n, 3)}
    for c, sel in parts.items():
        _write_bundled_fastqs(bundled, cells, c, sel)
    out = os.path.join(d, "out")
    assert main_singlecell(["--file_anchored_cds", fa, "--fastq_dir", cells, "--out_folder", out]) == 0
    names = read_names(bundled["names1"])
    want = sorted((names[int(h["read_id"]) >> 1], int(h["read_id"]) & 1, int(h["pos"])) for h in bundled["oracle_hits"])
    got = []
    for c in parts:
        _, _, recs = read_bam(os.path.join(out, "BCR", "work_dir", c, "BCR_fusion_anchored_reads.bam"))
        got += [(r["qname"], 1 if r["flag"] & 0x80 else 0, r["pos"]) for r in recs]
    assert sorted(got) == want


def test_two_genes_one_pass_equals_two_separate_runs(bundled, tmp_path):
    """A two-record CDS file: the bulk driver decodes the FASTQ pair once and scans it for both
    anchors; each gene's files equal those of a run with that gene alone."""
    from anchored_fusion_b200.bam import read_bam, sam_line
    from anchored_fusion_b200.cli import main_bulk
    d = str(tmp_path)
    a = bundled["anchor"]
    second = a[2500:6000][::-1].translate(str.maketrans("ACGT", "TGCA"))     # reverse complement of a BCR slice
    fa2 = os.path.join(d, "two.fa")
    open(fa2, "w").write(bundled["header"] + "\n" + a + "\n>NM_000000.1 RCBCR [organism=Homo sapiens]\n" + second + "\n")
    p1, p2 = _write_bundled_fastqs(bundled, d)
    out2 = os.path.join(d, "out2")
    assert main_bulk(["--file_anchored_cds", fa2, "--fastq1", p1, "--fastq2", p2, "--out_folder", out2]) == 0
    singles = {}
    for gene, seq in (("BCR", a), ("RCBCR", second)):
        fa1 = os.path.join(d, gene + ".fa")
        open(fa1, "w").write(">NM_1.1 " + gene + "\n" + seq + "\n")
        out1 = os.path.join(d, "out_" + gene)
        assert main_bulk(["--file_anchored_cds", fa1, "--fastq1", p1, "--fastq2", p2, "--out_folder", out1]) == 0
        singles[gene] = out1
    for gene in ("BCR", "RCBCR"):
        w2 = os.path.join(out2, gene + "_fusion", "work_dir", gene + "_fusion")
        w1 = os.path.join(singles[gene], gene + "_fusion", "work_dir", gene + "_fusion")
        r2, r1 = read_bam(w2 + "_anchored_reads.bam")[2], read_bam(w1 + "_anchored_reads.bam")[2]
        assert len(r2) > 300 and [sam_line(r) for r in r2] == [sam_line(r) for r in r1]
        for suffix in ("_tmp_1.fastq", "_tmp_2.fastq", "_split_points.txt"):
            assert open(w2 + suffix).read() == open(w1 + suffix).read()


def test_singlecell_two_genes_cells_dealt_to_ranks(bundled, tmp_path, monkeypatch):
    """Config-5 shape: several cells, two anchored genes.  Each cell's FASTQ pair is decoded once for both
    genes, and under torchrun's RANK / WORLD_SIZE the cells are dealt to the ranks with no exchange: two
    'ranks' (run one after the other here) leave exactly the files of a single-process run."""
    from anchored_fusion_b200.bam import read_bam, sam_line
    from anchored_fusion_b200.cli import main_singlecell
    d = str(tmp_path)
    a = bundled["anchor"]
    second = a[2500:6000][::-1].translate(str.maketrans("ACGT", "TGCA"))
    fa = os.path.join(d, "two.fa")
    open(fa, "w").write(bundled["header"] + "\n" + a + "\n>NM_000000.1 RCBCR [organism=Homo sapiens]\n" + second + "\n")
    cells = os.path.join(d, "cells")
    os.mkdir(cells)
    n = len(bundled["seqs1"])
    names = ["c%02d" % i for i in range(5)]
    for i, c in enumerate(names):
        _write_bundled_fastqs(bundled, cells, c, range(i, n, 5))
    out1, out2 = os.path.join(d, "single"), os.path.join(d, "ranks")
    assert main_singlecell(["--file_anchored_cds", fa, "--fastq_dir", cells, "--out_folder", out1]) == 0
    for rank in (0, 1):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "2")
        monkeypatch.setenv("LOCAL_RANK", "0")
        assert main_singlecell(["--file_anchored_cds", fa, "--fastq_dir", cells, "--out_folder", out2]) == 0
        done = [c for c in names if os.path.exists(os.path.join(out2, "BCR", "work_dir", c, "BCR_fusion_anchored_reads.bam"))]
        assert done == (names[0::2] if rank == 0 else names)          # rank 0 took cells 0, 2, 4 only
    total = 0
    for gene in ("BCR", "RCBCR"):
        for c in names:
            w1 = os.path.join(out1, gene, "work_dir", c, gene + "_fusion")
            w2 = os.path.join(out2, gene, "work_dir", c, gene + "_fusion")
            r1, r2 = read_bam(w1 + "_anchored_reads.bam")[2], read_bam(w2 + "_anchored_reads.bam")[2]
            assert [sam_line(r) for r in r1] == [sam_line(r) for r in r2]
            total += len(r1)
            for suffix in ("_tmp_1.fastq", "_tmp_2.fastq", "_split_points.txt"):
                assert open(w1 + suffix).read() == open(w2 + suffix).read()
    assert total > 1500
