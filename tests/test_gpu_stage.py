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
    # the two BAMs against the specification, with a reader that shares no code with the writer
    from bam_spec_check import check_bam
    chk = check_bam(w + "_anchored_reads.bam")
    assert chk["refs"] == [("BCR", 6783)] and [(r["qname"], r["pos"], r["cigar"], r["seq"]) for r in chk["records"]] == \
        [(r["qname"], r["pos"], r["cigar"], r["seq"]) for r in recs]
    assert len(check_bam(w + "_realign_reads.bam", expect_sorted=False)["records"]) == len(rl)
    # second run: outputs exist -> skipped, like the reference's existence guards
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out]) == 0


def test_bulk_cli_with_2x300_reads(tmp_path):
    """2x300 MiSeq-shaped input (ragged 250..300, a few N) through the reference's CLI flags: the anchored-reads BAM holds
    the oracle's records, read for read (the round-1 limit of 256 bases made this input fail)."""
    import numpy as np
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.bam import read_bam
    from anchored_fusion_b200.cli import main_bulk
    from anchored_fusion_b200.records import cigar_string, sort_hits
    from oracle import oracle
    d = str(tmp_path)
    spec = af.synth_spec(seed=31, ref_len=300_000, anchor_start=100_000, anchor_len=5000, read_len=300,
                         frag_mean=650, frag_sd=40, sub_ppm=12_000, fusion_ppm=40_000)
    anchor = af.synth_anchor(spec)
    anchor = anchor.decode() if isinstance(anchor, bytes) else anchor
    n = 20_000
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    rng = np.random.default_rng(3)
    lens = rng.integers(250, 301, (2, n))
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    seqs = [[], []]
    for m, mm in enumerate((m1, m2)):
        for i in range(n):
            r = mm[i, : lens[m, i]].copy()
            if i % 50 == 7:
                r[int(rng.integers(0, len(r)))] = 4
            seqs[m].append(lut[r].tobytes().decode())
    fa = os.path.join(d, "target_gene.fasta")
    with open(fa, "w") as fh:
        fh.write(">ABL9\n")
        for i in range(0, len(anchor), 70):
            fh.write(anchor[i:i + 70] + "\n")
    p1, p2 = os.path.join(d, "s_1.fastq.gz"), os.path.join(d, "s_2.fastq.gz")
    with gzip.open(p1, "wt", compresslevel=1) as f1, gzip.open(p2, "wt", compresslevel=1) as f2:
        for i in range(n):
            f1.write("@r%d/1\n%s\n+\n%s\n" % (i, seqs[0][i], "F" * len(seqs[0][i])))
            f2.write("@r%d/2\n%s\n+\n%s\n" % (i, seqs[1][i], "F" * len(seqs[1][i])))
    out = os.path.join(d, "out")
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out,
                      "--not_filter_false_positive", "--thread", "4"]) == 0
    w = os.path.join(out, "ABL9_fusion", "work_dir", "ABL9_fusion")
    codes = np.full((2 * n, 300), 4, dtype=np.uint8)
    ol = np.empty(2 * n, dtype=np.uint16)
    for i in range(n):
        for m in range(2):
            c = oracle.encode(seqs[m][i])
            codes[2 * i + m, : len(c)] = c
            ol[2 * i + m] = len(c)
    want = sort_hits(oracle.anchor_reads(oracle.encode(anchor), codes, lens=ol, threads=4))
    _, refs, recs = read_bam(w + "_anchored_reads.bam")
    assert refs == [("ABL9", 5000)] and len(recs) == len(want) > 500
    assert int(want["m_len"].max()) > 256
    for r, h in zip(recs, want):
        rid = int(h["read_id"])
        seq = seqs[rid & 1][rid >> 1]
        assert (r["qname"], r["pos"], r["cigar"]) == ("r%d" % (rid >> 1), int(h["pos"]), cigar_string(h))
        assert r["seq"] == (_rc(seq) if int(h["score_strand"]) & 1 else seq)


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


# ---- round 2 ---------------------------------------------------------------------------------------------------
def _bulk_files(w):
    return {s: open(w + s, "rb").read() for s in ("_anchored_reads.bam", "_realign_reads.bam", "_tmp_1.fastq", "_tmp_2.fastq",
                                                 "_anchored_reads.raw.sam", "_split_points.txt")}


def test_bulk_cli_two_torchrun_ranks_leave_byte_identical_files(bundled, tmp_path):
    """Multi-GPU through the drop-in boundary: `torchrun --nproc-per-node 2 Anchored_Fusion.py ...` deals the read
    batches to the ranks, each rank anchors its share on its GPU (the ranks share GPU 0 on a one-GPU box), rank 0
    gets every rank's records and writes the one set of files -- byte for byte those of the one-process run."""
    import subprocess
    import sys
    from conftest import ROOT
    from anchored_fusion_b200.cli import main_bulk
    d = str(tmp_path)
    a = bundled["anchor"]
    second = a[2500:6000][::-1].translate(str.maketrans("ACGT", "TGCA"))
    fa = os.path.join(d, "two.fa")
    open(fa, "w").write(bundled["header"] + "\n" + a + "\n>NM_000000.1 RCBCR [organism=Homo sapiens]\n" + second + "\n")
    p1, p2 = _write_bundled_fastqs(bundled, d)
    os.environ["AF_BATCH_PAIRS"] = "1024"                     # 11 batches: every rank gets several
    try:
        out1 = os.path.join(d, "one")
        assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out1, "--thread", "3"]) == 0
        out2 = os.path.join(d, "two")
        env = dict(os.environ, AF_BATCH_PAIRS="1024")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29571", os.path.join(ROOT, "Anchored_Fusion.py"), "--file_anchored_cds", fa, "--fastq1", p1,
               "--fastq2", p2, "--out_folder", out2, "--thread", "3"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    finally:
        os.environ.pop("AF_BATCH_PAIRS", None)
    for gene in ("BCR", "RCBCR"):
        w1 = os.path.join(out1, gene + "_fusion", "work_dir", gene + "_fusion")
        w2 = os.path.join(out2, gene + "_fusion", "work_dir", gene + "_fusion")
        f1, f2 = _bulk_files(w1), _bulk_files(w2)
        assert len(f1["_anchored_reads.raw.sam"]) > 10_000
        for k in f1:
            assert f1[k] == f2[k], (gene, k)
        assert not [f for f in os.listdir(os.path.dirname(w2)) if f.endswith(".partial")]


def test_h2d_traffic_does_not_depend_on_the_number_of_genes(bundled):
    """f4: a batch is copied to the GPU once and scanned for every anchor index while resident."""
    import numpy as np
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import anchor_host_multi
    a = bundled["anchor"]
    seqs = [a, a[2500:6000][::-1].translate(str.maketrans("ACGT", "TGCA")), a[1000:4000], a[3000:]]
    idx = [af.AnchorIndex(s) for s in seqs]
    engs = [af.Anchorer(i, 0) for i in idx]
    host = af.pack_pairs(bundled["seqs1"], bundled["seqs2"], pad_byte=idx[0].pad_byte)
    singles = [e.anchor_host(host, slot_pairs=4096, n_slots=3)[0].copy() for e in engs]
    one = engs[0].pipeline_h2d_bytes()
    for e in engs:
        e.close_pipeline()
    multi = anchor_host_multi(engs, host, slot_pairs=4096, n_slots=3)
    assert engs[0].pipeline_h2d_bytes() == one == af.layout(host.max_read_len, host.n_pairs).packed_bytes
    for (h, st), want in zip(multi, singles):
        assert len(want) > 200 and np.array_equal(h.view(np.uint8), want.view(np.uint8))


def test_singlecell_20m_pairs_in_2000_cells_batched_across_cells(tmp_path):
    """configs[4] at size: 2 000 cells x 10 000 pairs of per-cell FASTQ.gz files through the single-cell CLI.  The
    cells are decoded concurrently and packed back to back into shared 1 M-pair GPU batches; the per-cell
    files must hold exactly the oracle's records of that cell's pairs, and a cell run alone through the bulk
    stage must leave byte-identical files."""
    import time
    import numpy as np
    from oracle import oracle
    from anchored_fusion_b200.bam import read_bam
    from anchored_fusion_b200.cli import main_singlecell
    from anchored_fusion_b200.stage import GeneAnchorer, anchor_stage
    n_cells, ppc = 2000, 10_000
    spec = oracle.synth_spec(seed=5, ref_len=10_000_000, anchor_start=2_000_000, anchor_len=6783, read_len=150,
                             frag_mean=300, sub_ppm=10_000, fusion_ppm=2_000)
    d = str(tmp_path)
    cells = os.path.join(d, "cells")
    os.mkdir(cells)
    f1 = [os.path.join(cells, "cell%05d_1.fastq.gz" % c) for c in range(n_cells)]
    f2 = [os.path.join(cells, "cell%05d_2.fastq.gz" % c) for c in range(n_cells)]
    threads = os.cpu_count() or 1
    for k in range(0, n_cells, 500):
        oracle.synth_fastq(spec, k * ppc, ppc, f1[k:k + 500], f2[k:k + 500], oracle.FASTQ_GZIP, 1, threads=threads)
    fa = os.path.join(d, "g.fa")
    anchor = oracle.synth_anchor(spec).decode()
    open(fa, "w").write(">NM_1.1 SYNX [organism=synthetic]\n" + anchor + "\n")
    out = os.path.join(d, "out")
    t0 = time.time()
    assert main_singlecell(["--file_anchored_cds", fa, "--fastq_dir", cells, "--out_folder", out, "--thread", "0"]) == 0
    dt = time.time() - t0
    # the oracle on all 20 M pairs, cell by cell
    acodes = oracle.encode(anchor)
    total = 0
    for k in range(0, n_cells, 200):
        reads = oracle.synth_reads(spec, k * ppc, 200 * ppc, threads=threads)
        want = oracle.anchor_reads(acodes, reads, threads=threads)
        cell_of = (want["read_id"] >> 1) // ppc
        for c in range(200):
            w = want[cell_of == c]
            _, _, recs = read_bam(os.path.join(out, "SYNX", "work_dir", "cell%05d" % (k + c), "SYNX_fusion_anchored_reads.bam"))
            got = sorted((r["pos"], 1 if r["flag"] & 0x10 else 0, r["cigar"], 1 if r["flag"] & 0x80 else 0) for r in recs)
            exp = sorted((int(h["pos"]), int(h["score_strand"]) & 1,
                          ("%dS" % h["clip_l"] if h["clip_l"] else "") + "%dM" % h["m_len"] + ("%dS" % h["clip_r"] if h["clip_r"] else ""),
                          int(h["read_id"]) & 1) for h in w)
            assert got == exp, "cell %d" % (k + c)
            total += len(recs)
    assert total > 20_000
    # three cells alone through the bulk stage: byte-identical files
    ga = GeneAnchorer(os.path.join(out, "SYNX", "work_dir", "SYNX_fusion_anchored_gene_sequence.fa"), "-1", "SYNX")
    for c in (0, 777, n_cells - 1):
        pre = os.path.join(d, "alone%d" % c)
        anchor_stage(None, f1[c], f2[c], pre, thread="2", gene_anchorer=ga)
        w = os.path.join(out, "SYNX", "work_dir", "cell%05d" % c, "SYNX_fusion")
        for s in ("_anchored_reads.bam", "_realign_reads.bam", "_tmp_1.fastq", "_tmp_2.fastq", "_anchored_reads.raw.sam"):
            assert open(pre + s, "rb").read() == open(w + s, "rb").read(), (c, s)
    rate = n_cells * ppc / dt
    print("single-cell 20 M pairs / 2000 cells: %.1f s, %.2f M pairs/s, %d anchored reads" % (dt, rate / 1e6, total))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "singlecell_20m_test.json") if os.path.isdir(
            os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")) else os.devnull, "w") as fh:
        json.dump({"cells": n_cells, "pairs_per_cell": ppc, "seconds": dt, "pairs_per_s": rate, "anchored_reads": total,
                   "host_threads": threads}, fh)
