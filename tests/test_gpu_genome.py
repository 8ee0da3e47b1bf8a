"""Genome pass of the contiguity filter (replaces `bwa mem` on the genome inside
del_too_many_reads, /root/reference functions.py:716) on the B200, bit-exact against the CPU oracle run with
the concatenated genome as its anchor."""
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEP = 256


def _rand_seq(rng, n):
    return "".join("ACGT"[c] for c in rng.integers(0, 4, n))


def _revcomp(s):
    return s.translate(str.maketrans("ACGTN", "TGCAN"))[::-1]


def _mutate(rng, s, rate):
    b = list(s)
    for i in range(len(b)):
        if rng.random() < rate:
            b[i] = "ACGT"[(("ACGT".index(b[i]) if b[i] in "ACGT" else 0) + 1 + int(rng.integers(0, 3))) % 4]
    return "".join(b)


def _toy_genome(rng):
    """Three contigs with the things an assembly has: N runs, soft-masked (lower-case) stretches, a repeat family,
    a poly-A tract, a tandem repeat."""
    element = _rand_seq(rng, 300)
    contigs = []
    for name, n in (("chrA", 1_500_000), ("chrB", 700_000), ("chrC description ignored", 250_000)):
        s = list(_rand_seq(rng, n))
        for _ in range(20):                                   # repeat family, 4 % diverged copies, both strands
            at = int(rng.integers(1000, n - 1000))
            copy = _mutate(rng, element, 0.04)
            if rng.random() < 0.5:
                copy = _revcomp(copy)
            s[at:at + 300] = copy
        for _ in range(3):                                    # N runs
            at, ln = int(rng.integers(1000, n - 5000)), int(rng.integers(1, 3000))
            s[at:at + ln] = "N" * ln
        at = int(rng.integers(1000, n - 1000))
        s[at:at + 60] = "A" * 60                              # poly-A
        at = int(rng.integers(1000, n - 1000))
        s[at:at + 120] = "CA" * 60                            # tandem repeat
        at = int(rng.integers(1000, n - 20000))
        s[at:at + 10000] = "".join(s[at:at + 10000]).lower()  # soft-masked
        contigs.append((name, "".join(s)))
    return contigs, element


def _concat(contigs):
    """The sequence the library aligns to: [256 N] contig [256 N] contig ... [256 N]."""
    return "N" * SEP + "".join(seq + "N" * SEP for _, seq in contigs)


def _reads(rng, contigs, element):
    reads = []
    for _ in range(260):                                      # plain reads, both strands, 1.5 % substitutions
        name, seq = contigs[int(rng.integers(0, len(contigs)))]
        L = int(rng.choice([36, 76, 101, 150, 150, 150, 250, 256]))
        at = int(rng.integers(0, len(seq) - L))
        r = _mutate(rng, seq[at:at + L].upper(), 0.015)
        reads.append(_revcomp(r) if rng.random() < 0.5 else r)
    for _ in range(80):                                       # chimeras: two loci, or a locus + random sequence
        (_, s1), (_, s2) = contigs[int(rng.integers(0, 3))], contigs[int(rng.integers(0, 3))]
        cut = int(rng.integers(20, 130))
        a1, a2 = int(rng.integers(0, len(s1) - 150)), int(rng.integers(0, len(s2) - 150))
        right = s2[a2:a2 + 150 - cut].upper() if rng.random() < 0.6 else _rand_seq(rng, 150 - cut)
        r = s1[a1:a1 + cut].upper() + right
        reads.append(_revcomp(r) if rng.random() < 0.5 else r)
    for _ in range(40):                                       # reads of the repeat family: many seeded diagonals
        at = int(rng.integers(0, 150))
        reads.append(_mutate(rng, element[at:at + 150], 0.02))
    for _ in range(30):                                       # reads with N
        _, seq = contigs[0]
        at = int(rng.integers(0, len(seq) - 150))
        r = list(seq[at:at + 150].upper())
        for _ in range(int(rng.integers(1, 5))):
            r[int(rng.integers(0, 150))] = "N"
        reads.append("".join(r))
    for name, seq in contigs:                                 # reads hanging over the ends of a contig
        reads.append(_rand_seq(rng, 50) + seq[:100].upper())
        reads.append(seq[-100:].upper() + _rand_seq(rng, 50))
        reads.append(seq[-149:].upper() + "C")
        reads.append("G" + seq[:149].upper())
    reads += [_rand_seq(rng, 150) for _ in range(40)]         # noise
    reads += ["A" * 150, "A" * 40 + _rand_seq(rng, 110), "CA" * 75, "ACGT" * 30, "", "ACGTACGT", _rand_seq(rng, 18), _rand_seq(rng, 19)]
    _, seq = contigs[1]
    for L in (19, 20, 23, 30, 31, 32, 33, 64, 65):            # short exact reads around the seed length / word edges
        at = int(rng.integers(0, len(seq) - L))
        reads.append(seq[at:at + L].upper())
    return reads


def _oracle_hits(concat, reads, params=None):
    from oracle import oracle
    codes = np.full((len(reads), 256), 4, dtype=np.uint8)
    lens = np.zeros(len(reads), dtype=np.uint16)
    for i, r in enumerate(reads):
        codes[i, :len(r)] = oracle.encode(r)
        lens[i] = len(r)
    return oracle.anchor_reads(oracle.encode(concat), codes, lens=lens, params=params, threads=8)


def _same(gpu, ora):
    assert len(gpu) == len(ora), (len(gpu), len(ora))
    for f in ("read_id", "pos", "clip_l", "m_len", "clip_r", "score_strand"):
        assert np.array_equal(gpu[f].astype(np.int64), ora[f].astype(np.int64)), f


@pytest.fixture(scope="module")
def toy():
    rng = np.random.default_rng(20260219)
    contigs, element = _toy_genome(rng)
    reads = _reads(rng, contigs, element)
    concat = _concat(contigs)
    return {"contigs": contigs, "reads": reads, "concat": concat, "oracle": _oracle_hits(concat, reads)}


def test_genome_records_equal_the_oracle(toy):
    from anchored_fusion_b200.genome import Genome
    g = Genome.from_contigs(toy["contigs"])
    assert g.length == len(toy["concat"])
    assert [(c[0], c[2]) for c in g.contigs] == [(n, len(s)) for n, s in toy["contigs"]]
    hits = g.align(toy["reads"])
    assert len(toy["oracle"]) > 350                           # the case is not vacuous
    assert len(set(toy["oracle"]["score_strand"] & 1)) == 2
    _same(hits, toy["oracle"])
    st = g.last_stats
    assert st["n_passes"] >= 4 and st["n_seeds"] >= len(hits) and st["scan_ms"] > 0
    # as SAM every record lies inside one contig (an extension that ran one mismatch into the separator is clipped)
    import re
    lens = dict((c[0], c[2]) for c in g.contigs)
    lines = g.sam_lines(["r%d" % i for i in range(len(toy["reads"]))], toy["reads"], hits)
    assert len(lines) == len(toy["reads"]) and sum(1 for l in lines if l.split("\t")[1] != "4") == len(hits)
    for l in lines:
        a = l.split("\t")
        if a[1] == "4":
            continue
        ops = re.findall(r"(\d+)([SM])", a[5])
        assert sum(int(n) for n, _ in ops) == len(a[9]) and [o for _, o in ops].count("M") == 1
        m = [int(n) for n, o in ops if o == "M"][0]
        assert 1 <= int(a[3]) and int(a[3]) + m - 1 <= lens[a[2]]
    g.close()


def test_pass_size_and_buffer_growth_do_not_change_the_records(toy, monkeypatch):
    from anchored_fusion_b200.genome import Genome
    g = Genome.from_contigs(toy["contigs"])
    _same(g.align(toy["reads"], reads_per_pass=7), toy["oracle"])
    assert g.last_stats["n_passes"] == (len(toy["reads"]) + 6) // 7
    g.close()
    monkeypatch.setenv("AF_GENOME_TEST_CAP", "64")           # candidate and seed buffers start at 64 entries
    g = Genome.from_contigs(toy["contigs"])
    _same(g.align(toy["reads"]), toy["oracle"])
    assert g.last_stats["n_retries"] >= 1
    g.close()


def test_other_scores_and_thresholds(toy):
    from anchored_fusion_b200 import default_params
    from anchored_fusion_b200.genome import Genome
    from oracle import oracle
    g = Genome.from_contigs(toy["contigs"])
    sub = toy["reads"][:200] + toy["reads"][-40:]
    for over in ({"B": 2, "X": 12, "T": 40}, {"A": 2, "B": 7, "clip5": 0, "clip3": 11, "T": 50}):
        _same(g.align(sub, params=default_params(**over)), _oracle_hits(toy["concat"], sub, oracle.default_params(**over)))
    g.close()


def test_fasta_loader_plain_gzip_crlf(toy, tmp_path):
    from anchored_fusion_b200.genome import Genome
    plain, gz = tmp_path / "g.fa", tmp_path / "g.fa.gz"
    text = []
    for i, (name, seq) in enumerate(toy["contigs"]):
        eol = "\r\n" if i == 1 else "\n"
        text.append(">" + name + eol)
        width = (60, 70, 1000)[i]
        text += [seq[k:k + width] + eol for k in range(0, len(seq), width)]
    text = "".join(text)
    plain.write_text(text[:-1] if text.endswith("\n") else text)   # no newline at the end of the file
    with gzip.open(gz, "wt") as fh:
        fh.write(text)
    sub = toy["reads"][:120]
    want = toy["oracle"][toy["oracle"]["read_id"] < 120]
    for path in (plain, gz):
        g = Genome.from_fasta(path)
        assert [(c[0], c[2]) for c in g.contigs] == [(n.split()[0], len(s)) for n, s in toy["contigs"]]
        _same(g.align(sub), want)
        g.close()
    from anchored_fusion_b200 import AnchoredFusionError
    with pytest.raises(AnchoredFusionError):
        Genome.from_fasta(tmp_path / "missing.fa")
    (tmp_path / "bad.fa").write_text("ACGT\n")
    with pytest.raises(AnchoredFusionError, match="header"):
        Genome.from_fasta(tmp_path / "bad.fa")


def test_synthetic_genome_matches_the_host_generator():
    """af_genome_synth (device-generated measurement genome) holds the bases the seeded host generator gives."""
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.genome import Genome
    n = 3_000_000
    g = Genome.synthetic(7, n)
    ref = af.synth_anchor(af.synth_spec(seed=7, ref_len=n, anchor_start=0, anchor_len=n))
    if isinstance(ref, bytes):
        ref = ref.decode()
    rng = np.random.default_rng(5)
    reads = []
    for _ in range(150):
        at = int(rng.integers(0, n - 150))
        r = _mutate(rng, ref[at:at + 150], 0.01)
        reads.append(_revcomp(r) if rng.random() < 0.5 else r)
    reads += [ref[:150], ref[-150:], _rand_seq(rng, 150)]
    hits = g.align(reads)
    _same(hits, _oracle_hits("N" * SEP + ref + "N" * SEP, reads))
    assert len(hits) >= 150
    g.close()


def test_del_too_many_reads_with_the_gpu_genome_pass(toy, tmp_path):
    """del_too_many_reads end to end: SAM text of anchored reads -> 2-op reads -> genome pass on the GPU ->
    contiguity decision, against the same decision taken on SAM lines built from the oracle's records."""
    from anchored_fusion_b200.functions import contiguity_filter, del_too_many_reads, two_op_records
    from anchored_fusion_b200.genome import Genome
    fa = tmp_path / "genome.fa"
    fa.write_text("".join(">%s\n%s\n" % (n, s) for n, s in toy["contigs"]))
    rng = np.random.default_rng(3)
    sam = []
    usable = [r for r in toy["reads"] if len(r) >= 60 and set(r) <= set("ACGTN")]
    for i, r in enumerate(usable):
        cut = int(rng.integers(20, len(r) - 20))
        cigar = "%dS%dM" % (cut, len(r) - cut) if i % 2 else "%dM%dS" % (cut, len(r) - cut)
        if i % 7 == 0:
            cigar = "%dM" % len(r)                            # 1-op: not a candidate
        sam.append("\t".join(["r%d" % i, "0", "GENE", str(100 + i), "60", cigar, "*", "0", "0", r, "I" * len(r)]))
    f_read = tmp_path / "anchored.sam"
    f_read.write_text("\n".join(sam) + "\n")
    out = tmp_path / "out.sam"
    del_too_many_reads(str(f_read), str(out), str(tmp_path / "w"), str(fa), "4")
    recs = list(two_op_records(sam))
    g = Genome.from_contigs(toy["contigs"])
    ora = _oracle_hits(toy["concat"], [s for _, s in recs])
    want = contiguity_filter(g.sam_lines([t for t, _ in recs], [s for _, s in recs], ora))
    got = out.read_text().splitlines(keepends=True)
    assert got == want
    assert 0 < len(got) < len(recs)                           # some reads are explained by the genome, some are not
    assert not os.path.exists(str(tmp_path / "w") + "_del_tmp.fa")
    g.close()


def test_bulk_cli_runs_the_contiguity_stage_when_given_a_genome(bundled, tmp_path):
    """--file_ref_seq <genome FASTA>: the CLI goes on to del_too_many_reads (Anchored_Fusion.py:202-203) with the GPU
    genome pass and leaves <w>_anchored_reads.sam, the file the reference's next stages read."""
    from anchored_fusion_b200.bam import read_bam, sam_line
    from anchored_fusion_b200.cli import main_bulk
    from anchored_fusion_b200.functions import contiguity_filter, two_op_records
    from anchored_fusion_b200.genome import Genome
    d = str(tmp_path)
    fa = os.path.join(d, "target_gene.fasta")
    with open(fa, "w") as fh:
        fh.write(bundled["header"] + "\n" + bundled["anchor"] + "\n")
    p1, p2 = os.path.join(d, "s_1.fastq"), os.path.join(d, "s_2.fastq")
    q = bundled["qual_char"] * bundled["read_len"]
    with open(p1, "w") as f1, open(p2, "w") as f2:
        for i in range(len(bundled["seqs1"])):
            f1.write("@%s\n%s\n+\n%s\n" % (bundled["names1"][i], bundled["seqs1"][i], q))
            f2.write("@%s\n%s\n+\n%s\n" % (bundled["names2"][i], bundled["seqs2"][i], q))
    # toy genome: the gene itself, and a contig that holds every third junction-spanning read in full (those reads
    # are explained by the genome without a junction and must go), in random flanks
    rng = np.random.default_rng(1)
    hits = bundled["oracle_hits"]
    clipped = [h for h in hits if int(h["clip_l"]) >= 20 or int(h["clip_r"]) >= 20]
    whole = []
    for h in clipped[::3]:
        rid = int(h["read_id"])
        whole.append((bundled["seqs1"], bundled["seqs2"])[rid & 1][rid >> 1])
    contigs = [("chrG", _rand_seq(rng, 5000) + bundled["anchor"] + _rand_seq(rng, 5000)),
               ("chrJ", _rand_seq(rng, 300).join(whole) + _rand_seq(rng, 300))]
    gfa = os.path.join(d, "genome.fa")
    with open(gfa, "w") as fh:
        for n, s in contigs:
            fh.write(">%s\n" % n + "".join(s[i:i + 80] + "\n" for i in range(0, len(s), 80)))
    out = os.path.join(d, "out")
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", p1, "--fastq2", p2, "--out_folder", out, "--file_ref_seq", gfa,
                      "--not_filter_false_positive", "--thread", "4"]) == 0
    w = os.path.join(out, "BCR_fusion", "work_dir", "BCR_fusion")
    got = open(w + "_anchored_reads.sam").read().splitlines(keepends=True)
    recs = list(two_op_records([sam_line(r) for r in read_bam(w + "_anchored_reads.bam")[2]]))
    g = Genome.from_contigs(contigs)
    ora = _oracle_hits(_concat(contigs), [s for _, s in recs])
    want = contiguity_filter(g.sam_lines([t for t, _ in recs], [s for _, s in recs], ora))
    assert got == want
    assert 0 < len(got) < len(recs) and len(recs) - len(got) >= len(whole) // 2
    assert all(len(l.split("\t")) == 11 for l in got)
    assert os.path.exists(w + "_split_points_filtered.txt")
    g.close()


def test_singlecell_cli_runs_the_contiguity_stage_per_cell(bundled, tmp_path):
    """Single-cell layout with --file_ref_seq: every cell gets its <w>_anchored_reads.sam, and the union of the cells'
    survivors is what the bulk run on the concatenation keeps (the filter judges each read on its own)."""
    import gzip
    from anchored_fusion_b200.cli import main_bulk, main_singlecell
    d = str(tmp_path)
    fa = os.path.join(d, "t.fa")
    open(fa, "w").write(bundled["header"] + "\n" + bundled["anchor"] + "\n")
    rng = np.random.default_rng(2)
    gfa = os.path.join(d, "genome.fa")
    hits = bundled["oracle_hits"]
    whole = []
    for h in [h for h in hits if int(h["clip_l"]) >= 20 or int(h["clip_r"]) >= 20][::4]:
        rid = int(h["read_id"])
        whole.append((bundled["seqs1"], bundled["seqs2"])[rid & 1][rid >> 1])
    with open(gfa, "w") as fh:
        fh.write(">chrG\n" + _rand_seq(rng, 3000) + bundled["anchor"] + _rand_seq(rng, 3000) + "\n>chrJ\n" + _rand_seq(rng, 200).join(whole) + "\n")
    q = bundled["qual_char"] * bundled["read_len"]
    n = len(bundled["seqs1"])

    def write(prefix, sel):
        with gzip.open(prefix + "_1.fastq.gz", "wt") as f1, gzip.open(prefix + "_2.fastq.gz", "wt") as f2:
            for i in sel:
                f1.write("@%s\n%s\n+\n%s\n" % (bundled["names1"][i], bundled["seqs1"][i], q))
                f2.write("@%s\n%s\n+\n%s\n" % (bundled["names2"][i], bundled["seqs2"][i], q))
    cells = os.path.join(d, "cells")
    os.mkdir(cells)
    parts = {"cellA": range(0, n, 3), "cellB": range(1, n, 3), "cellC": range(2, n, 3)}
    for c, sel in parts.items():
        write(os.path.join(cells, c), sel)
    write(os.path.join(d, "all"), range(n))
    out_sc, out_bulk = os.path.join(d, "sc"), os.path.join(d, "bulk")
    assert main_singlecell(["--file_anchored_cds", fa, "--fastq_dir", cells, "--out_folder", out_sc, "--file_ref_seq", gfa, "--thread", "4"]) == 0
    assert main_bulk(["--file_anchored_cds", fa, "--fastq1", os.path.join(d, "all_1.fastq.gz"), "--fastq2", os.path.join(d, "all_2.fastq.gz"),
                      "--out_folder", out_bulk, "--file_ref_seq", gfa, "--thread", "4"]) == 0
    bulk = open(os.path.join(out_bulk, "BCR_fusion", "work_dir", "BCR_fusion_anchored_reads.sam")).read().splitlines()
    union = []
    for c in parts:
        w = os.path.join(out_sc, "BCR", "work_dir", c, "BCR_fusion")
        assert os.path.exists(w + "_split_points_filtered.txt")
        union += open(w + "_anchored_reads.sam").read().splitlines()
    assert sorted(union) == sorted(bulk) and len(bulk) > 30
    assert os.path.exists(gfa + ".af2bit")                         # the packed genome was cached beside the FASTA
