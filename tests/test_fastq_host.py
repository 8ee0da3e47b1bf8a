"""Host ingest: the C++ FASTQ(.gz) reader + packer (no GPU needed) and the BAM writer."""
import gzip

import numpy as np
import pytest


def _write_fastq(path, names, seqs, quals, gz):
    opener = gzip.open if gz else open
    with opener(path, "wt") as fh:
        for n, s, q in zip(names, seqs, quals):
            fh.write("@%s\n%s\n+\n%s\n" % (n, s, q))


@pytest.mark.parametrize("gz", [True, False])
def test_reader_batches_equal_pack_pairs(tmp_path, gz):
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader, peek_max_read_len
    rng = np.random.default_rng(3)
    n = 2500
    s1 = ["".join(rng.choice(list("ACGTN"), p=[.245, .245, .245, .245, .02], size=rng.integers(20, 101))) for _ in range(n)]
    s2 = ["".join(rng.choice(list("ACGT"), size=rng.integers(20, 101))) for _ in range(n)]
    q1 = ["".join(chr(33 + int(x)) for x in rng.integers(2, 40, len(s))) for s in s1]
    q2 = ["I" * len(s) for s in s2]
    names = ["frag%d extra comment" % i for i in range(n)]
    ext = ".fastq.gz" if gz else ".fastq"
    p1, p2 = str(tmp_path / ("x_1" + ext)), str(tmp_path / ("x_2" + ext))
    _write_fastq(p1, [m + "/1" if i % 2 else m for i, m in enumerate(names)], s1, q1, gz)
    _write_fastq(p2, [("frag%d/2" % i) for i in range(n)], s2, q2, gz)
    mrl = max(peek_max_read_len(p1, 50), 112)
    reader = FastqPairReader(p1, p2, mrl, 0xE4, 1000)
    seen = 0
    for bi in range(4):
        b = reader.next_batch()
        if bi == 3:
            assert b is None
            break
        lo, hi = seen, min(seen + 1000, n)
        assert b.n_pairs == hi - lo
        want = af.pack_pairs(s1[lo:hi], s2[lo:hi], max_read_len=mrl, pad_byte=0xE4)
        lay = af.layout(mrl, b.n_pairs)
        assert np.array_equal(b.packed[: lay.packed_bytes // 4], want.packed)
        assert np.array_equal(b.lens, want.lens)
        assert (b.n_nreads == 0 and want.n_nreads == 0) or np.array_equal(b.nread_ids, want.nread_ids)
        for rid in (0, 1, 2 * (hi - lo) - 1):
            name, seq, qual = reader.record(rid)
            g = lo + rid // 2
            assert name == "frag%d" % g                      # comment and /1 /2 dropped, as bwa does
            assert (seq, qual) == ((s1, s2)[rid & 1][g], (q1, q2)[rid & 1][g])
        seen = hi
    reader.close()


def test_reader_errors(tmp_path):
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader
    p1, p2 = str(tmp_path / "a_1.fastq"), str(tmp_path / "a_2.fastq")
    _write_fastq(p1, ["r0", "r1"], ["ACGT", "ACGT"], ["IIII", "IIII"], False)
    _write_fastq(p2, ["r0"], ["ACGT"], ["IIII"], False)
    with pytest.raises(af.AnchoredFusionError, match="out of step"):
        FastqPairReader(p1, p2, 16, 0xE4, 10).next_batch()
    with pytest.raises(af.AnchoredFusionError, match="cannot open"):
        FastqPairReader(str(tmp_path / "missing.fastq"), p2, 16, 0xE4, 10)
    open(p2, "w").write("@r0\nACGT\n+\nIII\n@r1\nACGT\n+\nIIII\n")
    with pytest.raises(af.AnchoredFusionError, match="quality length"):
        FastqPairReader(p1, p2, 16, 0xE4, 10).next_batch()
    _write_fastq(p2, ["r0", "r1"], ["ACGT", "ACGTACGTACGTACGTACGT"], ["IIII", "I" * 20], False)
    with pytest.raises(af.AnchoredFusionError, match="max_read_len"):
        FastqPairReader(p1, p2, 16, 0xE4, 10).next_batch()


def test_reader_line_endings_blank_lines_and_unterminated_tail(tmp_path):
    """CRLF files, stray blank lines between records, a last line without '\\n', a record cut short."""
    from anchored_fusion_b200.stage import FastqPairReader
    import anchored_fusion_b200 as af
    p1, p2 = str(tmp_path / "e_1.fastq"), str(tmp_path / "e_2.fastq")
    open(p1, "wb").write(b"@r0/1 c\r\nACGTAC\r\n+\r\nIIIIII\r\n\r\n@r1\r\nGGGG\r\n+r1\r\n!!!!")      # CRLF, blank, no final newline
    open(p2, "wb").write(b"\n@r0/2\nTTTTT\n+\n#####\n\n\n@r1\nCC\n+\nII\n")
    rd = FastqPairReader(p1, p2, 16, 0xE4, 10)
    b = rd.next_batch()
    assert b.n_pairs == 2 and list(b.lens) == [6, 5, 4, 2]
    assert rd.record(0) == ("r0", "ACGTAC", "IIIIII") and rd.record(1) == ("r0", "TTTTT", "#####")
    assert rd.record(2) == ("r1", "GGGG", "!!!!") and rd.record(3) == ("r1", "CC", "II")
    assert rd.next_batch() is None
    rd.close()
    open(p2, "w").write("@r0\nTTTTT\n+\n#####\n@r1\nCC\n+\n")
    with pytest.raises(af.AnchoredFusionError, match="truncated"):
        FastqPairReader(p1, p2, 16, 0xE4, 10).next_batch()


@pytest.mark.parametrize("gz", [True, False])
def test_reader_records_across_decode_blocks_and_early_close(tmp_path, gz):
    """~30 MB of text per file: records straddle the 4 MB decode blocks, the inflate threads run ahead of
    the consumer (bounded queue), and closing a reader mid-file joins them cleanly."""
    from anchored_fusion_b200.stage import FastqPairReader
    rng = np.random.default_rng(11)
    n = 90_000
    codes = rng.integers(0, 4, (2 * n, 150)).astype(np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [lut[c].tobytes().decode() for c in codes]
    ext = ".fastq.gz" if gz else ".fastq"
    p1, p2 = str(tmp_path / ("b_1" + ext)), str(tmp_path / ("b_2" + ext))
    opener = (lambda p: gzip.open(p, "wt", compresslevel=1)) if gz else (lambda p: open(p, "w"))
    for path, mate in ((p1, 0), (p2, 1)):
        with opener(path) as fh:
            fh.write("".join("@read_number_%d/%d\n%s\n+\n%s\n" % (i, mate + 1, seqs[2 * i + mate], "F" * 150) for i in range(n)))
    rd = FastqPairReader(p1, p2, 150, 0xE4, 25_000)
    seen = 0
    while True:
        b = rd.next_batch()
        if b is None:
            break
        for rid in (0, 2 * b.n_pairs - 1, b.n_pairs):
            g = seen + rid // 2
            assert rd.record(rid) == ("read_number_%d" % g, seqs[2 * g + (rid & 1)], "F" * 150)
        seen += b.n_pairs
    assert seen == n
    rd.close()
    rd = FastqPairReader(p1, p2, 150, 0xE4, 1000)      # read a little, then close with the queues full
    assert rd.next_batch().n_pairs == 1000
    rd.close()


def test_bam_roundtrip(tmp_path):
    from anchored_fusion_b200.bam import BamWriter, read_bam, sam_line
    p = str(tmp_path / "t.bam")
    with BamWriter(p, "BCR", 6783) as w:
        w.write("r1", 0x41, 100, 60, [(40, "S"), (61, "M")], "ACGTN" * 20 + "A", "2" * 101, next_pos=300)
        w.write("r2", 0x85, 100, 0, [], "ACG", "!!I", next_pos=100, mapped=False)
        for i in range(3000):                                  # spans several BGZF blocks
            w.write("q%d" % i, 0x91, 200 + i, 60, [(101, "M")], "ACGT" * 25 + "C", "2" * 101)
    text, refs, recs = read_bam(p)
    assert refs == [("BCR", 6783)] and "SO:coordinate" in text and len(recs) == 3002
    assert sam_line(recs[0]) == "r1\t65\tBCR\t100\t60\t40S61M\t=\t300\t0\t" + "ACGTN" * 20 + "A\t" + "2" * 101 + "\n"
    assert (recs[1]["flag"], recs[1]["cigar"], recs[1]["seq"], recs[1]["qual"]) == (133, "*", "ACG", "!!I")
    assert recs[-1]["pos"] == 3199
    raw = gzip.open(p).read()                                   # every BGZF block is a gzip member
    assert raw[:4] == b"BAM\x01"
    assert open(p, "rb").read()[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_cli_gene_names_and_anchor_split(tmp_path, bundled):
    from anchored_fusion_b200.cli import discover_cells, parse_gene_names, split_anchor_fasta
    fa = tmp_path / "t.fa"
    fa.write_text("%s\n%s\n>NM_005157.6 ABL1 [organism=Homo sapiens] [GeneID=25]\nACGTACGT\nTTTT\n"
                  % (bundled["header"], bundled["anchor"]))
    assert parse_gene_names(str(fa)) == ["BCR", "ABL1"]
    gl = tmp_path / "genes.txt"
    gl.write_text("G1\n\nG2\n")
    assert parse_gene_names(str(fa), str(gl)) == ["G1", "G2"]
    outs = split_anchor_fasta(str(fa), ["BCR", "ABL1"], lambda g: str(tmp_path / (g + ".fa")))
    assert open(outs[0]).read() == ">BCR\n" + bundled["anchor"] + "\n"
    assert open(outs[1]).read() == ">ABL1\nACGTACGT\nTTTT\n"
    d = tmp_path / "cells"
    d.mkdir()
    for f in ("c1_1.fastq.gz", "c1_2.fastq.gz", "c2_1.fq", "c2_2.fq", "c3_1.fastq", "junk.txt", "c4_1.fq.gz", "c4_2.fq.gz"):
        (d / f).write_text("")
    assert discover_cells(str(d)) == [("c1", "c1_1.fastq.gz", "c1_2.fastq.gz"), ("c2", "c2_1.fq", "c2_2.fq"),
                                      ("c4", "c4_1.fq.gz", "c4_2.fq.gz")]
