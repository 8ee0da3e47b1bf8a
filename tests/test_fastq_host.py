"""Host ingest: the C++ FASTQ(.gz) reader + packer (no GPU needed) and the BAM writer."""
import gzip

import numpy as np
import pytest


def _write_fastq(path, names, seqs, quals, gz):
    opener = gzip.open if gz else open
    with opener(path, "wt") as fh:
        for n, s, q in zip(names, seqs, quals):
            fh.write("@%s\n%s\n+\n%s\n" % (n, s, q))


@pytest.mark.parametrize("longest", [100, 300])           # 300: the long-read layout (W rounded up to 20 words, 512-bit N masks)
@pytest.mark.parametrize("gz", [True, False])
def test_reader_batches_equal_pack_pairs(tmp_path, gz, longest):
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader, peek_max_read_len
    rng = np.random.default_rng(3)
    n = 2500
    s1 = ["".join(rng.choice(list("ACGTN"), p=[.245, .245, .245, .245, .02], size=rng.integers(20, longest + 1))) for _ in range(n)]
    s2 = ["".join(rng.choice(list("ACGT"), size=rng.integers(20, longest + 1))) for _ in range(n)]
    q1 = ["".join(chr(33 + int(x)) for x in rng.integers(2, 40, len(s))) for s in s1]
    q2 = ["I" * len(s) for s in s2]
    names = ["frag%d extra comment" % i for i in range(n)]
    ext = ".fastq.gz" if gz else ".fastq"
    p1, p2 = str(tmp_path / ("x_1" + ext)), str(tmp_path / ("x_2" + ext))
    _write_fastq(p1, [m + "/1" if i % 2 else m for i, m in enumerate(names)], s1, q1, gz)
    _write_fastq(p2, [("frag%d/2" % i) for i in range(n)], s2, q2, gz)
    mrl = max(peek_max_read_len(p1, 50), 112 if longest == 100 else 304)
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


def test_bam_passes_an_independent_spec_check(tmp_path):
    """The writer's twin (read_bam) proves nothing about validity: bam_spec_check unpacks the BGZF container with
    Python's gzip module and checks every field against the SAM/BAM specification."""
    from anchored_fusion_b200.bam import BamWriter
    from bam_spec_check import check_bam
    rng = np.random.default_rng(2)
    p = str(tmp_path / "s.bam")
    rows = []
    with BamWriter(p, "GENE", 200_000) as w:
        for i in range(5000):                                               # > 64 KB of records: several BGZF blocks
            L = int(rng.integers(30, 151))
            cl, cr = int(rng.integers(0, 12)), int(rng.integers(0, 12))
            ops = ([(cl, "S")] if cl else []) + [(L - cl - cr, "M")] + ([(cr, "S")] if cr else [])
            seq = "".join(rng.choice(list("ACGTN"), size=L))
            qual = "".join(chr(33 + int(x)) for x in rng.integers(0, 42, L))
            pos = 1 + 39 * i
            flag = 0x1 | (0x40 if i & 1 else 0x80) | (0x10 if i % 3 == 0 else 0)
            w.write("read%d" % i, flag, pos, 60, ops, seq, qual, next_pos=pos)
            rows.append(("read%d" % i, flag, pos, "".join("%d%s" % o for o in ops), seq, qual))
            if i % 50 == 0:
                w.write("mate%d" % i, 0x1 | 0x4 | 0x80, pos, 0, [], "ACGTA", "IIIII", next_pos=pos, mapped=False)
    res = check_bam(p)
    assert res["blocks"] >= 4 and res["refs"] == [("GENE", 200_000)]
    mapped = [r for r in res["records"] if not r["flag"] & 4]
    assert [(r["qname"], r["flag"], r["pos"], r["cigar"], r["seq"], r["qual"]) for r in mapped] == rows
    assert sum(1 for r in res["records"] if r["flag"] & 4) == 100
    empty = str(tmp_path / "e.bam")
    BamWriter(empty, "GENE", 10).close()
    assert check_bam(empty)["records"] == []


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


# ---- round 2: the task-parallel reader (BGZF, multi-member gzip, many files, threads) ---------------------
def _bgzf_bytes(data, level=6, block=0xFF00):
    import struct
    import zlib
    out = bytearray()
    for i in range(0, len(data), block):
        blk = data[i:i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(blk) + co.flush()
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp
        out += struct.pack("<II", zlib.crc32(blk) & 0xFFFFFFFF, len(blk))
    return bytes(out) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _fastq_text(names, seqs, quals):
    return "".join("@%s\n%s\n+\n%s\n" % t for t in zip(names, seqs, quals)).encode()


def _random_pairs(n, seed, lo=30, hi=151, with_n=True):
    rng = np.random.default_rng(seed)
    alpha, p = (list("ACGTN"), [.2475, .2475, .2475, .2475, .01]) if with_n else (list("ACGT"), None)
    s1 = ["".join(rng.choice(alpha, p=p, size=rng.integers(lo, hi))) for _ in range(n)]
    s2 = ["".join(rng.choice(alpha, p=p, size=rng.integers(lo, hi))) for _ in range(n)]
    q1 = ["".join(chr(33 + int(x)) for x in rng.integers(2, 41, len(s))) for s in s1]
    q2 = ["F" * len(s) for s in s2]
    return s1, s2, q1, q2


def _read_all(reader):
    """[(n_pairs, packed copy, lens copy, nids copy, first pair)] of every batch + records of all reads."""
    out, recs = [], []
    while True:
        b = reader.next_batch()
        if b is None:
            break
        lay_words = __import__("anchored_fusion_b200").layout(b.max_read_len, b.n_pairs).packed_bytes // 4
        out.append((b.n_pairs, b.packed[:lay_words].copy(), b.lens.copy(), None if b.nread_ids is None else b.nread_ids.copy(),
                    reader.first_pair))
        recs += reader.records(np.arange(2 * b.n_pairs))
    return out, recs


@pytest.mark.parametrize("fmt", ["bgzf", "multi_member", "gzip", "plain", "bgzf_then_gzip"])
@pytest.mark.parametrize("threads", [1, 3])
def test_reader_formats_and_thread_counts_give_identical_batches(tmp_path, fmt, threads):
    """The same 40 k pairs as BGZF, as concatenated gzip members, as one gzip member, as plain text and as a
    BGZF file with plain gzip members appended: identical packed batches, lengths, N lists and record text,
    for any worker count (the parallel paths cut the text at other places than the serial one)."""
    import gzip as gz
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader
    n = 40_000
    s1, s2, q1, q2 = _random_pairs(n, 5)
    names = ["r%d/1 comment" % i for i in range(n)]
    t1, t2 = _fastq_text(names, s1, q1), _fastq_text(["r%d/2" % i for i in range(n)], s2, q2)

    def enc(t):
        if fmt == "bgzf":
            return _bgzf_bytes(t)
        if fmt == "multi_member":
            return b"".join(gz.compress(t[i:i + 700_001], compresslevel=1 + (i // 700_001) % 9) for i in range(0, len(t), 700_001))
        if fmt == "gzip":
            return gz.compress(t, compresslevel=6)
        if fmt == "bgzf_then_gzip":
            cut = len(t) // 2
            cut = t.index(b"\n@r", cut) + 1
            return _bgzf_bytes(t[:cut])[:-28] + gz.compress(t[cut:])
        return t
    ext = ".fastq" if fmt == "plain" else ".fastq.gz"
    p1, p2 = str(tmp_path / ("f_1" + ext)), str(tmp_path / ("f_2" + ext))
    open(p1, "wb").write(enc(t1))
    open(p2, "wb").write(enc(t2))
    rd = FastqPairReader(p1, p2, 160, 0xE4, 16_384, threads=threads)
    assert rd.threads == threads
    batches, recs = _read_all(rd)
    rd.close()
    assert sum(b[0] for b in batches) == n and [b[4] for b in batches] == [0, 16_384, 32_768]
    assert recs == [x for i in range(n) for x in (("r%d" % i, s1[i], q1[i]), ("r%d" % i, s2[i], q2[i]))]
    lo = 0
    for npairs, packed, lens, nids, _ in batches:
        want = af.pack_pairs(s1[lo:lo + npairs], s2[lo:lo + npairs], max_read_len=160, pad_byte=0xE4)
        assert np.array_equal(packed, want.packed) and np.array_equal(lens, want.lens)
        assert (nids is None and want.n_nreads == 0) or np.array_equal(nids, want.nread_ids)
        lo += npairs


def test_reader_many_cells_as_one_stream_and_skip(tmp_path):
    """Single-cell layout: 37 small file pairs (gzip, BGZF, plain, one empty cell) read as ONE stream; batches
    span cells, file_starts() gives each cell's first pair, skip_batch() steps over a batch without packing."""
    import gzip as gz
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader
    rng = np.random.default_rng(8)
    f1, f2, all1, all2, sizes = [], [], [], [], []
    for c in range(37):
        n = 0 if c == 5 else int(rng.integers(1, 900))
        s1, s2, q1, q2 = _random_pairs(n, 100 + c, 40, 101, with_n=False)
        t1 = _fastq_text(["c%d_%d/1" % (c, i) for i in range(n)], s1, q1)
        t2 = _fastq_text(["c%d_%d/2" % (c, i) for i in range(n)], s2, q2)
        kind = c % 3
        ext = (".fastq.gz", ".fq.gz", ".fastq")[kind]
        p1, p2 = str(tmp_path / ("cell%02d_1%s" % (c, ext))), str(tmp_path / ("cell%02d_2%s" % (c, ext)))
        for p, t in ((p1, t1), (p2, t2)):
            open(p, "wb").write(gz.compress(t) if kind == 0 else _bgzf_bytes(t) if kind == 1 else t)
        f1.append(p1)
        f2.append(p2)
        all1 += s1
        all2 += s2
        sizes.append(n)
    total = sum(sizes)
    rd = FastqPairReader(f1, f2, 112, 0xE4, 4096, threads=4)
    got1, got2, i = [], [], 0
    while True:
        if i % 3 == 1:
            n = rd.skip_batch()
            if n == 0:
                break
            recs = rd.records(np.arange(2 * n))                 # text of a skipped batch is still there
        else:
            b = rd.next_batch()
            if b is None:
                break
            n = b.n_pairs
            lo = rd.first_pair
            want = af.pack_pairs(all1[lo:lo + n], all2[lo:lo + n], max_read_len=112, pad_byte=0xE4)
            assert np.array_equal(b.packed[: len(want.packed)], want.packed)
            recs = rd.records(np.arange(2 * n))
        got1 += [r[1] for r in recs[0::2]]
        got2 += [r[1] for r in recs[1::2]]
        i += 1
    assert got1 == all1 and got2 == all2 and len(got1) == total
    assert list(rd.file_starts()) == list(np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    rd.close()


def test_reader_reports_damage(tmp_path):
    """Truncated gzip (no silent partial sample), a flipped byte (CRC), a cell whose two files differ in size,
    a BGZF block cut short."""
    import gzip as gz
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader
    s1, s2, q1, q2 = _random_pairs(30_000, 3, 100, 101, with_n=False)
    names = ["r%d" % i for i in range(len(s1))]
    g1, g2 = gz.compress(_fastq_text(names, s1, q1)), gz.compress(_fastq_text(names, s2, q2))
    good1, good2 = str(tmp_path / "g_1.fastq.gz"), str(tmp_path / "g_2.fastq.gz")
    open(good1, "wb").write(g1)
    open(good2, "wb").write(g2)

    def drain(p1, p2):
        rd = FastqPairReader(p1, p2, 112, 0xE4, 8192, threads=2)
        try:
            while rd.next_batch() is not None:
                pass
        finally:
            rd.close()

    drain(good1, good2)
    bad = str(tmp_path / "bad_1.fastq.gz")
    open(bad, "wb").write(g1[: len(g1) * 2 // 3])                          # both cut at the same fraction would once pass silently
    cut2 = str(tmp_path / "bad_2.fastq.gz")
    open(cut2, "wb").write(g2[: len(g2) * 2 // 3])
    with pytest.raises(af.AnchoredFusionError, match="truncated"):
        drain(bad, cut2)
    flipped = bytearray(g1)
    flipped[len(flipped) // 2] ^= 0x10
    open(bad, "wb").write(bytes(flipped))
    with pytest.raises(af.AnchoredFusionError, match="CRC|corrupt|truncated|FASTQ"):
        drain(bad, good2)
    b1 = _bgzf_bytes(_fastq_text(names, s1, q1))
    open(bad, "wb").write(b1[: len(b1) // 2])
    with pytest.raises(af.AnchoredFusionError, match="truncated|corrupt"):
        drain(bad, good2)
    # two cells, the second one's mate file is one record short
    c1a, c1b = str(tmp_path / "ca_1.fastq"), str(tmp_path / "ca_2.fastq")
    c2a, c2b = str(tmp_path / "cb_1.fastq"), str(tmp_path / "cb_2.fastq")
    open(c1a, "wb").write(_fastq_text(names[:10], s1[:10], q1[:10]))
    open(c1b, "wb").write(_fastq_text(names[:9], s2[:9], q2[:9]))
    open(c2a, "wb").write(_fastq_text(names[:10], s1[:10], q1[:10]))
    open(c2b, "wb").write(_fastq_text(names[:11], s2[:11], q2[:11]))
    with pytest.raises(af.AnchoredFusionError, match="out of step"):
        rd = FastqPairReader([c1a, c2a], [c1b, c2b], 112, 0xE4, 8192, threads=2)
        try:
            while rd.next_batch() is not None:
                pass
        finally:
            rd.close()


def test_inflater_equals_zlib_on_many_streams(tmp_path):
    """af_inflate.h against zlib through the reader: stored / fixed / dynamic blocks, every compression level
    and strategy, tiny and empty members, long runs, binary noise (compared as 'FASTQ' whose lines are the data)."""
    import zlib
    from anchored_fusion_b200.stage import FastqPairReader
    rng = np.random.default_rng(17)
    n = 6000
    seqs = []
    for i in range(n):
        kind = i % 4
        L = int(rng.integers(1, 200))
        if kind == 0:
            seqs.append("".join(rng.choice(list("ACGT"), size=L)))
        elif kind == 1:
            seqs.append("ACGT"[i % 4] * L)                                   # runs: distance-1 matches
        elif kind == 2:
            seqs.append(("ACGTTGCA" * 30)[:L])                               # short-period repeats: overlapping copies
        else:
            seqs.append("".join(rng.choice(list("ACGTNacgtnRYKM"), size=L)))
    quals = ["".join(chr(33 + int(x)) for x in rng.integers(0, 60, len(s))) for s in seqs]
    text = _fastq_text(["n%d" % i for i in range(n)], seqs, quals)
    variants = []
    for level in (0, 1, 3, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
            for wbits in (31, 25):                                             # 32 KB and 512 B windows
                co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9 if level else 1, strategy)
                variants.append(co.compress(text) + co.flush())
    # sync-flushed stream (empty stored blocks inside) and a stream of tiny members
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    variants.append(b"".join(co.compress(text[i:i + 5000]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(text), 5000)) + co.flush())
    import gzip as gz
    variants.append(b"".join(gz.compress(text[i:i + 997]) for i in range(0, len(text), 997)) + gz.compress(b""))
    other = str(tmp_path / "o_2.fastq")
    open(other, "wb").write(text)
    for k, blob in enumerate(variants):
        p = str(tmp_path / ("v%d_1.fastq.gz" % k))
        open(p, "wb").write(blob + b"\0" * 16 if k % 7 == 3 else blob)          # trailing garbage is ignored, as gzip does
        rd = FastqPairReader(p, other, 208, 0xE4, 4096, threads=2)
        _, recs = _read_all(rd)
        rd.close()
        assert [r[1] for r in recs[0::2]] == seqs and [r[2] for r in recs[0::2]] == quals, k


def test_reader_crc32_equals_zlib_on_every_length_and_alignment():
    """af_crc32.h (PCLMULQDQ folding, zlib below 64 bytes and on CPUs without it) against zlib.crc32."""
    import zlib
    from anchored_fusion_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(9)
    sizes = list(range(0, 200)) + [255, 256, 257, 1023, 4096, 4097, 65535, 65536, 65537, (1 << 20) + 5]
    for n in sizes:
        for off in (0, 1, 3, 15):
            raw = rng.integers(0, 256, n + off, dtype=np.uint8)
            view = raw[off:]
            assert L.af_debug_crc32(view.ctypes.data, n) == zlib.crc32(view.tobytes()), (n, off)


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("chunk", [4096, 50_000])
def test_single_member_gzip_decoded_by_several_workers(tmp_path, monkeypatch, level, chunk):
    """One gzip member per file, decoded in parallel (af_inflate_par.h: block starts found by search, symbolic decode,
    chaining, marker resolution): identical batches and record text to the serial path, for streams of real gzip at
    several levels (back-references across every chunk boundary) and for two large members back to back."""
    import gzip as gz
    from anchored_fusion_b200.stage import FastqPairReader
    n = 30_000
    s1, s2, q1, q2 = _random_pairs(n, 11 + level)
    t1 = _fastq_text(["INSTR:%d:FLOW:1:%d:%d:%d 1:N:0:ACGT" % (7, 1100 + i // 1000, 1000 + i * 3, 2000 + i * 7) for i in range(n)], s1, q1)
    t2 = _fastq_text(["INSTR:%d:FLOW:1:%d:%d:%d 2:N:0:ACGT" % (7, 1100 + i // 1000, 1000 + i * 3, 2000 + i * 7) for i in range(n)], s2, q2)
    p1, p2 = str(tmp_path / "s_1.fastq.gz"), str(tmp_path / "s_2.fastq.gz")
    open(p1, "wb").write(gz.compress(t1, compresslevel=level))
    half = t2.index(b"\n@INSTR", len(t2) // 2) + 1
    open(p2, "wb").write(gz.compress(t2[:half], compresslevel=level) + gz.compress(t2[half:], compresslevel=level))   # two members
    monkeypatch.setenv("AF_GZIP_SERIAL", "1")
    rd = FastqPairReader(p1, p2, 160, 0xE4, 8192, threads=4)
    want_batches, want_recs = _read_all(rd)
    rd.close()
    monkeypatch.delenv("AF_GZIP_SERIAL")
    monkeypatch.setenv("AF_GZIP_PAR_CHUNK", str(chunk))
    for threads in (2, 5):
        rd = FastqPairReader(p1, p2, 160, 0xE4, 8192, threads=threads)
        got_batches, got_recs = _read_all(rd)
        rd.close()
        assert got_recs == want_recs and len(got_recs) == 2 * n
        assert len(got_batches) == len(want_batches)
        for g, w in zip(got_batches, want_batches):
            assert g[0] == w[0] and np.array_equal(g[1], w[1]) and np.array_equal(g[2], w[2]) and g[4] == w[4]
            assert (g[3] is None and w[3] is None) or np.array_equal(g[3], w[3])


def test_parallel_gzip_path_reports_damage(tmp_path, monkeypatch):
    """Truncation, a flipped bit in the deflate data (caught by the member CRC at the latest) and trailing garbage on the
    parallel single-member path."""
    import gzip as gz
    import anchored_fusion_b200 as af
    from anchored_fusion_b200.stage import FastqPairReader
    n = 20_000
    s1, s2, q1, q2 = _random_pairs(n, 3, with_n=False)
    t1, t2 = _fastq_text(["a%d" % i for i in range(n)], s1, q1), _fastq_text(["a%d" % i for i in range(n)], s2, q2)
    good1, good2 = gz.compress(t1, 6), gz.compress(t2, 6)
    monkeypatch.setenv("AF_GZIP_PAR_CHUNK", "30000")

    def run(b1, b2):
        p1, p2 = str(tmp_path / "d_1.fastq.gz"), str(tmp_path / "d_2.fastq.gz")
        open(p1, "wb").write(b1)
        open(p2, "wb").write(b2)
        rd = FastqPairReader(p1, p2, 160, 0xE4, 4096, threads=4)
        try:
            total = 0
            while True:
                b = rd.next_batch()
                if b is None:
                    return total
                total += b.n_pairs
        finally:
            rd.close()
    assert run(good1, good2) == n
    assert run(good1 + b"trailing garbage that is not gzip", good2) == n            # ignored, as gzip does
    with pytest.raises(af.AnchoredFusionError, match="truncated|CRC|corrupt"):
        run(good1[: len(good1) * 2 // 3], good2)
    for at in (len(good1) // 3, len(good1) // 2, len(good1) - 5):
        bad = bytearray(good1)
        bad[at] ^= 0x04
        with pytest.raises(af.AnchoredFusionError, match="CRC|corrupt|truncated|record|FASTQ"):
            run(bytes(bad), good2)


def test_chunked_gunzip_equals_zlib_and_finds_its_block_starts():
    """af_debug_gunzip_chunks: the parallel decode alone, any number of chunks, against the source text."""
    import ctypes
    import gzip as gz
    from anchored_fusion_b200._lib import check, lib
    n = 15_000
    s1, _, q1, _ = _random_pairs(n, 21)
    text = _fastq_text(["M01:%d:%d" % (i // 500, i) for i in range(n)], s1, q1)
    for level in (1, 6, 9):
        blob = np.frombuffer(gz.compress(text, level), dtype=np.uint8)
        for k in (1, 2, 7, 23):
            out = np.zeros(len(text) + 64, dtype=np.uint8)
            m, redone = ctypes.c_int64(), ctypes.c_int32()
            check(lib().af_debug_gunzip_chunks(blob.ctypes.data, len(blob), k, out.ctypes.data, len(out), ctypes.byref(m), ctypes.byref(redone)))
            assert out[: m.value].tobytes() == text and redone.value <= 1, (level, k, redone.value)
