"""Host-side mirror of the reference's record interpretation vs the REFERENCE's own outputs
(tests/golden/ref_functions.json, produced by importing /root/reference/functions.py)."""
import json
import os

import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "ref_functions.json")) as fh:
        return json.load(fh)


def _dump(groups):
    return [{"chrom": g.chrom, "breakpoint": int(g.breakpoint), "type": g.type_, "cnt": int(g.cnt),
             "reads": list(g.reads), "seq_left": g.seq_left, "seq_right": g.seq_right} for g in groups]


def test_deal_cigar_matches_reference(golden):
    from anchored_fusion_b200.functions import deal_cigar
    for case in golden["deal_cigar"]:
        ops, seq = deal_cigar(case["cigar"], case["seq"])
        assert [list(o) for o in ops] == case["ops"], case["cigar"]
        assert seq == case["seq_out"], case["cigar"]


def test_deal_cigar_survey_known_answers():
    from anchored_fusion_b200.functions import deal_cigar
    assert deal_cigar("40S61M", "A" * 101)[0] == [[40, 40, "S"], [101, 61, "M"]]
    assert deal_cigar("20H81M", "A" * 81)[0] == [[81, 81, "M"]]
    assert len(deal_cigar("5S90M6S", "A" * 101)[0]) == 3
    ops, seq = deal_cigar("50M2D41M10S", "A" * 101)
    assert ops == [[93, 93, "M"], [103, 10, "S"]] and seq[50:52] == "NN" and len(seq) == 103


def test_reverse_matches_reference(golden):
    from anchored_fusion_b200.functions import reverse
    for case in golden["reverse"]:
        assert reverse(case["seq"]) == case["out"]
    with pytest.raises(KeyError):
        reverse("ACGU")


@pytest.mark.parametrize("name", ["hand", "synthetic_junctions"])
def test_contact_reads_matches_reference(golden, tmp_path, name):
    from anchored_fusion_b200.functions import contact_reads
    case = next(c for c in golden["contact_reads"] if c["name"] == name)
    p = tmp_path / "anchored_reads.sam"
    p.write_text("".join(case["lines"]))
    assert _dump(contact_reads(str(p), "", "", "1")) == case["out"]


def test_contact_reads_on_bundled_sample_matches_reference(golden, bundled, tmp_path):
    """Oracle records of the bundled sample -> pseudo-SAM in the product's order -> our
    contact_reads == the reference's contact_reads on the same lines (split points, types,
    read sets, consensus sequences)."""
    from anchored_fusion_b200.functions import contact_reads
    from anchored_fusion_b200.records import pseudo_sam_lines, read_names
    case = next(c for c in golden["contact_reads"] if c["name"] == "bundled_c1")
    names = read_names(bundled["names1"])
    lines = pseudo_sam_lines(bundled["oracle_hits"], "BCR", names, bundled["seqs1"], bundled["seqs2"])
    assert len(lines) == case["n_lines"]
    p = tmp_path / "anchored_reads.sam"
    p.write_text("".join(lines))
    got = _dump(contact_reads(str(p), "", "", "1"))
    assert got == case["out"]
    # the survey's indicative landmark: the dominant split point is BCR:3235, type MS
    top = max(got, key=lambda g: g["cnt"])
    assert (top["breakpoint"], top["type"]) == (3235, "MS")


def test_two_op_records(golden):
    from anchored_fusion_b200.functions import two_op_records
    case = next(c for c in golden["contact_reads"] if c["name"] == "hand")
    heads = [h for h, _ in two_op_records(case["lines"])]
    assert "r1$BCR$100$40S61M" in heads and "r3$BCR$100$10S91M" in heads
    assert not any(h.startswith("r8$") for h in heads)      # 101M has one op


def test_fine_block_first_loop_matches_reference(golden, tmp_path):
    """the spanning_anchored objects and the FASTA that Find_fine_block's first loop
    (functions.py:513-529) makes of the pseudo-SAM records -- reference run with blat stubbed."""
    from anchored_fusion_b200.functions import fine_block_candidates, fine_block_candidates_from_file
    case = golden["find_fine_block_first_loop"]
    rows = [l.split("\t") for l in case["lines"]]
    cands, fasta = fine_block_candidates((a[0], a[2], a[3], a[5], a[9]) for a in rows)
    assert [[c.type_, c.left_length, c.right_length, c.read_name] for c in cands] == case["candidates"]
    assert fasta == case["fasta"]
    p = tmp_path / "split.sam"
    p.write_text("".join(case["lines"]))
    again = fine_block_candidates_from_file(str(p), str(tmp_path / "w"))
    assert len(again) == len(cands) and (tmp_path / "w_prb_spanning.fa").read_text() == case["fasta"]


@pytest.mark.parametrize("case", range(6))
def test_del_too_many_reads_matches_reference(golden, tmp_path, case):
    """SURVEY 8(a8) + 8(f)-3: the 2-op selection handed to the genome aligner and the contiguity
    decision on a prepared genome SAM text (flags 0/16/256/2048, H/S clips, D/I/N, unmapped, header
    lines first and last) -- reference run with samtools / bwa replaced by the same texts."""
    from anchored_fusion_b200.functions import contiguity_filter, del_too_many_reads, two_op_records
    g = golden["del_too_many_reads"][case]
    fasta = "".join(">%s\n%s\n" % (tag, seq) for tag, seq in two_op_records(g["anchored"]))
    assert fasta == g["fasta"]
    assert "".join(contiguity_filter(g["genome_sam"])) == g["out_sam"]
    # the whole function, same signature, files and clean-up
    f_read = tmp_path / "anchored.sam"
    f_read.write_text("@HD\tVN:1.6\n" + "\n".join(g["anchored"]) + "\n")
    out_sam = tmp_path / "out.sam"
    del_too_many_reads(str(f_read), str(out_sam), str(tmp_path / "w"), "genome.fa", "1", genome_sam=g["genome_sam"])
    assert out_sam.read_text() == g["out_sam"]
    assert not (tmp_path / "w_del_tmp.fa").exists() and not (tmp_path / "w_del_tmp.sam").exists()


def test_del_too_many_reads_needs_an_aligner(tmp_path, monkeypatch):
    from anchored_fusion_b200.functions import del_too_many_reads
    monkeypatch.setenv("PATH", str(tmp_path))
    f_read = tmp_path / "anchored.sam"
    f_read.write_text("r\t0\tBCR\t10\t60\t40S61M\t=\t1\t0\t" + "A" * 101 + "\t*\n")
    with pytest.raises(RuntimeError):
        del_too_many_reads(str(f_read), str(tmp_path / "o.sam"), str(tmp_path / "w"), "genome.fa", "1")
    assert not (tmp_path / "w_del_tmp.fa").exists()
