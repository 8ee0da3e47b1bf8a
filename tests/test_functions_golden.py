"""Host-side mirror of the reference's record interpretation vs the REFERENCE's own outputs
(tests/golden/ref_functions.json, produced by importing /root/reference/functions.py)."""
import json
import os

import numpy as np
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
