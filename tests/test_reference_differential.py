"""Live differential tests against the REFERENCE's own functions.py -- only where /root/reference is
mounted (the build container; the GPU box has no reference and skips this file).  The committed
goldens (tests/golden/ref_functions.json) pin a fixed set of cases; here the same comparisons run on
a few thousand freshly drawn ones."""
import importlib.util
import os
import random
import tempfile

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/functions.py"),
                                reason="the reference is not mounted here")


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_bridge import load_reference_functions
    return load_reference_functions()


@pytest.fixture(scope="module")
def gen():
    spec = importlib.util.spec_from_file_location("make_goldens", os.path.join(GOLDEN, "make_goldens.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _both(f_ref, f_mine, *args):
    """(result or exception type) of both implementations."""
    out = []
    for f in (f_ref, f_mine):
        try:
            out.append(("ok", f(*args)))
        except Exception as e:      # the reference's error behaviour is part of the contract
            out.append(("raise", type(e).__name__))
    return out


def test_deal_cigar_random_cigars(ref):
    from anchored_fusion_b200.functions import deal_cigar
    rng = random.Random(2024)
    n_checked = 0
    for _ in range(3000):
        n_ops = rng.randint(1, 6)
        ops, last = [], None
        for k in range(n_ops):
            op = rng.choice("MMMMSSIDNH")
            if op == last:
                continue
            last = op
            ops.append((rng.randint(1, 60), op))
        cigar = "".join("%d%s" % t for t in ops)
        qlen = sum(n for n, op in ops if op in "MIS")
        seq = "".join(rng.choice("ACGT") for _ in range(qlen))
        a, b = _both(ref.deal_cigar, deal_cigar, cigar, seq)
        assert a[0] == b[0], (cigar, a, b)
        if a[0] == "ok":
            assert [list(x) for x in a[1][0]] == [list(x) for x in b[1][0]] and a[1][1] == b[1][1], cigar
        else:
            assert a[1] == b[1], cigar
        n_checked += 1
    assert n_checked == 3000


def test_reverse_random_strings(ref):
    from anchored_fusion_b200.functions import reverse
    rng = random.Random(7)
    for _ in range(500):
        s = "".join(rng.choice("ACGTNHacgtX") if rng.random() < 0.02 else rng.choice("ACGTNH") for _ in range(rng.randint(0, 120)))
        a, b = _both(ref.reverse, reverse, s)
        assert a == b, s


def test_contact_reads_random_junction_sets(ref, gen, bundled):
    """Random split reads around random junctions (mismatches, both types, duplicates, near-by breakpoints):
    the merged Split_reads lists of both implementations agree field by field."""
    from anchored_fusion_b200.functions import contact_reads
    anchor = bundled["anchor"]
    rng = random.Random(99)
    for trial in range(40):
        lines, k = [], 0
        centers = [rng.randint(200, len(anchor) - 200) for _ in range(rng.randint(1, 5))]
        bps = sorted(set(c + rng.randint(-4, 4) for c in centers for _ in range(rng.randint(1, 3))))
        for bp in bps:
            for rep in range(rng.randint(1, 8)):
                typ = rng.choice(["SM", "MS"])
                m = rng.randint(20, 90)
                s = 101 - m
                partner = "".join(random.Random(bp * 7 + (typ == "SM") + trial).choice("ACGT") for _ in range(101))
                if typ == "SM":
                    seq, cg, pos = partner[-s:] + anchor[bp - 1: bp - 1 + m], "%dS%dM" % (s, m), bp
                else:
                    seq, cg, pos = anchor[bp - m: bp] + partner[:s], "%dM%dS" % (m, s), bp - m + 1
                seq = list(seq)
                for _ in range(rng.randint(0, 3)):
                    seq[rng.randrange(101)] = rng.choice("ACGTN")
                lines.append((pos, gen.pseudo_sam("t%d_%d" % (trial, k), "BCR", pos, cg, "".join(seq))))
                k += 1
        lines = [l for _, l in sorted(lines, key=lambda t: t[0])]
        want = gen.run_contact_reads(ref, lines)
        with tempfile.NamedTemporaryFile("w", suffix=".sam", delete=False) as fh:
            fh.writelines(lines)
            path = fh.name
        try:
            got = [gen.dump_split(b) for b in contact_reads(path, "", "", "1")]
        finally:
            os.remove(path)
        assert got == want, trial


def test_del_too_many_reads_random_genome_alignments(ref, gen):
    """The 2-op selection and the contiguity decision on freshly drawn anchored records and genome SAM
    texts (reference run with samtools / bwa replaced by those texts, as in make_goldens.py)."""
    from anchored_fusion_b200.functions import contiguity_filter, two_op_records
    for seed in range(25):
        rng = random.Random(1000 + seed)
        g = gen.run_del_too_many(ref, rng, seed % 6)
        fasta = "".join(">%s\n%s\n" % (tag, seq) for tag, seq in two_op_records(g["anchored"]))
        assert fasta == g["fasta"], seed
        assert "".join(contiguity_filter(g["genome_sam"])) == g["out_sam"], seed


def test_fine_block_first_loop_random_records(ref, gen):
    from anchored_fusion_b200.functions import fine_block_candidates
    rng = random.Random(5)
    for trial in range(20):
        lines = []
        for k in range(rng.randint(1, 60)):
            L = rng.choice([76, 101, 150])
            seq = "".join(rng.choice("ACGTN") for _ in range(L))
            cg = gen.random_cigar(rng, L, ["SM", "MS", "M", "SMS", "MDM", "MIM", "HM", "SM", "MS"])
            if "H" in cg:
                seq = seq[int(cg.split("H")[0]):]
            lines.append(gen.pseudo_sam("f%d_%d" % (trial, k), "BCR", rng.randint(1, 6000), cg, seq))
        want = gen.run_find_fine_block(ref, lines)
        rows = [l.split("\t") for l in lines]
        cands, fasta = fine_block_candidates((a[0], a[2], a[3], a[5], a[9]) for a in rows)
        assert [[c.type_, c.left_length, c.right_length, c.read_name] for c in cands] == want["candidates"], trial
        assert fasta == want["fasta"], trial
