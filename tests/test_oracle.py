"""The oracle pinned: C restatement vs the index-free pure-Python brute force, hand-checked
known answers of the spec, and the committed golden of the bundled sample."""
import numpy as np

from conftest import hits_equal


def _rc(r):
    return [3 - c if c < 4 else 4 for c in reversed(r)]


def test_c_oracle_equals_bruteforce_on_random_cases():
    from oracle import bruteforce, oracle
    rng = np.random.default_rng(1)
    a = rng.integers(0, 4, 400).astype(np.uint8)
    a[[50, 300]] = 4
    reads = []
    for t in range(240):
        L = 60
        kind = t % 5
        if kind == 0:
            r = rng.integers(0, 4, L)
        else:
            p = int(rng.integers(-20, 360))
            r = np.array([a[i] if 0 <= i < 400 and a[i] < 4 else rng.integers(0, 4) for i in range(p, p + L)])
            j = int(rng.integers(5, L - 5))
            if kind == 2:
                r[j:] = rng.integers(0, 4, L - j)
            if kind == 3:
                r[:j] = rng.integers(0, 4, j)
            for _ in range(int(rng.integers(0, 4))):
                r[rng.integers(0, L)] = rng.integers(0, 5)
            if t % 2:
                r = np.array(_rc(list(r)))
        reads.append(r.astype(np.uint8))
    hits = oracle.anchor_reads(a, np.stack(reads))
    got = {int(h["read_id"]): (int(h["pos"]), int(h["clip_l"]), int(h["m_len"]), int(h["clip_r"]),
                               int(h["score_strand"]) & 1, int(h["score_strand"]) >> 1) for h in hits}
    assert len(got) > 80
    for i, r in enumerate(reads):
        assert bruteforce.anchor_read(list(r), list(a)) == got.get(i), i


def test_spec_known_answers():
    """Hand-worked cases of 'Anchoring spec v1' (bwa-mem scores: +1/-4, clip 5, T 30, k 19)."""
    from oracle import oracle
    rng = np.random.default_rng(3)
    a = rng.integers(0, 4, 500).astype(np.uint8)

    def one(read):
        h = oracle.anchor_reads(a, np.array([read], dtype=np.uint8))
        if not len(h):
            return None
        h = h[0]
        return int(h["pos"]), int(h["clip_l"]), int(h["m_len"]), int(h["clip_r"]), int(h["score_strand"]) & 1, int(h["score_strand"]) >> 1

    exact = a[100:200].copy()
    assert one(exact) == (101, 0, 100, 0, 0, 100)                       # full match: 100M
    assert one(np.array(_rc(list(exact)))) == (101, 0, 100, 0, 1, 100)  # reverse strand, same POS/CIGAR
    assert one(a[100:129]) is None                                       # 29 < T=30
    assert one(a[100:130]) == (101, 0, 30, 0, 0, 30)
    # one mismatch 2 bases from the 5' end: local best 97 (clip 3) vs end-to-end 95 > 97-5 -> no clip
    mm = exact.copy(); mm[2] = (mm[2] + 1) % 4
    assert one(mm) == (101, 0, 100, 0, 0, 95)
    # two mismatches at 1 and 3: local best 96 (clip 4) vs end-to-end 90 <= 96-5 -> soft clip 4
    mm = exact.copy(); mm[1] = (mm[1] + 1) % 4; mm[3] = (mm[3] + 1) % 4
    assert one(mm) == (105, 4, 96, 0, 0, 96)
    mm = exact.copy(); mm[98] = (mm[98] + 1) % 4; mm[96] = (mm[96] + 1) % 4
    assert one(mm) == (101, 0, 96, 4, 0, 96)
    # a fusion junction: 60 anchor bases then foreign sequence -> 60M40S, split point POS+60-1
    junction = exact.copy(); junction[60:] = (a[160:200] + 1 + np.arange(40) % 3) % 4
    assert one(junction) == (101, 0, 60, 40, 0, 60)
    assert one(np.array(_rc(list(junction)))) == (101, 0, 60, 40, 1, 60)
    # a read hanging off the anchor's start / end is clipped at the boundary
    off = np.concatenate([(a[0:10] + 2) % 4, a[0:90]]).astype(np.uint8)
    assert one(off) == (1, 10, 90, 0, 0, 90)
    off = np.concatenate([a[440:500], (a[0:40] + 1) % 4]).astype(np.uint8)
    assert one(off) == (441, 0, 60, 40, 0, 60)
    # N is a mismatch and cannot sit inside a seed
    n = exact.copy(); n[50] = 4
    assert one(n) == (101, 0, 100, 0, 0, 95)
    n = a[100:136].copy(); n[18] = 4          # 36 bases, N in the middle: no 19-mer without N
    assert one(n) is None


def test_bundled_golden_is_what_the_oracle_computes(bundled):
    from oracle import oracle
    hits = oracle.anchor_reads(oracle.encode(bundled["anchor"]), bundled["codes"], threads=4)
    assert hits_equal(hits, bundled["oracle_hits"])
    assert len(hits) == 1261 and len(set(hits["read_id"] >> 1)) == 647
    # multi-thread == single-thread, and lens == stride is the same as lens=None
    sub = bundled["codes"][:4000]
    h1 = oracle.anchor_reads(oracle.encode(bundled["anchor"]), sub, threads=1)
    h2 = oracle.anchor_reads(oracle.encode(bundled["anchor"]), sub, lens=np.full(4000, 101, np.uint16), threads=3)
    assert hits_equal(h1, h2)


def test_bundled_sample_hits_agree_with_the_simulators_ground_truth(bundled):
    """The reference's bundled reads are wgsim simulations whose names carry the truth
    (`contig_start_end_...`).  The BCR part of EU216071.1 (the BCR-ABL1 transcript) is the anchored CDS
    in two colinear pieces (CDS position = transcript position + 452 resp. + 1175).  Every anchored read
    of the frozen oracle output sits where wgsim drew it (+-3 for its indels), and every read that lies
    inside one of the two pieces is anchored: an external pin of the positions (hence of strand and
    clipping, which decide where a read's first base lands) that does not depend on this repo's own
    arithmetic."""
    hits, names, L = bundled["oracle_hits"], bundled["names1"], bundled["read_len"]
    offsets = (452, 1175)

    def truth(name):
        parts = name.split("_")
        return int(parts[-5]), int(parts[-4])

    spans = {off: [10 ** 9, 0] for off in offsets}
    for h in hits:
        start, end = truth(names[int(h["read_id"]) >> 1])
        first = int(h["pos"]) - int(h["clip_l"])                  # CDS position of the read's first base
        where = [(base, off) for base in (start, end - L + 1) for off in offsets if abs(first - base - off) <= 3]
        assert where, (names[int(h["read_id"]) >> 1], h)          # specificity: 1 261 of 1 261
        base, off = where[0]
        if not (h["clip_l"] or h["clip_r"]):
            spans[off][0] = min(spans[off][0], base)
            spans[off][1] = max(spans[off][1], base + L - 1)
    assert spans[452][0] < 50 and spans[1175][1] > 2000
    by_id = {int(h["read_id"]): h for h in hits}
    expected = missed = 0
    for p, name in enumerate(names):
        if not name.startswith("EU216071.1"):
            assert 2 * p not in by_id and 2 * p + 1 not in by_id       # the other five contigs never anchor
            continue
        start, end = truth(name)
        for base in (start, end - L + 1):
            for off, (lo, hi) in spans.items():
                if base >= lo and base + L - 1 <= hi:
                    expected += 1
                    ok = any(abs(int(by_id[r]["pos"]) - int(by_id[r]["clip_l"]) - base - off) <= 3
                             for r in (2 * p, 2 * p + 1) if r in by_id)
                    missed += not ok
    assert expected > 1100 and missed == 0                             # sensitivity: 1 159 of 1 159


def test_c_oracle_equals_bruteforce_with_other_parameters_and_repeats():
    """Non-default scores / clip penalties / X-drop / T and an anchor made of a tandem repeat (many
    diagonals tie: score desc, strand 0 first, smaller d) -- C oracle vs the index-free brute force."""
    from oracle import bruteforce, oracle
    rng = np.random.default_rng(11)
    unit = rng.integers(0, 4, 37).astype(np.uint8)
    a = np.concatenate([rng.integers(0, 4, 120), np.tile(unit, 5), rng.integers(0, 4, 120)]).astype(np.uint8)
    G = len(a)
    for over in ({"B": 2, "X": 12, "T": 35, "clip5": 3, "clip3": 8}, {"A": 2, "B": 5, "X": 30, "T": 50, "clip5": 0, "clip3": 0},
                 {"k": 23, "T": 23, "X": 6}):
        P = dict(bruteforce.DEFAULT, **over)
        cp = oracle.default_params(**over)
        reads = []
        for t in range(150):
            L = 70
            p = int(rng.integers(-15, G - L + 15))
            r = np.array([a[i] if 0 <= i < G else rng.integers(0, 4) for i in range(p, p + L)])
            if t % 3 == 0:
                j = int(rng.integers(8, L - 8))
                r[j:] = rng.integers(0, 4, L - j)
            for _ in range(int(rng.integers(0, 5))):
                r[rng.integers(0, L)] = rng.integers(0, 5)
            if t % 2:
                r = np.array(_rc(list(r)))
            reads.append(r.astype(np.uint8))
        hits = oracle.anchor_reads(a, np.stack(reads), params=cp)
        got = {int(h["read_id"]): (int(h["pos"]), int(h["clip_l"]), int(h["m_len"]), int(h["clip_r"]),
                                   int(h["score_strand"]) & 1, int(h["score_strand"]) >> 1) for h in hits}
        assert len(got) > 60, over
        for i, r in enumerate(reads):
            assert bruteforce.anchor_read(list(r), list(a), P) == got.get(i), (over, i)


def test_cpu_arm_generator_equals_the_product_host_generator():
    """oracle/af_synth.cpp (what bench.py's reference arm and the full-size parity tests feed the oracle) and
    the product's af_synth_pairs_host are the same pure function of (seed, pair index), Ns included."""
    import anchored_fusion_b200 as af
    from oracle import oracle
    spec = af.synth_spec(seed=3, ref_len=200_000, anchor_start=50_000, anchor_len=6783, read_len=150, sub_ppm=10_000,
                         fusion_ppm=20_000, n_ppm=700)
    m1, m2 = af.synth_pairs_host(spec, 987, 3000)
    r = oracle.synth_reads(spec, 987, 3000, threads=3)
    assert (r[0::2] == m1).all() and (r[1::2] == m2).all()
    assert oracle.synth_anchor(spec) == af.synth_anchor(spec)
