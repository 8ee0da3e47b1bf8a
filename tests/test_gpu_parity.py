"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.
Bit-exact (integer records), on the bundled sample, on seeded synthetic pairs and on edge cases."""
import numpy as np
import pytest

from conftest import hits_equal

pytestmark = pytest.mark.gpu


def _interleave(m1, m2):
    r = np.empty((2 * m1.shape[0], m1.shape[1]), dtype=np.uint8)
    r[0::2], r[1::2] = m1, m2
    return r


@pytest.fixture(scope="module")
def af():
    import torch
    assert torch.cuda.is_available()
    import anchored_fusion_b200 as af
    return af


@pytest.mark.parametrize("kp", [12, 13])
def test_bundled_sample_matches_oracle(af, bundled, kp):
    """config 1: test/target_gene.fasta + test/test_sample_{1,2}.fastq.gz (11 258 pairs, 2x101)."""
    from oracle import oracle
    index = af.AnchorIndex(bundled["anchor"], kp=kp)
    eng = af.Anchorer(index, 0)
    host = af.pack_pairs(bundled["seqs1"], bundled["seqs2"], pad_byte=index.pad_byte)
    hits, stats = eng.anchor(host.to_device(0))
    assert hits_equal(hits, bundled["oracle_hits"])           # committed golden of the frozen oracle
    live = oracle.anchor_reads(oracle.encode(bundled["anchor"]), bundled["codes"], threads=4)
    assert hits_equal(hits, live)
    assert stats["flagged"] >= len(hits)
    # wgsim names carry the truth: only EU216071.1 (BCR-ABL1) fragments can anchor to BCR
    for rid in hits["read_id"]:
        assert bundled["names1"][rid >> 1].startswith("EU216071.1")


@pytest.mark.parametrize("kp,mode,threads", [(12, 3, 768), (13, 3, 640), (12, 0, 256), (13, 3, 64)])
def test_seed_scan_flags_equal_the_filter_emulation(af, bundled, kp, mode, threads):
    """The scan kernel alone: its flag words equal a numpy emulation of the filter probes bit for
    bit (so false positives are exactly the filter's, and every oracle hit is flagged)."""
    from filter_emulator import expected_flags
    from anchored_fusion_b200._lib import check, lib
    index = af.AnchorIndex(bundled["anchor"], kp=kp)
    eng = af.Anchorer(index, 0)
    host = af.pack_pairs(bundled["seqs1"], bundled["seqs2"], pad_byte=index.pad_byte)
    check(lib().af_seed_scan_config(threads, mode))
    try:
        flags = eng.seed_scan(host.to_device(0)).cpu().numpy().view(np.uint32)
    finally:
        check(lib().af_seed_scan_config(0, 0))
    n = host.n_pairs
    got = np.zeros(2 * n, bool)
    for mate in (0, 1):
        bits = (flags[:, mate][:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1
        got[mate::2] = bits.reshape(-1)[:n].astype(bool)
    want = expected_flags(index, bundled["codes"])
    assert np.array_equal(got, want)
    assert got[bundled["oracle_hits"]["read_id"]].all()


def test_device_generator_matches_host_generator(af):
    spec = af.synth_spec(seed=11, ref_len=100_000, anchor_start=20_000, anchor_len=5000, read_len=150,
                         sub_ppm=20_000, fusion_ppm=50_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    n = 1000 + 7
    dev = af.synth_pairs_device(spec, 123, n, index.pad_byte, 0)
    m1, m2 = af.synth_pairs_host(spec, 123, n)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    host = af.pack_pairs([lut[r].tobytes() for r in m1], [lut[r].tobytes() for r in m2], pad_byte=index.pad_byte)
    assert np.array_equal(dev.packed.cpu().numpy().view(np.uint32), host.packed)


@pytest.mark.parametrize("read_len,kp,n", [(150, 12, 200_000), (101, 12, 50_000), (150, 13, 50_000),
                                           (250, 12, 20_000), (76, 13, 20_000), (36, 12, 5_000)])
def test_synthetic_pairs_match_oracle(af, read_len, kp, n):
    from oracle import oracle
    spec = af.synth_spec(seed=read_len + kp, ref_len=300_000, anchor_start=100_000, anchor_len=10_000,
                         read_len=read_len, frag_mean=max(2 * read_len, 300), sub_ppm=15_000, fusion_ppm=30_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor, kp=kp)
    eng = af.Anchorer(index, 0)
    hits, stats = eng.anchor(af.synth_pairs_device(spec, 0, n, index.pad_byte, 0))
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8)
    assert len(want) > 50
    assert hits_equal(hits, want)


@pytest.mark.parametrize("mode", [11, 8])
def test_other_middle_stages_give_the_same_records(af, mode):
    """af_seed_scan_config(0, 11): the second look at flagged reads happens inside the scan (refine queue) and
    the survivors go straight to k_extend; (0, 8): every flagged read goes to k_extend.  Default is 7,
    k_verify_smem between scan and extension."""
    from oracle import oracle
    from anchored_fusion_b200._lib import check, lib
    spec = af.synth_spec(seed=77, ref_len=300_000, anchor_start=100_000, anchor_len=8000, read_len=150,
                         sub_ppm=15_000, fusion_ppm=30_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    n = 150_000
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8)
    dev = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    ref, st_ref = eng.anchor(dev)
    check(lib().af_seed_scan_config(0, mode))
    try:
        hits, st = eng.anchor(dev)
    finally:
        check(lib().af_seed_scan_config(0, 7))
    assert hits_equal(ref, want) and hits_equal(hits, want)
    assert st["flagged"] == st_ref["flagged"]            # both count the reads that pass the plain filter


def test_variable_length_reads_with_n(af):
    """ragged lengths (19..150), Ns (mismatch; no seed across an N), reads off both anchor ends."""
    from oracle import oracle
    rng = np.random.default_rng(5)
    G = 3000
    anchor = rng.integers(0, 4, G).astype(np.uint8)
    anchor[[100, 101, 1500]] = 4                       # N in the anchor too
    n, stride = 4000, 150
    reads = np.full((2 * n, stride), 4, dtype=np.uint8)
    lens = rng.integers(19, stride + 1, 2 * n).astype(np.uint16)
    for i in range(2 * n):
        L = int(lens[i])
        kind = i % 4
        if kind == 0:
            r = rng.integers(0, 4, L)
        else:
            p = int(rng.integers(-40, G - L + 40))
            r = np.array([anchor[j] if 0 <= j < G and anchor[j] < 4 else rng.integers(0, 4) for j in range(p, p + L)])
            if kind == 2:
                j = int(rng.integers(0, L))
                r[j:] = rng.integers(0, 4, L - j)
            for _ in range(int(rng.integers(0, 4))):
                r[rng.integers(0, L)] = rng.integers(0, 5)
            if i % 8 >= 4:
                r = np.array([3 - c if c < 4 else 4 for c in r[::-1]])
        reads[i, :L] = r
    want = oracle.anchor_reads(anchor, reads, lens=lens)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    index = af.AnchorIndex(lut[anchor].tobytes())
    seqs = [lut[reads[i, : lens[i]]].tobytes() for i in range(2 * n)]
    host = af.pack_pairs(seqs[0::2], seqs[1::2], pad_byte=index.pad_byte)
    assert host.uniform_len == 0 and host.n_nreads > 100
    eng = af.Anchorer(index, 0)
    hits, _ = eng.anchor(host.to_device(0))
    assert len(want) > 500
    assert hits_equal(hits, want)
    # and the same batch through the host-buffer pipeline, in small chunks
    hits2, _ = eng.anchor_host(host, slot_pairs=512, n_slots=3)
    assert hits_equal(hits2, want)


@pytest.mark.parametrize("read_len", [300, 257, 400, 448, 512])
def test_long_reads_uniform(af, read_len):
    """Reads beyond 256 bases (2x300 MiSeq, merged pairs up to 512): the long-read instances of the scan (W = 20..32,
    no register double buffer), the bitmap verify with 32 words per read and the base-by-base extension masks --
    every record equals the oracle's, through device tiles, host tiles and the wire format."""
    from oracle import oracle
    spec = af.synth_spec(seed=900 + read_len, ref_len=400_000, anchor_start=150_000, anchor_len=7000, read_len=read_len,
                         frag_mean=2 * read_len + 100, frag_sd=40, sub_ppm=15_000, fusion_ppm=30_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    n = 40_000
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8)
    assert len(want) > 1000 and int(want["m_len"].max()) > 256
    hits, st = eng.anchor(af.synth_pairs_device(spec, 0, n, index.pad_byte, 0))
    assert hits_equal(hits, want)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    host = af.pack_pairs([lut[r].tobytes() for r in m1], [lut[r].tobytes() for r in m2], pad_byte=index.pad_byte)
    assert host.uniform_len == read_len
    hits2, _ = eng.anchor_host(host, slot_pairs=4096, n_slots=3)
    assert hits_equal(hits2, want)
    wire = af.wire_from_packed(np.asarray(host.packed).view(np.uint32), read_len, n)
    hits3, _ = eng.anchor_host(af.PackedBatch(wire, n, read_len, read_len), slot_pairs=4096, n_slots=3, wire=True)
    assert hits_equal(hits3, want)


def test_long_reads_ragged_with_n(af):
    """ragged lengths 19..500 in one batch, Ns in reads and anchor, reads hanging over both anchor ends, both strands."""
    from oracle import oracle
    rng = np.random.default_rng(11)
    G = 4000
    anchor = rng.integers(0, 4, G).astype(np.uint8)
    anchor[[100, 101, 2500]] = 4
    n, stride = 3000, 500
    reads = np.full((2 * n, stride), 4, dtype=np.uint8)
    lens = rng.integers(19, stride + 1, 2 * n).astype(np.uint16)
    lens[:8] = [500, 499, 257, 256, 300, 301, 19, 480]
    for i in range(2 * n):
        L = int(lens[i])
        kind = i % 4
        if kind == 0:
            r = rng.integers(0, 4, L)
        else:
            p = int(rng.integers(-60, G - L + 60))
            r = np.array([anchor[j] if 0 <= j < G and anchor[j] < 4 else rng.integers(0, 4) for j in range(p, p + L)])
            if kind == 2:
                j = int(rng.integers(0, L))
                r[j:] = rng.integers(0, 4, L - j)
            for _ in range(int(rng.integers(0, 6))):
                r[rng.integers(0, L)] = rng.integers(0, 5)
            if i % 8 >= 4:
                r = np.array([3 - c if c < 4 else 4 for c in r[::-1]])
        reads[i, :L] = r
    want = oracle.anchor_reads(anchor, reads, lens=lens)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    index = af.AnchorIndex(lut[anchor].tobytes())
    seqs = [lut[reads[i, : lens[i]]].tobytes() for i in range(2 * n)]
    host = af.pack_pairs(seqs[0::2], seqs[1::2], pad_byte=index.pad_byte)
    assert host.uniform_len == 0 and host.n_nreads > 100 and host.max_read_len > 256
    eng = af.Anchorer(index, 0)
    hits, _ = eng.anchor(host.to_device(0))
    assert len(want) > 500 and hits_equal(hits, want)
    hits2, _ = eng.anchor_host(host, slot_pairs=512, n_slots=3)
    assert hits_equal(hits2, want)


@pytest.mark.parametrize("over", [{"B": 2, "X": 12, "T": 35, "clip5": 3, "clip3": 8},
                                  {"A": 2, "B": 5, "X": 30, "T": 50, "clip5": 0, "clip3": 0}])
def test_other_scores_and_a_repetitive_anchor(af, over):
    """Non-default match / mismatch scores, clip penalties, X-drop and threshold, on an anchor holding a
    37-base unit five times and a 300-base segment three times (many seeded diagonals per read; ties go
    to the higher score, then strand 0, then the smaller diagonal)."""
    from oracle import oracle
    rng = np.random.default_rng(23)
    unit, seg = rng.integers(0, 4, 37), rng.integers(0, 4, 300)
    codes = np.concatenate([rng.integers(0, 4, 500), np.tile(unit, 5), rng.integers(0, 4, 400), seg, rng.integers(0, 4, 200), seg,
                            rng.integers(0, 4, 300), seg, rng.integers(0, 4, 500)]).astype(np.uint8)
    G = len(codes)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    n, L = 20_000, 100
    reads = rng.integers(0, 4, (2 * n, L)).astype(np.uint8)
    for i in range(0, 2 * n, 2):                       # every other read comes from the anchor
        p = int(rng.integers(-20, G - L + 20))
        src = np.array([codes[j] if 0 <= j < G else rng.integers(0, 4) for j in range(p, p + L)], dtype=np.uint8)
        if i % 3 == 0:
            j = int(rng.integers(10, L - 10))
            src[j:] = rng.integers(0, 4, L - j)
        for _ in range(int(rng.integers(0, 4))):
            src[rng.integers(0, L)] = rng.integers(0, 4)
        reads[i] = src if i % 4 else (3 - src)[::-1]
    want = oracle.anchor_reads(codes, reads, params=oracle.default_params(**over), threads=8)
    index = af.AnchorIndex(lut[codes].tobytes(), params=af.default_params(**over))
    host = af.pack_pairs([lut[r].tobytes() for r in reads[0::2]], [lut[r].tobytes() for r in reads[1::2]], pad_byte=index.pad_byte)
    hits, _ = af.Anchorer(index, 0).anchor(host.to_device(0))
    assert len(want) > 5000 and hits_equal(hits, want)


def test_unsupported_seed_length_is_refused(af):
    with pytest.raises(af.AnchoredFusionError, match="k=15 must be 19"):
        af.AnchorIndex("ACGT" * 100, params=af.default_params(k=15))


@pytest.mark.parametrize("anchor_len,bloom", [(40_000, True), (40_000, False), (20_000, False), (100_000, True)])
def test_long_anchor_stays_exact_with_bloom_filter_or_saturated_buckets(af, anchor_len, bloom, monkeypatch):
    """Anchors beyond ~12 kb overflow the 3-slot filter buckets ("always hit").  By default the index then carries a
    blocked Bloom filter in the same shared memory (far fewer flagged reads); AF_NO_BLOOM keeps the saturated buckets,
    where most reads are flagged and the exact verify stage decides.  Either way the records equal the oracle's."""
    from oracle import oracle
    if not bloom and anchor_len > 30_000:
        monkeypatch.setenv("AF_NO_BLOOM", "1")
    spec = af.synth_spec(seed=40, ref_len=600_000, anchor_start=200_000, anchor_len=anchor_len, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=15_000, fusion_ppm=20_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    assert index.bloom == bloom and index.info.n_overflow * 200 > index.info.n_buckets
    eng = af.Anchorer(index, 0)
    n = 60_000
    dev = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    hits, stats = eng.anchor(dev)
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8)
    assert len(want) > 2500 and hits_equal(hits, want)
    if anchor_len == 40_000:
        assert (stats["flagged"] < 0.7 * n) if bloom else (stats["flagged"] > n)      # of 2n reads: < 35 % against > 50 %
    # the scan's flag words are exactly what the filter words say (numpy emulation of the probe)
    from filter_emulator import expected_flags
    flags = eng.seed_scan(dev).cpu().numpy().view(np.uint32)
    sub = 4096
    codes = _interleave(m1, m2)[: 2 * sub]
    got = np.array([(flags[(r >> 1) >> 5, r & 1] >> ((r >> 1) & 31)) & 1 for r in range(2 * sub)], dtype=bool)
    assert np.array_equal(got, expected_flags(index, codes))
    # ... and the host-buffer pipeline takes the same route
    host = af.PackedBatch(dev.packed.cpu().numpy().view(np.uint32), n, 150, 150)
    hits2, _ = eng.anchor_host(host, slot_pairs=16_384, n_slots=2)
    assert hits_equal(hits2, want)


def test_robustness_set_n_bases_and_trimmed_reads(af):
    """SURVEY 8d robustness set: 200 k synthetic pairs with 0.1 % N bases, every third read trimmed to a
    random length (adapter / quality trimming), 1 % fusion fragments -- resident path and host pipeline."""
    from oracle import oracle
    spec = af.synth_spec(seed=77, ref_len=400_000, anchor_start=150_000, anchor_len=6783, read_len=150,
                         frag_mean=300, frag_sd=30, sub_ppm=15_000, fusion_ppm=10_000, n_ppm=1_000)
    anchor = af.synth_anchor(spec)
    n = 200_000
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    reads = _interleave(m1, m2)
    rng = np.random.default_rng(3)
    lens = np.full(2 * n, 150, dtype=np.uint16)
    cut = np.arange(0, 2 * n, 3)
    lens[cut] = rng.integers(30, 151, len(cut))
    for i in cut:
        reads[i, lens[i]:] = 4
    assert (reads == 4).sum() > 10_000
    want = oracle.anchor_reads(oracle.encode(anchor), reads, lens=lens, threads=8)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    text = lut[reads]
    seqs = [text[i, : lens[i]].tobytes() for i in range(2 * n)]
    index = af.AnchorIndex(anchor)
    host = af.pack_pairs(seqs[0::2], seqs[1::2], pad_byte=index.pad_byte)
    assert host.uniform_len == 0 and host.n_nreads > 20_000
    eng = af.Anchorer(index, 0)
    hits, _ = eng.anchor(host.to_device(0))
    assert len(want) > 3000
    assert hits_equal(hits, want)
    hits2, _ = eng.anchor_host(host, slot_pairs=1 << 16, n_slots=3)
    assert hits_equal(hits2, want)


@pytest.mark.parametrize("read_len", [101, 300])
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1025])
def test_edge_batch_sizes(af, n, read_len):
    from oracle import oracle
    spec = af.synth_spec(seed=2, ref_len=20_000, anchor_start=5_000, anchor_len=4000, read_len=read_len, frag_mean=2 * read_len + 50)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    hits, _ = eng.anchor(af.synth_pairs_device(spec, 0, n, index.pad_byte, 0))
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2))
    assert hits_equal(hits, want)


def test_pipeline_equals_resident_path_and_counts_launches(af):
    import ctypes
    from anchored_fusion_b200._lib import lib
    spec = af.synth_spec(seed=9, ref_len=500_000, anchor_start=100_000, anchor_len=6783, read_len=150,
                         sub_ppm=10_000, fusion_ppm=10_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, 0)
    n = 300_000
    dev = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    hits, _ = eng.anchor(dev)
    host = af.PackedBatch(dev.packed.cpu().numpy().view(np.uint32), n, 150, 150)
    before = lib().af_kernel_launches()
    hits2, _ = eng.anchor_host(host, slot_pairs=65_536, n_slots=3)
    assert hits_equal(hits, hits2)
    assert lib().af_kernel_launches() - before in (6 * ((n + 65_535) // 65_536), 2 * ((n + 65_535) // 65_536))   # six-kernel path / candidate-stream path (scan + k_tail)


def test_fused_scan_verify_kernel_gives_the_same_records(af, bundled):
    """af_seed_scan_config(0, 4): the warp-specialised fused scan+verify kernel (20 scan warps feed
    4 verify warps through shared-memory queues) instead of the separate kernels."""
    from oracle import oracle
    from anchored_fusion_b200._lib import check, lib
    check(lib().af_seed_scan_config(0, 4))
    try:
        index = af.AnchorIndex(bundled["anchor"])
        eng = af.Anchorer(index, 0)
        host = af.pack_pairs(bundled["seqs1"], bundled["seqs2"], pad_byte=index.pad_byte)
        hits, st = eng.anchor(host.to_device(0))
        assert hits_equal(hits, bundled["oracle_hits"]) and st["flagged"] >= len(hits)
        for L, kp in ((150, 12), (250, 13), (36, 12)):
            spec = af.synth_spec(seed=L, ref_len=300_000, anchor_start=100_000, anchor_len=9000, read_len=L,
                                 frag_mean=max(2 * L, 300), sub_ppm=15_000, fusion_ppm=30_000, )
            anchor = af.synth_anchor(spec)
            idx = af.AnchorIndex(anchor, kp=kp)
            e2 = af.Anchorer(idx, 0)
            n = 60_000
            h, _ = e2.anchor(af.synth_pairs_device(spec, 0, n, idx.pad_byte, 0))
            m1, m2 = af.synth_pairs_host(spec, 0, n)
            assert hits_equal(h, oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8))
    finally:
        check(lib().af_seed_scan_config(0, 5))


def test_pipeline_retry_after_a_capacity_error_returns_clean_records(af):
    """A too small host hit buffer fails with AF_ERR_CAPACITY; the natural retry with a larger buffer on the
    same pipeline must not see anything of the failed run (its in-flight slots are drained)."""
    from oracle import oracle
    spec = af.synth_spec(seed=19, ref_len=200_000, anchor_start=50_000, anchor_len=6783, read_len=150,
                         sub_ppm=10_000, fusion_ppm=20_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    n = 200_000
    dev = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    host = af.PackedBatch(dev.packed.cpu().numpy().view(np.uint32), n, 150, 150)
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    want = oracle.anchor_reads(oracle.encode(anchor), _interleave(m1, m2), threads=8)
    assert len(want) > 2000
    small = np.zeros(len(want) // 3, dtype=af.HIT_DTYPE)         # overflows in the second or third chunk
    with pytest.raises(af.AnchoredFusionError):
        eng.anchor_host(host, slot_pairs=32_768, n_slots=3, hits_out=small)
    hits, _ = eng.anchor_host(host, slot_pairs=32_768, n_slots=3)
    assert hits_equal(hits, want)


def test_capacity_overflow_is_reported_not_dropped(af):
    spec = af.synth_spec(seed=4, ref_len=20_000, anchor_start=1_000, anchor_len=15_000, read_len=150)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, 0)
    dev = af.synth_pairs_device(spec, 0, 50_000, index.pad_byte, 0)
    with pytest.raises(af.AnchoredFusionError):
        eng.anchor(dev, cand_cap=1000, hits_cap=1000)


@pytest.mark.parametrize("modes,read_len", [((), 150), ((11,), 150), ((4,), 150), ((10,), 150), ((8,), 150), ((), 300), ((), 500)])
def test_kernels_stay_inside_their_buffers(af, modes, read_len):
    """compute-sanitizer is not available on this pool, so out-of-bounds WRITES are looked for directly:
    workspace, hit list, counters and the packed batch sit between 64 KB canary zones, which must come
    back untouched (default path; the scan with the in-kernel refinement queue; the fused scan+verify kernel;
    the bitmap verify kernel; no middle stage at all), with a pair
    count that is not a multiple of anything and capacities small enough to be hit exactly."""
    import ctypes
    import torch
    from anchored_fusion_b200._lib import check, lib
    spec = af.synth_spec(seed=91, ref_len=300_000, anchor_start=100_000, anchor_len=5000, read_len=read_len,
                         frag_mean=2 * read_len, sub_ppm=15_000, fusion_ppm=50_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, 0)
    n = 77_777 if read_len == 150 else 33_333
    ref_hits, stats = eng.anchor(af.synth_pairs_device(spec, 0, n, index.pad_byte, 0))
    G, dev = 1 << 16, torch.device("cuda", 0)

    def guarded(nbytes):
        big = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device=dev)
        return big, big[G: G + nbytes]

    lay = af.layout(read_len, n)
    src = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    big_p, packed = guarded(lay.packed_bytes)
    packed.copy_(src.packed.view(torch.uint8)[: lay.packed_bytes])
    cand_cap = stats["flagged"] + 1                      # just enough
    hits_cap = len(ref_hits)                             # exactly enough
    ws_bytes = lib().af_workspace_bytes_len(n, cand_cap, read_len)    # exactly what this read length needs
    big_w, ws = guarded(ws_bytes)
    big_h, hits = guarded(hits_cap * 16)
    big_c, counts = guarded(32)
    batch = af.PackedBatch(packed.view(torch.int32), n, read_len, read_len)
    cb = batch.c_struct()
    for m in modes:
        check(lib().af_seed_scan_config(0, m))
    try:
        check(lib().af_anchor_batch(eng.dindex._h, ctypes.byref(cb), ws.data_ptr(), ws_bytes, cand_cap, hits.data_ptr(), hits_cap,
                                    counts.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.synchronize()
    finally:
        check(lib().af_seed_scan_config(0, 5))
        check(lib().af_seed_scan_config(0, 9))
        check(lib().af_seed_scan_config(0, 7))
    c = counts.view(torch.int32).cpu().numpy()
    assert c[2] == 0 and c[1] == len(ref_hits)
    got = hits.cpu().numpy().view(af.HIT_DTYPE)
    assert hits_equal(got, ref_hits)
    for name, big, nbytes in (("packed", big_p, lay.packed_bytes), ("workspace", big_w, ws_bytes), ("hits", big_h, hits_cap * 16),
                              ("counts", big_c, 32)):
        assert bool((big[:G] == 0xA5).all()) and bool((big[G + nbytes:] == 0xA5).all()), name
