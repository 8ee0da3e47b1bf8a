"""The seed-scan probe sequence (af_scan_read, the template k_seed_scan runs) executed on the
host for every read length class, against the numpy emulation of the filter: no false negatives,
and exactly the filter's positives."""
import ctypes

import numpy as np
import pytest


def _scan_pairs_host(af, index, batch, refine=False):
    from anchored_fusion_b200._lib import check, lib
    lay = af.layout(batch.max_read_len, batch.n_pairs)
    W, Q = lay.words_per_read, lay.quads_per_pair
    L = batch.uniform_len if batch.uniform_len > 0 else batch.max_read_len
    packed = batch.packed.reshape(-1)
    out = np.zeros(2 * batch.n_pairs, bool)
    f1, f2 = ctypes.c_int32(), ctypes.c_int32()
    for p in range(batch.n_pairs):
        tile, lane = p >> 5, p & 31
        words = np.concatenate([packed[((tile * Q + q) * 32 + lane) * 4: ((tile * Q + q) * 32 + lane) * 4 + 4] for q in range(Q)])[: 2 * W]
        words = np.ascontiguousarray(words, dtype=np.uint32)
        check(lib().af_debug_scan_pair(index._h, words.ctypes.data, W, L, int(refine), ctypes.byref(f1), ctypes.byref(f2)))
        out[2 * p], out[2 * p + 1] = bool(f1.value), bool(f2.value)
    return out


@pytest.mark.parametrize("read_len", [36, 76, 101, 125, 150, 151, 250, 300, 410])
@pytest.mark.parametrize("kp", [12, 13])
def test_host_twin_equals_emulation_and_has_no_false_negatives(read_len, kp):
    import anchored_fusion_b200 as af
    from filter_emulator import expected_flags
    from oracle import oracle
    spec = af.synth_spec(seed=read_len * 7 + kp, ref_len=60_000, anchor_start=20_000, anchor_len=8000, read_len=read_len,
                         frag_mean=max(2 * read_len, 200), sub_ppm=20_000, fusion_ppm=100_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor, kp=kp)
    n = 1500
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    batch = af.pack_pairs([lut[r].tobytes() for r in m1], [lut[r].tobytes() for r in m2], pad_byte=index.pad_byte)
    got = _scan_pairs_host(af, index, batch)
    codes = np.empty((2 * n, read_len), dtype=np.uint8)
    codes[0::2], codes[1::2] = m1, m2
    want = expected_flags(index, codes)
    assert np.array_equal(got, want)
    hits = oracle.anchor_reads(oracle.encode(anchor), codes, threads=4)
    assert len(hits) > 50 and got[hits["read_id"]].all()
    # the optional neighbour test (a sample counts only if the k'-mer at p-H or p+H passes too) loses
    # no anchored read either and removes most chance hits
    ref = _scan_pairs_host(af, index, batch, refine=True)
    assert np.array_equal(ref, expected_flags(index, codes, refine=True))
    assert ref[hits["read_id"]].all() and not (ref & ~got).any() and ref.sum() <= got.sum()


@pytest.mark.parametrize("anchor_len", [40_000, 100_000])
def test_long_anchor_switches_to_the_bloom_filter_without_false_negatives(anchor_len, monkeypatch):
    """Beyond ~12 kb the 3-slot buckets overflow; the index then holds a blocked Bloom filter in the same words.  The
    probe sequence the kernel runs (host twin) equals the numpy emulation, flags every anchored read, and flags far
    fewer random reads than the overflowing buckets would."""
    import anchored_fusion_b200 as af
    from filter_emulator import expected_flags
    from oracle import oracle
    read_len, n = 150, 1200
    spec = af.synth_spec(seed=anchor_len, ref_len=200_000, anchor_start=50_000, anchor_len=anchor_len, read_len=read_len,
                         frag_mean=300, sub_ppm=15_000, fusion_ppm=50_000)
    anchor = af.synth_anchor(spec)
    index = af.AnchorIndex(anchor)
    assert index.bloom and index.info.n_overflow * 25 > index.info.n_buckets
    assert not af.AnchorIndex(anchor[:6783]).bloom and not af.AnchorIndex(anchor[:20_000]).bloom
    m1, m2 = af.synth_pairs_host(spec, 0, n)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    batch = af.pack_pairs([lut[r].tobytes() for r in m1], [lut[r].tobytes() for r in m2], pad_byte=index.pad_byte)
    got = _scan_pairs_host(af, index, batch)
    codes = np.empty((2 * n, read_len), dtype=np.uint8)
    codes[0::2], codes[1::2] = m1, m2
    assert np.array_equal(got, expected_flags(index, codes))
    hits = oracle.anchor_reads(oracle.encode(anchor), codes, threads=4)
    assert len(hits) > 100 and got[hits["read_id"]].all()
    # reads that do not touch the anchor: the Bloom filter lets far fewer through than the saturated buckets
    rng = np.random.default_rng(1)
    rnd = rng.integers(0, 4, (2 * n, read_len)).astype(np.uint8)
    rb = af.pack_pairs([lut[r].tobytes() for r in rnd[0::2]], [lut[r].tobytes() for r in rnd[1::2]], pad_byte=index.pad_byte)
    bloom_rate = _scan_pairs_host(af, index, rb).mean()
    monkeypatch.setenv("AF_NO_BLOOM", "1")
    plain = af.AnchorIndex(anchor)
    assert not plain.bloom
    bucket_rate = _scan_pairs_host(af, plain, af.pack_pairs([lut[r].tobytes() for r in rnd[0::2]], [lut[r].tobytes() for r in rnd[1::2]],
                                                            pad_byte=plain.pad_byte)).mean()
    assert bloom_rate < 0.7 * bucket_rate and bloom_rate < (0.25 if anchor_len == 40_000 else 0.7), (bloom_rate, bucket_rate)


@pytest.mark.parametrize("kp", [12, 13])
def test_sample_grid_covers_every_window_of_19_matches(kp):
    """The sample grid (af_sample0 / af_nsamples, csrc/af_common.h) starts at base 4 (k' = 12) or 6 (k' = 13) and has
    one sample fewer than a zero-based grid.  Exhaustive check of its guarantee through the kernel's probe sequence
    (host twin): a read that matches the anchor in exactly one window of 19 bases [q, q+19) -- everything else
    mismatching -- is flagged for EVERY q in 0..L-19, at every read length class including both ends of each word
    count and the long-read instances."""
    import anchored_fusion_b200 as af
    rng = np.random.default_rng(kp)
    G = 3000
    anchor = rng.integers(0, 4, G).astype(np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    index = af.AnchorIndex(lut[anchor].tobytes(), kp=kp)
    for L in (19, 20, 26, 27, 28, 36, 50, 76, 101, 125, 144, 145, 150, 151, 160, 161, 250, 256, 257, 300, 301, 320, 321, 512):
        qs = list(range(L - 18))
        reads = []
        for q in qs:
            a0 = int(rng.integers(L, G - 2 * L))
            r = (anchor[a0 - q: a0 - q + L] + 1 + rng.integers(0, 3, L)) % 4          # every base differs from the anchor's ...
            r[q:q + 19] = anchor[a0: a0 + 19]                                          # ... except the window
            if q % 2:                                                                  # odd q: the read is the reverse strand
                r = (3 - r)[::-1]
            reads.append(lut[r.astype(np.uint8)].tobytes())
        if len(reads) % 2:
            reads.append(reads[-1])
        batch = af.pack_pairs(reads[0::2], reads[1::2], pad_byte=index.pad_byte)
        got = _scan_pairs_host(af, index, batch)
        assert got.all(), (L, [qs[i] for i in np.flatnonzero(~got[: len(qs)])])
