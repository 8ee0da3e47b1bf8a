"""The C-ABI library loads and exports every symbol include/anchored_fusion.h declares
(no compute calls that need a GPU), and the host-side entry points behave."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "anchored_fusion.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(af_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from anchored_fusion_b200._lib import LIB_PATH, SIGNATURES
    names = _declared_symbols()
    assert len(names) >= 30
    handle = ctypes.CDLL(LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "libafb200.so lacks %s" % n
        assert n in SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(SIGNATURES) == names


def test_struct_sizes_match_the_header():
    from anchored_fusion_b200 import _lib
    assert ctypes.sizeof(_lib.Params) == 28
    assert _lib.HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(_lib.Layout) == 40
    assert ctypes.sizeof(_lib.Batch) == 56
    assert ctypes.sizeof(_lib.Synth) == 56
    assert ctypes.sizeof(_lib.IndexInfo) == 44
    assert _lib.GENOME_HIT_DTYPE.itemsize == 24          # af_genome_hit_t
    assert ctypes.sizeof(_lib.GenomeStats) == 56         # af_genome_stats_t


def test_no_cpu_fallback_device_entry_points_fail_loudly_without_a_gpu():
    import torch
    import anchored_fusion_b200 as af
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    idx = af.AnchorIndex("ACGT" * 50)
    with pytest.raises(af.AnchoredFusionError):
        idx.upload(0)
    with pytest.raises(af.AnchoredFusionError):
        af.Anchorer(idx, 0)
    from anchored_fusion_b200.dist import HitExchange
    with pytest.raises(af.AnchoredFusionError):
        HitExchange(0, 1, 1, 64, 0)              # the hit exchange lives in device memory: no GPU, no exchange
    from anchored_fusion_b200.genome import Genome
    with pytest.raises(af.AnchoredFusionError):
        Genome.from_contigs([("c", "ACGT" * 100)])    # the genome pass streams a genome resident in HBM: no GPU, no genome
    with pytest.raises(af.AnchoredFusionError):
        Genome.synthetic(1, 100000)


def test_only_the_seed_length_the_kernels_are_built_for_is_accepted():
    import anchored_fusion_b200 as af
    with pytest.raises(af.AnchoredFusionError, match="k=15 must be 19"):
        af.AnchorIndex("ACGT" * 100, params=af.default_params(k=15))
    with pytest.raises(af.AnchoredFusionError, match="kp=11"):
        af.AnchorIndex("ACGT" * 100, kp=11)
    af.AnchorIndex("ACGT" * 100, params=af.default_params(B=2, X=12, T=35, clip5=3, clip3=8))    # scores are free


def test_product_never_touches_the_oracle():
    """anchored_fusion_b200/ must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "anchored_fusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "af_oracle" not in text.replace("oracle/af_oracle.c", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_layout_and_errors():
    import anchored_fusion_b200 as af
    lay = af.layout(150, 10_000_000)
    assert (lay.words_per_read, lay.quads_per_pair, lay.n_tiles, lay.packed_bytes) == (10, 5, 312_500, 800_000_000)
    lay = af.layout(101, 33)
    assert (lay.words_per_read, lay.quads_per_pair, lay.n_tiles, lay.packed_bytes) == (7, 4, 2, 4096)
    lay = af.layout(300, 64)                            # beyond 256 bases W is rounded up to a multiple of 4 (long-read scan)
    assert (lay.words_per_read, lay.quads_per_pair, lay.n_tiles, lay.packed_bytes) == (20, 10, 2, 10240)
    assert af.layout(257, 1).words_per_read == 20 and af.layout(512, 1).words_per_read == 32
    with pytest.raises(af.AnchoredFusionError):
        af.layout(513, 1)
    with pytest.raises(af.AnchoredFusionError):
        af.AnchorIndex("ACGT" * 10, kp=20)
    with pytest.raises(af.AnchoredFusionError):
        af.pack_pairs(["A" * 200], ["C" * 10], max_read_len=150)


def test_index_contents(bundled):
    """Every k'-mer of both anchor strands is in the exact table at its position and passes the
    shared-memory filter (no false negatives by construction)."""
    import anchored_fusion_b200 as af
    from oracle import oracle
    for kp in (12, 13):
        idx = af.AnchorIndex(bundled["anchor"], kp=kp)
        info = idx.info
        assert (info.k, info.kp, info.stride, info.anchor_len) == (19, kp, 20 - kp, 6783)
        tab = idx.table_words()
        live = tab[tab[:, 0] != 0xFFFFFFFF]
        a = oracle.encode(bundled["anchor"]).astype(np.uint64)
        G = len(a)
        fwd = np.zeros(G - kp + 1, dtype=np.uint64)
        rc = np.zeros(G - kp + 1, dtype=np.uint64)
        for t in range(kp):
            fwd |= a[t: G - kp + 1 + t] << np.uint64(2 * t)
            rc |= (np.uint64(3) - a[kp - 1 - t: G - t]) << np.uint64(2 * t)
        want = sorted([(int(k), j) for j, k in enumerate(fwd)] + [(int(k), (1 << 31) | j) for j, k in enumerate(rc)])
        assert sorted((int(k), int(v)) for k, v in live) == want
        assert info.n_entries == len(want) and info.n_keys == len(set(k for k, _ in want))
        # filter membership of every key
        filt = idx.filter_words().astype(np.uint64)
        keys = np.array(sorted(set(k for k, _ in want)), dtype=np.uint64)
        from filter_emulator import filter_hash
        b, fp3 = filter_hash(keys, info.filter_mul, kp, info.n_buckets)
        v = filt[b] ^ fp3
        hit = ((v - np.uint64(0x40100401)) & ~v & np.uint64(0xA0080200)) != 0
        assert hit.all()
        # the pad pattern's k'-mers are not anchor k'-mers
        pad = [(info.pad_byte >> (2 * i)) & 3 for i in range(4)]
        for ph in range(4):
            key = sum(pad[(i + ph) & 3] << (2 * i) for i in range(kp))
            assert key not in set(int(k) for k in keys)


def test_pack_roundtrip_ragged_with_n():
    import anchored_fusion_b200 as af
    rng = np.random.default_rng(0)
    s1 = ["".join(rng.choice(list("ACGTN"), p=[.24, .24, .24, .24, .04], size=rng.integers(1, 151))) for _ in range(70)]
    s2 = ["".join(rng.choice(list("acgt"), size=rng.integers(1, 151))) for _ in range(70)]
    pad_byte = 0x1B
    b = af.pack_pairs(s1, s2, pad_byte=pad_byte, max_read_len=150)
    pad = [(pad_byte >> (2 * k)) & 3 for k in range(4)]
    assert b.uniform_len == 0 and b.n_pairs == 70 and b.packed.nbytes == 3 * 5 * 512
    nset = dict(zip(b.nread_ids.tolist(), b.nmask))
    for p in range(70):
        for m, ss in ((0, s1), (1, s2)):
            rid = 2 * p + m
            assert b.lens[rid] == len(ss[p])
            got = af.unpack_read(b, rid)
            want = ["ACGT".index(c.upper()) if c.upper() != "N" else pad[k & 3] for k, c in enumerate(ss[p])]
            assert got.tolist() == want
            npos = [k for k, c in enumerate(ss[p]) if c == "N"]
            if npos:
                mask = nset[rid]
                assert [k for k in range(512) if (mask[k >> 5] >> (k & 31)) & 1] == npos
            else:
                assert rid not in nset
    assert sorted(nset) == b.nread_ids.tolist()
    # uniform batch, empty batch
    u = af.pack_pairs(["ACGT" * 5] * 3, ["TTTT" * 5] * 3)
    assert u.uniform_len == 20 and u.nread_ids is None
    e = af.pack_pairs([], [], max_read_len=150)
    assert e.n_pairs == 0


def test_host_synth_is_deterministic_and_plants_the_anchor():
    import anchored_fusion_b200 as af
    from oracle import oracle
    spec = af.synth_spec(seed=5, ref_len=50_000, anchor_start=10_000, anchor_len=3000, read_len=100,
                         frag_mean=250, sub_ppm=0, fusion_ppm=200_000)
    a1, a2 = af.synth_pairs_host(spec, 1000, 500)
    b1, b2 = af.synth_pairs_host(spec, 1200, 300)
    assert np.array_equal(a1[200:], b1) and np.array_equal(a2[200:], b2)    # pure function of the pair index
    anchor = oracle.encode(af.synth_anchor(spec))
    reads = np.empty((1000, 100), dtype=np.uint8)
    reads[0::2], reads[1::2] = a1, a2
    hits = oracle.anchor_reads(anchor, reads)
    assert len(set(hits["read_id"] >> 1)) >= 0.15 * 500       # ~20 % fusion fragments + natural overlap
    assert ((hits["clip_l"] > 0) | (hits["clip_r"] > 0)).sum() > 20   # junction reads are soft-clipped


def test_wire_format_round_trip_on_the_host():
    """tiles -> wire (4 L bits per pair, no padding) -> tiles is the identity, for ragged reads with N at every word edge."""
    import anchored_fusion_b200 as af
    rng = np.random.default_rng(0)
    for L, n in ((150, 70), (36, 33), (101, 1), (250, 64), (256, 5), (16, 40), (17, 31), (1, 3), (32, 32), (257, 9), (300, 40), (301, 33), (512, 7)):
        lens = rng.integers(max(1, L - 20), L + 1, (2, n))
        lens[0, 0] = L
        seqs = [["".join("ACGTN"[c] for c in rng.choice(5, int(l), p=[.245, .245, .245, .245, .02])) for l in row] for row in lens]
        b = af.pack_pairs(seqs[0], seqs[1], max_read_len=L, pad_byte=0x6C)
        tiles = np.asarray(b.packed).view(np.uint32)
        wire = af.wire_from_packed(tiles, L, n)
        assert wire.nbytes == af.wire_bytes(L, n) == ((n + 31) // 32) * ((4 * L + 31) // 32) * 128
        assert np.array_equal(af.wire_to_packed(wire, L, n, 0x6C), tiles), (L, n)
    assert af.wire_bytes(150, 10_000_000) == 760_000_000          # the algorithmic 76 bytes per 2 x 150 bp pair


def test_python_constants_follow_the_header():
    """_lib.py mirrors a few #defines of include/anchored_fusion.h (array strides, limits); a drift would corrupt the
    N-mask arrays silently."""
    import anchored_fusion_b200 as af  # noqa: F401
    from anchored_fusion_b200 import _lib
    text = open(os.path.join(ROOT, "include", "anchored_fusion.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+(AF_[A-Z_]+)\s+(\d+)\b", text, flags=re.M)}
    assert defs["AF_MAX_READ_LEN"] == _lib.MAX_READ_LEN == 512
    assert defs["AF_NMASK_WORDS"] == _lib.NMASK_WORDS == defs["AF_MAX_READ_LEN"] // 32
    assert defs["AF_GENOME_MAX_READ_LEN"] == _lib.GENOME_MAX_READ_LEN
    assert defs["AF_ABI_VERSION"] == _lib.lib().af_abi_version()
