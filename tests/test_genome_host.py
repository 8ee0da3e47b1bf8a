"""Host half of the genome pass (csrc/af_genome_host.cpp): FASTA (plain / gzip / CRLF / no final newline) -> the
concatenation [256 N] contig [256 N] ... [256 N], checked against a sequence built here.  No GPU needed."""
import ctypes
import gzip

import numpy as np
import pytest

SEP = 256


def _fnv(codes):
    x = 1469598103934665603
    for c in codes:
        x = ((x ^ int(c)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return x


def _parse(path):
    from anchored_fusion_b200._lib import check, lib
    n, k, h = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_uint64()
    check(lib().af_debug_genome_fasta(str(path).encode(), ctypes.byref(n), ctypes.byref(k), ctypes.byref(h)))
    return n.value, k.value, h.value


def _codes(seq):
    lut = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = lut[ord(ch.lower())] = i
    return lut[np.frombuffer(seq.encode(), dtype=np.uint8)]


def test_fasta_forms_give_the_same_concatenation(tmp_path):
    rng = np.random.default_rng(2)
    contigs = []
    for name, n in (("chr1 first contig", 5000), ("chr2", 1), ("empty", 0), ("chr4\tx", 777)):
        contigs.append((name, "".join("ACGTNacgtnRY"[c] for c in rng.integers(0, 12, n))))
    concat = "N" * SEP + "".join(s + "N" * SEP for _, s in contigs)
    want = (len(concat), len(contigs), _fnv(_codes(concat)))

    def text(width, eol, final_newline=True):
        t = "".join(">" + n + eol + "".join(s[i:i + width] + eol for i in range(0, len(s), width)) for n, s in contigs)
        return t if final_newline else t.rstrip("\r\n")

    forms = {"a.fa": text(60, "\n"), "b.fa": text(7, "\r\n"), "c.fa": text(100000, "\n", final_newline=False), "d.fa": text(61, "\n").replace("\n", "\n\n", 3)}
    for fn, t in forms.items():
        (tmp_path / fn).write_text(t, newline="")
        assert _parse(tmp_path / fn) == want, fn
    with gzip.open(tmp_path / "e.fa.gz", "wt", newline="") as fh:
        fh.write(forms["a.fa"])
    assert _parse(tmp_path / "e.fa.gz") == want


def test_fasta_errors_are_reported(tmp_path):
    from anchored_fusion_b200 import AnchoredFusionError
    with pytest.raises(AnchoredFusionError, match="cannot open"):
        _parse(tmp_path / "missing.fa")
    (tmp_path / "nohdr.fa").write_text("ACGT\n>x\nAC\n")
    with pytest.raises(AnchoredFusionError, match="header"):
        _parse(tmp_path / "nohdr.fa")
    (tmp_path / "empty.fa").write_text("")
    with pytest.raises(AnchoredFusionError, match="no sequence"):
        _parse(tmp_path / "empty.fa")
    good = gzip.compress(b">c\n" + b"ACGT" * 5000 + b"\n")
    (tmp_path / "cut.fa.gz").write_bytes(good[: len(good) // 2])
    with pytest.raises(AnchoredFusionError):
        _parse(tmp_path / "cut.fa.gz")


def test_packed_genome_cache_is_used_validated_and_rebuilt(tmp_path, monkeypatch):
    """<fasta>.af2bit: written on the first load, read on the second, ignored (and rewritten) when the FASTA changed or
    the cache is damaged; AF_GENOME_CACHE=0 switches it off."""
    import os
    rng = np.random.default_rng(4)
    contigs = [("c%d" % i, "".join("ACGTN"[c] for c in rng.integers(0, 5, n))) for i, n in enumerate((3000, 10, 70000))]
    fa = tmp_path / "g.fa"
    fa.write_text("".join(">%s\n%s\n" % c for c in contigs))
    concat = "N" * SEP + "".join(s + "N" * SEP for _, s in contigs)
    want = (len(concat), len(contigs), _fnv(_codes(concat)))
    cache = tmp_path / "g.fa.af2bit"
    assert _parse(fa) == want and cache.exists()
    stamp = cache.stat().st_mtime_ns
    assert _parse(fa) == want and cache.stat().st_mtime_ns == stamp          # second load: cache read, not rewritten
    raw = bytearray(cache.read_bytes())
    cache.write_bytes(bytes(raw[: len(raw) // 2]))                             # truncated cache: ignored, rebuilt
    assert _parse(fa) == want and cache.stat().st_size == len(raw)
    flipped = bytearray(raw)
    flipped[len(raw) // 2] ^= 0x10                                             # one bit of the packed bases: the checksum catches it
    cache.write_bytes(bytes(flipped))
    assert _parse(fa) == want and cache.read_bytes() == bytes(raw)
    contigs2 = contigs + [("extra", "ACGT" * 100)]
    fa.write_text("".join(">%s\n%s\n" % c for c in contigs2))                  # the FASTA changed: size stamp differs
    concat2 = "N" * SEP + "".join(s + "N" * SEP for _, s in contigs2)
    assert _parse(fa) == (len(concat2), len(contigs2), _fnv(_codes(concat2)))
    raw = bytearray(cache.read_bytes())
    raw[16:24] = (1 << 40).to_bytes(8, "little")                               # header claims another FASTA size
    cache.write_bytes(bytes(raw))
    assert _parse(fa) == (len(concat2), len(contigs2), _fnv(_codes(concat2)))
    monkeypatch.setenv("AF_GENOME_CACHE", "0")
    os.remove(cache)
    assert _parse(fa)[0] == len(concat2) and not cache.exists()
