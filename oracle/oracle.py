"""ctypes wrapper around oracle/af_oracle.c (the CPU restatement of the anchoring pass).

TEST INFRASTRUCTURE ONLY -- see the header of af_oracle.c for what is restated
(Anchored_Fusion.py:172,182,194 -> bwa index / bwa mem -M / samtools view -F 772) and
why parity is unpinned at the bwa boundary.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaf_oracle.so")
_SRC = os.path.join(_HERE, "af_oracle.c")

HIT_DTYPE = np.dtype([("read_id", "<u4"), ("pos", "<i4"), ("clip_l", "<u2"), ("m_len", "<u2"),
                      ("clip_r", "<u2"), ("score_strand", "<u2")])


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("k", "A", "B", "clip5", "clip3", "T", "X")]


_SYNTH_SO = os.path.join(_HERE, "libaf_synth.so")
_SYNTH_SRC = os.path.join(_HERE, "af_synth.cpp")
_SYNTH_HDR = os.path.join(_HERE, "..", "anchored_fusion_b200", "csrc", "af_common.h")


def build(force=False):
    """Compile af_oracle.c with gcc (no reference sources involved) and the CPU arm's pair generator."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaf_oracle.so"], stdout=subprocess.DEVNULL, env=env)
    if force or not os.path.exists(_SYNTH_SO) or os.path.getmtime(_SYNTH_SO) < max(os.path.getmtime(_SYNTH_SRC), os.path.getmtime(_SYNTH_HDR)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaf_synth.so"], stdout=subprocess.DEVNULL, env=env)
    sw_so, sw_src = os.path.join(_HERE, "libaf_sw.so"), os.path.join(_HERE, "af_sw.c")   # exhaustive DP the oracle is checked against
    if force or not os.path.exists(sw_so) or os.path.getmtime(sw_so) < os.path.getmtime(sw_src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaf_sw.so"], stdout=subprocess.DEVNULL, env=env)
    return _SO


class Synth(ctypes.Structure):
    """af_synth_t (include/anchored_fusion.h): the seeded generator's description."""
    _fields_ = [("seed", ctypes.c_uint64), ("ref_len", ctypes.c_int64), ("anchor_start", ctypes.c_int64),
                ("anchor_len", ctypes.c_int32), ("read_len", ctypes.c_int32), ("frag_mean", ctypes.c_int32),
                ("frag_sd", ctypes.c_int32), ("sub_ppm", ctypes.c_uint32), ("fusion_ppm", ctypes.c_uint32),
                ("n_ppm", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


_synth = None


def _synth_lib():
    global _synth
    if _synth is None:
        build()
        _synth = ctypes.CDLL(_SYNTH_SO)
        _synth.afo_synth_anchor.restype = ctypes.c_int
        _synth.afo_synth_anchor.argtypes = [ctypes.POINTER(Synth), ctypes.c_void_p]
        _synth.afo_synth_reads.restype = ctypes.c_int
        _synth.afo_synth_reads.argtypes = [ctypes.POINTER(Synth), ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_int]
        _synth.afo_synth_fastq.restype = ctypes.c_int
        _synth.afo_synth_fastq.argtypes = [ctypes.POINTER(Synth), ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int]
    return _synth


FASTQ_PLAIN, FASTQ_GZIP, FASTQ_BGZF = 0, 1, 2


def synth_fastq(spec, first_pair, pairs_per_file, paths1, paths2, fmt=FASTQ_GZIP, level=6, threads=1):
    """Write the generator's pairs as FASTQ files (measurement input of the ingest path): file i of paths1 /
    paths2 holds mates 1 / 2 of pairs [first_pair + i * pairs_per_file, +pairs_per_file).  Illumina-style names,
    binned qualities; fmt: FASTQ_PLAIN, FASTQ_GZIP (one member) or FASTQ_BGZF."""
    spec = as_synth(spec)
    n = len(paths1)
    assert n == len(paths2) and n > 0
    arr = ctypes.c_char_p * n
    a1, a2 = arr(*[p.encode() for p in paths1]), arr(*[p.encode() for p in paths2])
    rc = _synth_lib().afo_synth_fastq(ctypes.byref(spec), first_pair, pairs_per_file, n, a1, a2, fmt, level, threads)
    assert rc == 0, "writing synthetic FASTQ files failed"


def synth_spec(seed=1, ref_len=10_000_000, anchor_start=1_000_000, anchor_len=6783, read_len=150, frag_mean=300,
               frag_sd=30, sub_ppm=0, fusion_ppm=0, n_ppm=0):
    return Synth(seed, ref_len, anchor_start, anchor_len, read_len, frag_mean, frag_sd, sub_ppm, fusion_ppm, n_ppm, 0)


def as_synth(spec):
    """Any object with af_synth_t's fields (e.g. the product's Synth structure) -> this module's Synth."""
    return Synth(*[getattr(spec, n) for n, _ in Synth._fields_])


def synth_anchor(spec):
    spec = as_synth(spec)
    buf = ctypes.create_string_buffer(spec.anchor_len)
    assert _synth_lib().afo_synth_anchor(ctypes.byref(spec), ctypes.cast(buf, ctypes.c_void_p)) == 0
    return buf.raw[: spec.anchor_len]


def synth_reads(spec, first_pair, n_pairs, threads=1, out=None):
    """(2*n_pairs, read_len) uint8 codes, mates interleaved, of the seeded generator the GPU arm uses."""
    spec = as_synth(spec)
    reads = out if out is not None else np.empty((2 * n_pairs, spec.read_len), dtype=np.uint8)
    assert reads.shape[0] >= 2 * n_pairs and reads.strides[1] == 1
    assert _synth_lib().afo_synth_reads(ctypes.byref(spec), first_pair, n_pairs, reads.ctypes.data, reads.strides[0], threads) == 0
    return reads[: 2 * n_pairs]


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = ctypes.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(_SO)
        _lib.afo_anchor_reads.restype = ctypes.c_int
        _lib.afo_anchor_reads.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(Params),
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64),
                                          ctypes.c_int]
        _lib.afo_diag_eval.restype = ctypes.c_int
        _lib.afo_diag_eval.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.POINTER(Params)] + [ctypes.POINTER(ctypes.c_int32)] * 3
        _lib.afo_default_params.argtypes = [ctypes.POINTER(Params)]
    return _lib


def default_params(**over):
    p = Params()
    lib().afo_default_params(ctypes.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq):
    """ASCII bases -> codes 0..3, everything else 4."""
    if isinstance(seq, str):
        seq = seq.encode()
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def decode(codes):
    return "".join("ACGTN"[c] for c in codes)


def anchor_reads(anchor_codes, reads_codes, lens=None, params=None, threads=1):
    """reads_codes: (n_reads, stride) uint8 codes.  Returns structured array HIT_DTYPE ordered by read_id."""
    anchor_codes = np.ascontiguousarray(anchor_codes, dtype=np.uint8)
    reads_codes = np.ascontiguousarray(reads_codes, dtype=np.uint8)
    assert reads_codes.ndim == 2
    n, stride = reads_codes.shape
    if lens is not None:
        lens = np.ascontiguousarray(lens, dtype=np.uint16)
        assert lens.shape == (n,)
    p = params or default_params()
    out = np.zeros(max(n, 1), dtype=HIT_DTYPE)
    n_out = ctypes.c_int64(0)
    rc = lib().afo_anchor_reads(anchor_codes.ctypes.data, len(anchor_codes), reads_codes.ctypes.data,
                                lens.ctypes.data if lens is not None else None, n, stride, ctypes.byref(p),
                                out.ctypes.data, len(out), ctypes.byref(n_out), threads)
    assert rc == 0
    return out[: n_out.value].copy()


def diag_eval(q_codes, anchor_codes, d, params=None):
    q = np.ascontiguousarray(q_codes, dtype=np.uint8)
    a = np.ascontiguousarray(anchor_codes, dtype=np.uint8)
    p = params or default_params()
    sc, qb, qe = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    ok = lib().afo_diag_eval(q.ctypes.data, len(q), a.ctypes.data, len(a), d, ctypes.byref(p),
                             ctypes.byref(sc), ctypes.byref(qb), ctypes.byref(qe))
    return (sc.value, qb.value, qe.value) if ok else None
