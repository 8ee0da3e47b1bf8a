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


def build(force=False):
    """Compile af_oracle.c with gcc (no reference sources involved)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaf_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


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
