"""The oracle's alignment arithmetic against a first-principles statement of bwa-mem's scoring model.

bwa is absent, so nothing the reference ships can confirm the records of the anchoring spec.  oracle/af_sw.c restates what
bwa-mem's scores optimise WITHOUT seeds, diagonals or X-drop: over all affine-gap local alignments (A 1, B 4, O 6, E 1)
maximise score + 5 for an alignment that starts at the read's first base + 5 for one that ends at its last (clipping an end
costs -L 5).  The oracle's record is one ungapped alignment, worth score + 5*[clip_l == 0] + 5*[clip_r == 0]:
  * the exhaustive optimum can never be smaller;
  * on reads without indels it must be EQUAL -- the seed-and-extend route with the greedy clip rule finds the optimum;
  * on reads with an indel it is larger by what a gap buys: the documented difference from bwa.
"""
import ctypes
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import ROOT

PARAMS = (1, 4, 6, 1, 5, 5)          # A, B, gap open, gap extend, clip5, clip3


@pytest.fixture(scope="module")
def sw():
    so = os.path.join(ROOT, "oracle", "libaf_sw.so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libaf_sw.so"], stdout=subprocess.DEVNULL, env=env)
    lib = ctypes.CDLL(so)
    lib.afo_sw_clip_objective.restype = ctypes.c_int
    lib.afo_sw_clip_objective.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_int] * 6 + \
                                         [ctypes.POINTER(ctypes.c_int32)] * 4
    return lib


def _objectives(sw, anchor, codes, lens, hits):
    """[(oracle objective, exhaustive objective, exhaustive (score, qb, qe, anchor end))] per oracle record."""
    anchor = np.ascontiguousarray(anchor, dtype=np.uint8)

    def one(h):
        rid, s, sc = int(h["read_id"]), int(h["score_strand"]) & 1, int(h["score_strand"]) >> 1
        L = int(lens[rid]) if lens is not None else codes.shape[1]
        r = codes[rid, :L]
        q = np.ascontiguousarray(np.where(r[::-1] < 4, 3 - r[::-1], 4) if s else r, dtype=np.uint8)
        out = [ctypes.c_int32() for _ in range(4)]
        obj = sw.afo_sw_clip_objective(q.ctypes.data, L, anchor.ctypes.data, len(anchor), *PARAMS, *out)
        mine = sc + PARAMS[4] * (int(h["clip_l"]) == 0) + PARAMS[5] * (int(h["clip_r"]) == 0)
        return mine, obj, tuple(o.value for o in out)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        return list(pool.map(one, hits))


def test_bundled_sample_records_are_optimal_under_bwa_scoring(sw, bundled):
    """All 1 261 anchored reads of the reference's bundled sample (wgsim: substitutions and sequencing errors, no indels in
    any anchored pair): the oracle's record attains the exhaustive optimum, and ends where the optimum ends."""
    from oracle import oracle
    hits = bundled["oracle_hits"]
    names = bundled["names1"]
    pat = re.compile(r".*_(\d+):(\d+):(\d+)_(\d+):(\d+):(\d+)_[0-9a-f]+(/[12])?$")
    assert all(int(pat.match(names[int(h["read_id"]) >> 1]).group(3)) == 0 and int(pat.match(names[int(h["read_id"]) >> 1]).group(6)) == 0
               for h in hits)                                  # wgsim's own bookkeeping: no indel in these pairs
    res = _objectives(sw, oracle.encode(bundled["anchor"]), bundled["codes"], None, hits)
    assert len(res) == 1261
    assert all(dp >= mine for mine, dp, _ in res)
    assert sum(dp == mine for mine, dp, _ in res) == len(res)
    L = bundled["read_len"]
    same_place = 0
    for h, (mine, dp, (sc, qb, qe, ae)) in zip(hits, res):
        same_place += (int(h["clip_l"]) == qb and L - int(h["clip_r"]) == qe and int(h["pos"]) + int(h["m_len"]) - 1 == ae)
    assert same_place >= 0.99 * len(res)                       # (equal-valued optima may sit elsewhere)


def test_substitution_only_reads_attain_the_optimum_and_indels_show_the_documented_gap(sw):
    from oracle import oracle
    rng = np.random.default_rng(8)
    G, L, n = 5000, 150, 1500
    anchor = rng.integers(0, 4, G).astype(np.uint8)
    anchor[[700, 701, 3000]] = 4
    clean, gapped = [], []
    for i in range(n):
        at = int(rng.integers(-60, G - L + 60))
        r = np.array([anchor[j] if 0 <= j < G and anchor[j] < 4 else rng.integers(0, 4) for j in range(at, at + L)], dtype=np.uint8)
        for _ in range(int(rng.integers(0, 5))):               # substitutions (and the odd N)
            r[int(rng.integers(0, L))] = rng.integers(0, 5)
        if i % 4 == 0:                                         # chimera: the right part is foreign
            cut = int(rng.integers(30, 120))
            r[cut:] = rng.integers(0, 4, L - cut)
        if i % 2:
            r = np.where(r[::-1] < 4, 3 - r[::-1], 4).astype(np.uint8)
        clean.append(r)
        g = r.copy()
        cut = int(rng.integers(40, 110))                       # one base deleted in the middle, one appended
        g = np.concatenate([g[:cut], g[cut + 1:], rng.integers(0, 4, 1).astype(np.uint8)])
        gapped.append(g)
    for reads, expect_equal in ((np.stack(clean), True), (np.stack(gapped), False)):
        hits = oracle.anchor_reads(anchor, reads, threads=4)
        assert len(hits) > 0.8 * n
        res = _objectives(sw, anchor, reads, None, hits)
        assert all(dp >= mine for mine, dp, _ in res)
        equal = sum(dp == mine for mine, dp, _ in res)
        if expect_equal:
            assert equal >= 0.995 * len(res), (equal, len(res))   # a gap can pay by chance next to a cluster of substitutions
        else:
            better = [dp - mine for mine, dp, _ in res if dp > mine]
            assert len(better) > 0.5 * len(res) and np.median(better) >= 20   # what gapped extension would add: the documented difference
