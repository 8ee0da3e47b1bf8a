"""Pure-Python brute force of "Anchoring spec v1"  -- TEST INFRASTRUCTURE ONLY.

Walks EVERY diagonal of both orientations, with no index at all, so it pins the C oracle's
k-mer schedule on small cases.  Same semantics as oracle/af_oracle.c (which cites the
reference lines it stands in for: Anchored_Fusion.py:172,182,194).
"""

DEFAULT = dict(k=19, A=1, B=4, clip5=5, clip3=5, T=30, X=100)


def _extend(bits, h0, qlen, A, B, X):
    cur, mx, off, g = h0, h0, 0, -1
    for j, m in enumerate(bits):
        cur += A if m else -B
        if cur <= 0:
            break
        if cur > mx:
            mx, off = cur, j + 1
        if j + 1 == qlen:
            g = cur
        if mx - cur > X:
            break
    return mx, off, g


def diag_eval(q, a, d, P=DEFAULT):
    L, G, k = len(q), len(a), P["k"]
    m = [0 <= i + d < G and q[i] < 4 and q[i] == a[i + d] for i in range(L)]
    qb0 = next((i for i in range(L - k + 1) if all(m[i:i + k])), None)
    if qb0 is None:
        return None
    sc, qb, qe = k * P["A"], 0, L
    if qb0 > 0:
        n = min(qb0, qb0 + d)
        mx, off, g = _extend([m[qb0 - 1 - j] for j in range(n)], sc, qb0, P["A"], P["B"], P["X"])
        if g <= 0 or g <= mx - P["clip5"]:
            qb, sc = qb0 - off, mx
        else:
            qb, sc = 0, g
    qe0 = qb0 + k
    if qe0 < L:
        n = min(L - qe0, G - (qe0 + d))
        mx, off, g = _extend([m[qe0 + j] for j in range(n)], sc, L - qe0, P["A"], P["B"], P["X"])
        if g <= 0 or g <= mx - P["clip3"]:
            qe, sc = qe0 + off, mx
        else:
            qe, sc = L, g
    return sc, qb, qe


def anchor_read(r, a, P=DEFAULT):
    """r, a: sequences of base codes.  Returns (pos, clip_l, m_len, clip_r, strand, score) or None."""
    L, G = len(r), len(a)
    best = None
    for s in (0, 1):
        q = list(r) if s == 0 else [3 - c if c < 4 else 4 for c in reversed(r)]
        for d in range(-(L - P["k"]), G - P["k"] + 1):
            res = diag_eval(q, a, d, P)
            if res is None:
                continue
            sc, qb, qe = res
            if best is None or sc > best[5]:
                best = (qb + d + 1, qb, qe - qb, L - qe, s, sc)
    if best is None or best[5] < P["T"]:
        return None
    return best
