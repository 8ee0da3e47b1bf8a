"""numpy emulation of the seed-scan kernel's filter probes (test helper): which reads must the
kernel flag, bit for bit, given the index's filter words and hash."""
import numpy as np


def filter_hash(key, fmul, kp, nb):
    """af_filter_hash (csrc/af_common.h) on uint64 numpy arrays: (bucket as int64, fp3 as uint64)."""
    m32 = np.uint64(0xFFFFFFFF)
    lo = (key * np.uint64(fmul)) & m32
    b = ((lo * np.uint64(nb)) >> np.uint64(32)).astype(np.int64)
    fp3 = ((lo & np.uint64(0x3FE)) * np.uint64(0x00100401) + np.uint64(0x00100401)) & m32
    return b, fp3


def expected_flags(index, codes, lens=None, refine=False):
    """codes: (n_reads, stride) base codes with N/pad positions already replaced by the pad
    pattern (i.e. what the packed words hold).  Returns bool[n_reads].  refine=True models the
    optional neighbour test (af_neighbour_ok): a sample counts only if the k'-mer at p-H or p+H passes too."""
    info = index.info
    KP, S, nb, fm = info.kp, info.stride, info.n_buckets, info.filter_mul
    filt = index.filter_words().astype(np.uint64)
    n, stride = codes.shape
    Lmax = int(lens.max()) if lens is not None else stride
    flag = np.zeros(n, bool)
    W = (stride + 15) // 16
    if W > 16:
        W = (W + 3) & ~3                          # af_layout: long reads round W up to a multiple of 4
    pad = [(info.pad_byte >> (2 * k)) & 3 for k in range(4)]
    full = np.empty((n, 16 * W), dtype=np.uint64)
    full[:, :stride] = codes
    for i in range(stride, 16 * W):
        full[:, i] = pad[i & 3]
    if lens is not None:
        for i in range(stride):
            full[lens <= i, i] = pad[i & 3]
    H = (19 - KP + 1) // 2

    bloom = getattr(index, "bloom", False)

    def probe(p):
        key = np.zeros(n, np.uint64)
        for t in range(KP):
            key |= full[:, p + t] << np.uint64(2 * t)
        b, fp3 = filter_hash(key, fm, KP, nb)
        if bloom:                                  # af_bloom_probe: three bits of the bucket's word
            lo = (key * np.uint64(fm)) & np.uint64(0xFFFFFFFF)
            one = np.uint64(1)
            mask = (one << ((lo >> np.uint64(1)) & np.uint64(31))) | (one << ((lo >> np.uint64(6)) & np.uint64(31))) | (one << ((lo >> np.uint64(11)) & np.uint64(31)))
            return (~filt[b] & mask & np.uint64(0xFFFFFFFF)) == 0
        v = filt[b] ^ fp3
        return (((v - np.uint64(0x40100401)) & ~v & np.uint64(0xA0080200)) & np.uint64(0xFFFFFFFF)) != 0

    # sample grid of af_common.h (af_sample0 / af_nsamples), L = the batch's longest read
    p0 = 4 if KP == 12 else 19 - KP
    nsamp = (Lmax - 19 - p0 + S - 1) // S + 1 if Lmax >= 19 else 0
    for p in range(p0, p0 + nsamp * S, S):
        hit = probe(p)
        if refine and hit.any():                  # neighbour test (af_neighbour_ok)
            ok = np.zeros(n, bool)
            if p - H >= 0:
                ok |= probe(p - H)
            if p + H + KP <= 16 * W:
                ok |= probe(p + H)
            hit &= ok
        flag |= hit
    return flag
