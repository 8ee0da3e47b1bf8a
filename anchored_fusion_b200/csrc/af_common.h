// af_common.h -- constants and hash functions shared by the host index builder and the
// sm_100a kernels.  Host and device MUST agree bit for bit on everything in this file.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/anchored_fusion.h"

#ifdef __CUDACC__
#define AF_HD __host__ __device__ __forceinline__
#else
#define AF_HD inline
#endif

// ---- shared-memory filter ----------------------------------------------------------------
// bucket word: three 10-bit fingerprint fields (0 = empty; stored fingerprints are odd, so
// never 0) and a 2-bit state field in bits 30..31: 01 = normal, 00 = overflowed (always hit).
// The probe tests all four fields for "== 0 after XOR" with one subtract and one LOP3.
static const uint32_t AF_F_ONES = 0x40100401u;   // low bit of each field
static const uint32_t AF_F_HIGH = 0xA0080200u;   // high bit of each field
static const uint32_t AF_F_REP = 0x00100401u;    // replicates a fingerprint into 3 fields
static const uint32_t AF_F_EMPTY = 0x40000000u;  // empty, normal bucket
static const int AF_F_SLOTS = 3;
static const uint32_t AF_MAX_BUCKETS = 50176;    // 196 KB of the 227 KB shared memory per CTA; the rest holds the fused kernel's queues
static const uint32_t AF_MIN_BUCKETS = 2048;

AF_HD uint32_t af_umulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// key -> bucket index and the fingerprint replicated into the three fields (state field 00).
// One 32-bit multiply: the bucket comes from the top bits of the product, the fingerprint
// from bits 1..9.  (Taking the fingerprint from the HIGH word of the 64-bit product looked
// natural but is 5x worse: keys that share a bucket differ by a lattice of deltas whose high
// words repeat.  Within a bucket the low bits of the product are free.)
AF_HD void af_filter_hash(uint32_t key, uint32_t fmul, uint32_t nb, uint32_t &bucket, uint32_t &fp3) {
    uint32_t lo = key * fmul;
    bucket = af_umulhi(lo, nb);
    fp3 = (lo & 0x3FEu) * AF_F_REP + AF_F_REP;  // fingerprint = (lo & 0x3FE) + 1, odd, 1..1023
}

// nonzero iff some field of the bucket equals the fingerprint, or the bucket overflowed
AF_HD uint32_t af_filter_test(uint32_t word, uint32_t fp3) {
    uint32_t v = word ^ fp3;
    return (v - AF_F_ONES) & ~v & AF_F_HIGH;
}

AF_HD uint32_t af_funnel_r(uint32_t lo, uint32_t hi, int sh) {   // bits [sh, sh+32) of hi:lo, 0 <= sh < 32
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// ---- the seed-scan probe sequence of one read (shared by k_seed_scan and its host twin) ------
// k'-mer starting at base P (compile-time) of the read in registers w[OFF..OFF+W)
template <int W, int KP, int OFF, int P, int NW>
AF_HD uint32_t af_kmer_at(const uint32_t (&w)[NW]) {
    constexpr uint32_t KMASK = (1u << (2 * KP)) - 1u;
    constexpr int o = 2 * P, wi = o >> 5, sh = o & 31;
    if (sh + 2 * KP <= 32) return (w[OFF + wi] >> sh) & KMASK;
    return af_funnel_r(w[OFF + wi], w[OFF + (wi + 1 < W ? wi + 1 : wi)], sh) & KMASK;
}

AF_HD uint32_t af_filter_probe(uint32_t key, const uint32_t *filt, uint32_t fmul, uint32_t nb) {
    uint32_t b, fp3;
    af_filter_hash(key, fmul, nb, b, fp3);
    const uint32_t v = filt[b] ^ fp3;
    return (v - AF_F_ONES) & ~v;                                             // & AF_F_HIGH != 0  <=>  hit
}

// Blocked-Bloom variant of the same filter, for anchors whose k'-mers overflow the 3-slot buckets (beyond ~12 kb the
// buckets turn "always hit" one by one and the scan flags most reads): a bucket word is 32 Bloom bits, a key sets three
// of them, chosen by bits 1..15 of the same product that picks the bucket.  No overflow state, so the false-positive
// rate degrades smoothly (40 kb anchor: ~0.3 % per probe against "always hit" in most buckets).  Three more integer
// instructions per probe than the fingerprint test, which is why the fingerprint buckets stay the default for anchors
// they can hold.  Returns AF_F_HIGH on a hit so that the callers' accumulate-and-test is the same.
AF_HD uint32_t af_bloom_mask(uint32_t lo) {
    return (1u << ((lo >> 1) & 31u)) | (1u << ((lo >> 6) & 31u)) | (1u << ((lo >> 11) & 31u));
}
AF_HD uint32_t af_bloom_probe(uint32_t key, const uint32_t *filt, uint32_t fmul, uint32_t nb) {
    const uint32_t lo = key * fmul;
    return (~filt[af_umulhi(lo, nb)] & af_bloom_mask(lo)) ? 0u : AF_F_HIGH;
}

// The rarely taken second look at a sample that passed the filter (compile-time position P):
// a true >= k match [q, q+k) around the sample also holds the k'-mer at P-H or the one at P+H,
// H = ceil((k-k')/2) -- left slack a = P-q and right slack b add up to k-k', so a >= H or b >= H.
// Chance k'-mer hits, which are ~96 % of what the plain filter flags, pass with probability ~0.3 %.
template <int W, int KP, int OFF, int P, int NW>
AF_HD bool af_neighbour_ok(const uint32_t (&w)[NW], const uint32_t *filt, uint32_t fmul, uint32_t nb) {
    constexpr int H = (19 - KP + 1) / 2;
    bool ok = false;
    if (P - H >= 0) ok = (af_filter_probe(af_kmer_at<W, KP, OFF, (P - H >= 0 ? P - H : 0)>(w), filt, fmul, nb) & AF_F_HIGH) != 0;
    if (P + H + KP <= 16 * W) {
        if (!ok) ok = (af_filter_probe(af_kmer_at<W, KP, OFF, (P + H + KP <= 16 * W ? P + H : 0)>(w), filt, fmul, nb) & AF_F_HIGH) != 0;
    }
    return ok;
}

// ---- sample positions -----------------------------------------------------------------------
// A read of L bases is sampled at p_j = P0 + j*S, S = k - k' + 1 (k = 19): a window of k matching bases [q, q+k),
// 0 <= q <= L-k, contains the sample that starts in [q, q+S-1] (it ends at p_j + k' <= q + k), so the grid needs
// P0 <= S-1 and a last sample at or beyond L-k: ceil((L-k-P0)/S) + 1 samples.  P0 = 0 takes 18 samples for 150 bases,
// any P0 in 3..7 takes 17 (11 instead of 12 for 101 bases).  For k' = 12, P0 = 4 puts the samples at bit offsets 8 and
// 24 of the packed words: every other k'-mer is one shift of one word (8 + 24 = 32 bits, no mask, no funnel).
// Every stage (scan, verify, extend, tail, host twin) takes the grid from these two functions.
AF_HD constexpr int af_sample0(int kp) { return kp == 12 ? 4 : 19 - kp; }
AF_HD constexpr int af_lmin(int W) { return 16 * (W > 16 ? W - 4 : W - 1) + 1; }   // shortest longest-read of a batch laid out with W words (af_layout rounds W to 20, 24, 28, 32 beyond 16)
AF_HD constexpr int af_nsamples(int L, int kp) { return L >= 19 ? (L - 19 - af_sample0(kp) + (20 - kp) - 1) / (20 - kp) + 1 : 0; }

template <int W, int KP, int OFF, int J, int NP, int NPMIN, bool REFINE, bool BLOOM, int NW>
AF_HD void af_scan_sample(const uint32_t (&w)[NW], int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                          uint32_t &acc) {
    if constexpr (J < NP) {
        constexpr int S = 20 - KP, PJ = af_sample0(KP) + J * S;
        // samples past NPMIN exist only for the longer reads this W can hold: a warp-uniform predicate on the
        // load alone (the arithmetic around it stays branch-free), so that a sample the batch does not have costs
        // no shared-memory wavefronts
        uint32_t t;
        if constexpr (BLOOM) {
            const uint32_t lo = af_kmer_at<W, KP, OFF, PJ>(w) * fmul;
            uint32_t word = 0u;
            if (J < NPMIN || J < nprobe) word = filt[af_umulhi(lo, nb)];
            t = (~word & af_bloom_mask(lo)) ? 0u : AF_F_HIGH;
        } else {
            uint32_t b, fp3;
            af_filter_hash(af_kmer_at<W, KP, OFF, PJ>(w), fmul, nb, b, fp3);
            uint32_t word = ~fp3;                                            // no field can match
            if (J < NPMIN || J < nprobe) word = filt[b];
            const uint32_t v = word ^ fp3;
            t = (v - AF_F_ONES) & ~v;
        }
        if constexpr (REFINE) {
            if (t & AF_F_HIGH) {                                             // rare: ~0.14 % of the samples
                if (!af_neighbour_ok<W, KP, OFF, PJ>(w, filt, fmul, nb)) t = 0u;
            }
        }
        acc |= t;
        af_scan_sample<W, KP, OFF, J + 1, NP, NPMIN, REFINE, BLOOM>(w, nprobe, filt, fmul, nb, acc);
    }
}

// All sample positions of one read held in registers w[OFF..OFF+W).  Fully unrolled (compile-time
// recursion): every shift and word index is a constant, the first NPMIN samples exist for every
// read length this W can hold, the last few are masked by a warp-uniform select.  With REFINE a
// sample only counts if its neighbour test passes too (af_neighbour_ok): that removes nearly all
// chance hits (362 k -> 16 k flagged reads per 10 M pairs) but the rarely taken branch per sample
// splits the probe sequence into 36 basic blocks and the kernel runs 2x slower (measured), so the
// production scan uses REFINE = false and leaves the false positives to k_verify.
template <int W, int KP, int OFF, int NW, bool REFINE = false, bool BLOOM = false>
AF_HD uint32_t af_scan_read(const uint32_t (&w)[NW], int nprobe, const uint32_t *filt, uint32_t fmul,
                              uint32_t nb) {
    constexpr int NP = af_nsamples(16 * W, KP);                              // samples when L == 16 W
    constexpr int NPMIN = af_nsamples(af_lmin(W), KP);                       // af_lmin: shortest L with this W
    uint32_t acc = 0;
    af_scan_sample<W, KP, OFF, 0, NP, NPMIN, REFINE, BLOOM>(w, nprobe, filt, fmul, nb, acc);
    return acc & AF_F_HIGH;
}

// Both reads of a pair (w[0..W) and w[W..2W)), as the production scan runs them: the samples every read length of
// this W has are branch-free; the last few, which only the longer lengths have, sit behind warp-uniform branches
// shared by the two reads -- a sample the batch lacks costs neither arithmetic nor shared-memory wavefronts.
// Same flags as af_scan_read on each read.
template <int W, int KP, int J, int NP, bool BLOOM, int NW>
AF_HD void af_scan_pair_tail(const uint32_t (&w)[NW], int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                             uint32_t &acc1, uint32_t &acc2) {
    if constexpr (J < NP) {
        if (J < nprobe) {
            af_scan_sample<W, KP, 0, J, J + 1, J + 1, false, BLOOM>(w, nprobe, filt, fmul, nb, acc1);
            af_scan_sample<W, KP, W, J, J + 1, J + 1, false, BLOOM>(w, nprobe, filt, fmul, nb, acc2);
            af_scan_pair_tail<W, KP, J + 1, NP, BLOOM>(w, nprobe, filt, fmul, nb, acc1, acc2);
        }
    }
}
template <int W, int KP, int NW, bool BLOOM = false>
AF_HD void af_scan_pair(const uint32_t (&w)[NW], int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                        uint32_t &a1, uint32_t &a2) {
    constexpr int NP = af_nsamples(16 * W, KP), NPMIN = af_nsamples(af_lmin(W), KP);
    uint32_t acc1 = 0, acc2 = 0;
    af_scan_sample<W, KP, 0, 0, NPMIN, NPMIN, false, BLOOM>(w, nprobe, filt, fmul, nb, acc1);
    af_scan_sample<W, KP, W, 0, NPMIN, NPMIN, false, BLOOM>(w, nprobe, filt, fmul, nb, acc2);
    af_scan_pair_tail<W, KP, NPMIN, NP, BLOOM>(w, nprobe, filt, fmul, nb, acc1, acc2);
    a1 = acc1 & AF_F_HIGH; a2 = acc2 & AF_F_HIGH;
}

// ---- exact table (global memory): open addressing, duplicates allowed ---------------------
static const uint32_t AF_T_EMPTY = 0xFFFFFFFFu;  // k' <= 15 keeps real keys below this
AF_HD uint32_t af_table_hash(uint32_t key, uint32_t tmask) {
    uint32_t h = key * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    return h & tmask;
}

// ---- synthetic generator ------------------------------------------------------------------
AF_HD uint64_t af_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
AF_HD uint32_t af_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    return x ^ (x >> 16);
}
// base x of the random reference: 32 bases per 64-bit hash word
AF_HD uint32_t af_ref_base(uint64_t seed, int64_t x) {
    uint64_t w = af_mix64(seed ^ ((uint64_t)(x >> 5) * 0xD6E8FEB86659FD93ull));
    return (uint32_t)(w >> (2 * (x & 31))) & 3u;
}

struct af_frag {       // one synthetic fragment, a pure function of (seed, pair index)
    int64_t start;     // reference start of the (left part of the) fragment
    int64_t start2;    // reference start of the right part when chimeric
    int32_t len;       // fragment length
    int32_t junction;  // chimeric: fragment offset where the right part begins; else len
    int32_t flip;      // 1: the fragment is the reverse complement of the reference interval
    uint32_t rseed;    // per-pair seed of the substitution / N stream
};

AF_HD af_frag af_make_frag(const af_synth_t &s, int64_t pair) {
    af_frag f;
    uint64_t h0 = af_mix64(s.seed * 0x9E3779B97F4A7C15ull + (uint64_t)pair);
    uint64_t h1 = af_mix64(h0), h2 = af_mix64(h1), h3 = af_mix64(h2);
    int64_t u = (int64_t)(h0 & 0xFFFF) + (int64_t)((h0 >> 16) & 0xFFFF) + (int64_t)((h0 >> 32) & 0xFFFF) +
                (int64_t)((h0 >> 48) & 0xFFFF) - 131070;       // ~N(0, 37837)
    int64_t len = (int64_t)s.frag_mean + (u * (int64_t)s.frag_sd) / 37837;
    if (len < s.read_len) len = s.read_len;
    if (len > s.ref_len / 2) len = s.ref_len / 2;
    f.len = (int32_t)len;
    f.flip = (int32_t)((h1 >> 63) & 1);
    f.rseed = (uint32_t)(h3 >> 32);
    f.junction = f.len;
    f.start = (int64_t)(h1 % (uint64_t)(s.ref_len - len));
    f.start2 = 0;
    bool fusion = (uint32_t)(h2 % 1000000ull) < s.fusion_ppm;
    if (fusion && s.anchor_len > 64 && f.len > 60) {
        // left part ends somewhere inside the anchor, right part starts elsewhere
        int32_t jx = 30 + (int32_t)((h2 >> 20) % (uint64_t)(f.len - 59));   // 30 .. len-30
        int64_t aend = s.anchor_start + 32 + (int64_t)((h2 >> 40) % (uint64_t)(s.anchor_len - 63));
        int64_t st = aend - jx;
        if (st < 0) st = 0;
        f.start = st;
        f.junction = jx;
        f.start2 = (int64_t)(h3 % (uint64_t)(s.ref_len - len));
    }
    return f;
}

// base i (0..len) of the fragment in fragment orientation
AF_HD uint32_t af_frag_base(const af_synth_t &s, const af_frag &f, int32_t i) {
    int32_t j = f.flip ? f.len - 1 - i : i;   // offset in reference orientation
    int64_t x = (j < f.junction) ? f.start + j : f.start2 + (j - f.junction);
    uint32_t b = af_ref_base(s.seed, x);
    return f.flip ? 3u - b : b;
}

// base i of mate m (0/1) of the pair, with substitutions and Ns; returns 0..4
AF_HD uint32_t af_read_base(const af_synth_t &s, const af_frag &f, int m, int32_t i) {
    uint32_t b = m == 0 ? af_frag_base(s, f, i) : 3u - af_frag_base(s, f, f.len - 1 - i);
    uint32_t e = af_mix32(f.rseed + (uint32_t)(m * 4096 + i) * 0x9E3779B1u);
    if (s.sub_ppm && (e % 1000003u) < s.sub_ppm) b = (b + 1u + ((e >> 24) % 3u)) & 3u;
    if (s.n_ppm) {
        uint32_t e2 = af_mix32(e ^ 0xA5A5A5A5u);
        if ((e2 % 1000003u) < s.n_ppm) b = 4u;
    }
    return b;
}

// ---- error reporting and host-side objects ------------------------------------------------
void af_set_error(const char *fmt, ...);

struct af_index {
    af_params_t P;
    int32_t kp, stride, G;
    std::vector<uint8_t> codes;     // anchor base codes 0..4
    uint32_t fmul, nb;
    std::vector<uint32_t> filter;   // nb bucket words
    uint32_t fmul2, nb2;
    std::vector<uint32_t> filter2;  // half-size copy (own multiplier) for k_verify, which shares the SM with read staging
    uint32_t tmask;
    std::vector<uint32_t> table;    // (tmask+1) x {key, value}; value = strand<<31 | anchor pos
    std::vector<uint32_t> member;   // 4^kp-bit bitmap: bit key set iff key is an anchor k'-mer
    std::vector<uint32_t> apk[2];   // anchor, 2 bit/base: [0] forward, [1] reverse complement (N packed as A)
    int32_t anchor_has_n;
    int32_t n_keys, n_entries, n_overflow, pad_byte;
    int32_t bloom = 0;              // filter holds Bloom bits (long anchor) instead of fingerprint buckets
};

// shared-memory filter over `keys` (distinct k'-mers) with nbk buckets: tries the first n_muls multipliers and keeps
// the one with the fewest false positives on 2^log2_probes pseudo-random keys (af_host.cpp)
void af_filter_pick(const std::vector<uint32_t> &keys, uint32_t kmask, uint32_t nbk, int n_muls, int log2_probes,
                    uint32_t &mul_out, std::vector<uint32_t> &filt_out, int32_t *ov_out);

// the genome pass's input on the host (af_genome_host.cpp): the concatenation, 2 bit/base (an N is stored as a
// position-hashed pseudo-random base) + 1 bit/base N map, and the contigs' places in it
struct af_genome_host {
    std::vector<uint32_t> pk, nm;
    int64_t n = 0;
    std::vector<std::string> names;
    std::vector<int64_t> starts, lens;
};
int af_genome_host_from_fasta(const char *path, af_genome_host &out);
int af_genome_host_from_contigs(const char *const *names, const char *const *seqs, const int64_t *lens, int32_t n, af_genome_host &out);

static inline uint8_t af_code_of(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// ---- multi-GPU hit exchange (af_exchange.cu; the appending kernel is k_hit_scatter) --------
static const int AF_LOG_HEADER_BYTES = 128;
struct af_log_header { unsigned long long tail; uint32_t status, n_batches; };
// what k_hit_scatter needs to append one batch to log (rank, slot) on every rank; by value
struct af_sink {
    int32_t world;
    uint32_t log_cap;                // records per region, markers included
    unsigned long long *state;       // writer side, local: [0..1] tail by batch parity, [2..3] n_batches<<32 | status
    uint32_t seq;                    // batches appended to this log since the last reset
    unsigned long long pair_base;
    char *region[AF_MAX_PEERS];      // region (rank, slot) inside each rank's buffer (own rank: local pointer)
};
struct af_exchange {
    int device, rank, world, n_slots;
    int64_t log_cap;
    size_t region_bytes, total_bytes;
    char *local;                     // world x n_slots regions
    char *peer[AF_MAX_PEERS];        // peer[r] = rank r's buffer mapped here (peer[rank] = local)
    unsigned long long *state;       // n_slots x 4
    uint32_t seq[64];                // per slot: batches appended since the last reset (host side)
    bool connected;
};
int af_exchange_sink(af_exchange *ex, int slot, int64_t pair_base, af_sink *out);
