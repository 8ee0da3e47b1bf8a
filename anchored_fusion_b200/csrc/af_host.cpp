// af_host.cpp -- host side of libafb200: anchor index build, batch layout, 2-bit packer,
// synthetic-read generator (host twin of the device generator), error plumbing.
// No CUDA here; see af_kernels.cu for the device side.
#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "af_common.h"

static thread_local char g_err[512] = "";

void af_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *af_last_error(void) { return g_err; }
extern "C" int af_abi_version(void) { return AF_ABI_VERSION; }
extern "C" void af_default_params(af_params_t *p) {
    p->k = 19; p->A = 1; p->B = 4; p->clip5 = 5; p->clip3 = 5; p->T = 30; p->X = 100;
}

// ------------------------------------------------------------------------------------------
// anchor index  (replaces `bwa index`, Anchored_Fusion.py:167-172)
// ------------------------------------------------------------------------------------------
static int build_filter(const std::vector<uint32_t> &keys, uint32_t fmul, uint32_t nb, std::vector<uint32_t> &out) {
    out.assign(nb, AF_F_EMPTY);
    int overflow = 0;
    for (uint32_t key : keys) {
        uint32_t b, fp3;
        af_filter_hash(key, fmul, nb, b, fp3);
        uint32_t fp = fp3 & 0x3FFu, w = out[b];
        if (!(w & AF_F_EMPTY)) continue;  // already always-hit
        bool placed = false;
        for (int s = 0; s < AF_F_SLOTS && !placed; s++) {
            uint32_t cur = (w >> (10 * s)) & 0x3FFu;
            if (cur == fp) placed = true;
            else if (cur == 0) { w |= fp << (10 * s); placed = true; }
        }
        if (!placed) { w = 0; overflow++; }  // state 00: every probe of this bucket hits
        out[b] = w;
    }
    return overflow;
}

// The filter hash is one multiply; how often foreign k'-mers collide with stored ones depends on
// the multiplier (the colliding key differences form a lattice).  Try a few and keep the one
// with the lowest measured false-positive rate on pseudo-random probe keys.
void af_filter_pick(const std::vector<uint32_t> &keys, uint32_t kmask, uint32_t nbk, int n_muls, int log2_probes,
                    uint32_t &mul_out, std::vector<uint32_t> &filt_out, int32_t *ov_out) {
    static const uint32_t muls[] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du,
                                    0xFD7046C5u, 0xB55A4F09u, 0x8DA6B343u, 0xD8163841u, 0xCB1AB31Fu, 0x9C06FAF5u,
                                    0x2545F491u, 0x6C8E9CF5u, 0xE7037ED1u, 0xA3B19535u};
    std::vector<uint32_t> cand;
    long best = -1;
    for (int mi = 0; mi < n_muls && mi < 16; mi++) {
        const uint32_t m = muls[mi];
        int ov = build_filter(keys, m, nbk, cand);
        long fp = 0;
        uint32_t x = 0x12345678u;
        for (int t = 0; t < (1 << log2_probes); t++) {
            x = af_mix32(x + 0x9E3779B9u);
            uint32_t b, fp3;
            af_filter_hash(x & kmask, m, nbk, b, fp3);
            fp += af_filter_test(cand[b], fp3) != 0;
        }
        if (best < 0 || fp < best) { best = fp; mul_out = m; filt_out = cand; if (ov_out) *ov_out = ov; }
    }
}

extern "C" int af_index_build(const char *anchor, int64_t len, const af_params_t *params, int32_t kp,
                              af_index_t **out) {
    if (!anchor || !out || len <= 0 || len >= (1ll << 28)) { af_set_error("af_index_build: bad anchor"); return AF_ERR_ARG; }
    af_params_t P;
    if (params) P = *params; else af_default_params(&P);
    if (kp == 0) kp = 12;
    // The kernels are built for bwa-mem's seed length: they sample a read's k'-mers at stride 20 - k'
    // (compile-time), which finds every exact match of >= 19 bases.  Scores, clip penalties, T and X are free.
    if (P.k != 19 || (kp != 12 && kp != 13) || P.A <= 0 || P.B < 0 || P.X < 0) {
        af_set_error("af_index_build: unsupported parameters (k=%d must be 19, kp=%d must be 12 or 13)", P.k, kp);
        return AF_ERR_ARG;
    }
    // a record stores score*2 + strand in 16 bits; the best possible score is A * read length
    if ((int64_t)2 * P.A * AF_MAX_READ_LEN + 1 > 65535 || P.B > 32767 || P.clip5 < 0 || P.clip3 < 0) {
        af_set_error("af_index_build: match score A=%d too large (2*A*%d+1 must fit 16 bits), or negative clip penalties", P.A, AF_MAX_READ_LEN);
        return AF_ERR_ARG;
    }
    af_index *idx = new af_index();
    idx->P = P;
    idx->kp = kp;
    idx->stride = 20 - kp;
    idx->G = (int32_t)len;
    idx->codes.resize((size_t)len);
    for (int64_t i = 0; i < len; i++) idx->codes[(size_t)i] = af_code_of(anchor[i]);

    // 2-bit packed anchor, both orientations, for word-parallel run checks
    idx->anchor_has_n = 0;
    for (int o = 0; o < 2; o++) {
        idx->apk[o].assign((size_t)(idx->G + 15) / 16 + 4, 0u);
        for (int32_t i = 0; i < idx->G; i++) {
            uint8_t c = o == 0 ? idx->codes[(size_t)i] : idx->codes[(size_t)(idx->G - 1 - i)];
            if (c >= 4) { idx->anchor_has_n = 1; continue; }
            if (o) c = (uint8_t)(3 - c);
            idx->apk[o][(size_t)i >> 4] |= (uint32_t)c << (2 * (i & 15));
        }
    }

    struct Ent { uint32_t key, val; };
    std::vector<Ent> ents;
    int run = 0;
    const uint32_t kmask = (1u << (2 * kp)) - 1;
    uint32_t kf = 0, kr = 0;  // forward k'-mer (base t at bits 2t) and its reverse complement
    for (int32_t i = 0; i < idx->G; i++) {
        uint8_t c = idx->codes[(size_t)i];
        if (c < 4) {
            kf = (kf >> 2) | ((uint32_t)c << (2 * (kp - 1)));
            kr = ((kr << 2) | (uint32_t)(3 - c)) & kmask;
            run++;
        } else { run = 0; kf = kr = 0; }
        if (run >= kp) {
            uint32_t j = (uint32_t)(i - kp + 1);
            ents.push_back({kf, j});
            ents.push_back({kr, (1u << 31) | j});
        }
    }
    idx->n_entries = (int32_t)ents.size();

    uint32_t slots = 1024;
    while (slots < 4 * ents.size()) slots <<= 1;
    idx->tmask = slots - 1;
    idx->table.assign((size_t)slots * 2, AF_T_EMPTY);
    for (const Ent &e : ents) {
        uint32_t s = af_table_hash(e.key, idx->tmask);
        while (idx->table[(size_t)s * 2] != AF_T_EMPTY) s = (s + 1) & idx->tmask;
        idx->table[(size_t)s * 2] = e.key;
        idx->table[(size_t)s * 2 + 1] = e.val;
    }

    std::vector<uint32_t> keys;
    keys.reserve(ents.size());
    for (const Ent &e : ents) keys.push_back(e.key);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    idx->n_keys = (int32_t)keys.size();
    idx->member.assign((size_t)1 << (2 * kp - 5), 0u);   // exact membership, one bit per possible k'-mer
    for (uint32_t key : keys) idx->member[key >> 5] |= 1u << (key & 31);

    uint64_t want = (uint64_t)keys.size() * 4;
    uint32_t nb = (uint32_t)std::min<uint64_t>(AF_MAX_BUCKETS, std::max<uint64_t>(AF_MIN_BUCKETS, want));
    nb = (nb + 31u) & ~31u;
    idx->nb = nb;
    auto pick = [&](uint32_t nbk, uint32_t &mul_out, std::vector<uint32_t> &filt_out, int32_t *ov_out) {
        af_filter_pick(keys, kmask, nbk, 16, 18, mul_out, filt_out, ov_out);
    };
    pick(nb, idx->fmul, idx->filter, &idx->n_overflow);
    // Long anchor: so many 3-slot buckets overflowed into "always hit" that the scan would flag most reads.  The same
    // words then hold a blocked Bloom filter instead (af_bloom_probe): three bits per key in the bucket's word, the
    // multiplier again chosen by measured false positives.  n_overflow keeps reporting what the buckets did.
    idx->bloom = 0;
    // (measured on B200, profiles/r02m: the Bloom probe costs the scan 45 % -- 0.27 instead of 0.19 ms per 10 M pairs -- and
    // its flagged reads take the exact-bitmap verify kernel, so it pays once more than ~4 % of the buckets have overflowed:
    // 40 kb anchor 1.86 -> 1.04 ms per step, 100 kb 4.34 -> 2.34 ms; at 20 kb the overflowing buckets are still faster)
    if ((long long)idx->n_overflow * 25 > (long long)nb && !getenv("AF_NO_BLOOM")) {
        static const uint32_t bmuls[] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu};
        long best = -1;
        std::vector<uint32_t> bits;
        for (uint32_t m : bmuls) {
            bits.assign(nb, 0u);
            for (uint32_t key : keys) { const uint32_t lo = key * m; bits[af_umulhi(lo, nb)] |= af_bloom_mask(lo); }
            long fp = 0;
            uint32_t x = 0x12345678u;
            for (int t = 0; t < (1 << 18); t++) {
                x = af_mix32(x + 0x9E3779B9u);
                fp += af_bloom_probe(x & kmask, bits.data(), m, nb) != 0;
            }
            if (best < 0 || fp < best) { best = fp; idx->fmul = m; idx->filter = bits; }
        }
        idx->bloom = 1;
    }
    idx->nb2 = ((nb / 2) + 31u) & ~31u;
    pick(idx->nb2, idx->fmul2, idx->filter2, nullptr);

    // 4-base pad pattern whose k'-mers (all 4 phases) are absent from the anchor, so padded
    // tails and N positions of a read never light the filter up by themselves.
    idx->pad_byte = 0xE4;
    for (int t = 0; t < 256; t++) {
        int b = (0xE4 + t * 37) & 0xFF;
        bool clean = true;
        for (int ph = 0; ph < 4 && clean; ph++) {
            uint32_t key = 0;
            for (int i = 0; i < kp; i++) key |= (uint32_t)((b >> (2 * ((i + ph) & 3))) & 3) << (2 * i);
            clean = !std::binary_search(keys.begin(), keys.end(), key);
        }
        if (clean) { idx->pad_byte = b; break; }
    }
    *out = idx;
    return AF_OK;
}

extern "C" void af_index_free(af_index_t *idx) { delete idx; }

extern "C" int af_index_info(const af_index_t *idx, af_index_info_t *info) {
    if (!idx || !info) { af_set_error("af_index_info: null"); return AF_ERR_ARG; }
    info->anchor_len = idx->G; info->k = idx->P.k; info->kp = idx->kp; info->stride = idx->stride;
    info->n_keys = idx->n_keys; info->n_entries = idx->n_entries; info->n_buckets = (int32_t)idx->nb;
    info->n_overflow = idx->n_overflow; info->table_slots = (int32_t)(idx->tmask + 1);
    info->filter_mul = idx->fmul; info->pad_byte = idx->pad_byte;
    return AF_OK;
}
extern "C" int af_index_filter_kind(const af_index_t *idx) { return idx ? idx->bloom : 0; }
extern "C" const uint32_t *af_index_filter(const af_index_t *idx) { return idx ? idx->filter.data() : nullptr; }
extern "C" const uint32_t *af_index_table(const af_index_t *idx) { return idx ? idx->table.data() : nullptr; }

// ------------------------------------------------------------------------------------------
// batch layout and the 2-bit packer  (replaces kseq/zlib ingest inside bwa, Anchored_Fusion.py:182)
// ------------------------------------------------------------------------------------------
extern "C" int af_layout(int32_t max_read_len, int64_t n_pairs, af_layout_t *out) {
    if (!out || max_read_len <= 0 || max_read_len > AF_MAX_READ_LEN || n_pairs < 0) {
        af_set_error("af_layout: max_read_len must be 1..%d", AF_MAX_READ_LEN);
        return AF_ERR_ARG;
    }
    out->max_read_len = max_read_len;
    out->words_per_read = (max_read_len + 15) / 16;
    if (out->words_per_read > 16) out->words_per_read = (out->words_per_read + 3) & ~3;   // long reads: the scan is built for W = 20, 24, 28, 32
    out->quads_per_pair = (2 * out->words_per_read + 3) / 4;
    out->reserved = 0;
    out->n_pairs = n_pairs;
    out->n_tiles = (n_pairs + AF_TILE_PAIRS - 1) / AF_TILE_PAIRS;
    out->packed_bytes = out->n_tiles * out->quads_per_pair * 512;
    return AF_OK;
}

// ---- wire format (include/anchored_fusion.h): a pair's 4 L bits back to back, word-interleaved per tile ----------
static inline int wire_words(int32_t L) { return (4 * L + 31) / 32; }
static inline uint32_t pad_word(int32_t pad_byte) {
    uint32_t padw = 0;
    for (int i = 0; i < 16; i++) padw |= (uint32_t)((pad_byte >> (2 * (i & 3))) & 3) << (2 * i);
    return padw;
}
extern "C" int64_t af_wire_bytes(int32_t max_read_len, int64_t n_pairs) {
    if (max_read_len <= 0 || max_read_len > AF_MAX_READ_LEN || n_pairs < 0) return -1;
    return ((n_pairs + AF_TILE_PAIRS - 1) / AF_TILE_PAIRS) * (int64_t)wire_words(max_read_len) * 128;
}
extern "C" int af_wire_from_packed(const void *packed, int32_t L, int64_t n_pairs, void *wire_out) {
    af_layout_t lay;
    int rc = af_layout(L, n_pairs, &lay);
    if (rc) return rc;
    if ((!packed || !wire_out) && lay.n_tiles) { af_set_error("af_wire_from_packed: null"); return AF_ERR_ARG; }
    const int W = lay.words_per_read, Q = lay.quads_per_pair, NW = wire_words(L);
    const uint32_t *in = (const uint32_t *)packed;
    uint32_t *out = (uint32_t *)wire_out;
    for (int64_t tile = 0; tile < lay.n_tiles; tile++)
        for (int lane = 0; lane < 32; lane++) {
            const uint32_t *pb = in + (tile * Q * 32 + lane) * 4;
            uint32_t o[2 * AF_MAX_READ_LEN / 16 + 1];
            memset(o, 0, sizeof(o));
            int bit = 0;
            for (int m = 0; m < 2; m++)
                for (int t = 0; t < W; t++) {
                    const int wi = m * W + t, nb = std::max(0, std::min(32, 2 * L - 32 * t));
                    if (nb == 0) continue;
                    uint32_t v = pb[(wi >> 2) * 128 + (wi & 3)];
                    if (nb < 32) v &= (1u << nb) - 1u;
                    const int a = bit >> 5, sh = bit & 31;
                    o[a] |= v << sh;
                    if (sh && sh + nb > 32) o[a + 1] |= v >> (32 - sh);
                    bit += nb;
                }
            for (int j = 0; j < NW; j++) out[(tile * NW + j) * 32 + lane] = o[j];
        }
    return AF_OK;
}
extern "C" int af_wire_to_packed(const void *wire, int32_t L, int64_t n_pairs, int32_t pad_byte, void *packed_out) {
    af_layout_t lay;
    int rc = af_layout(L, n_pairs, &lay);
    if (rc) return rc;
    if ((!wire || !packed_out) && lay.n_tiles) { af_set_error("af_wire_to_packed: null"); return AF_ERR_ARG; }
    const int W = lay.words_per_read, Q = lay.quads_per_pair, NW = wire_words(L);
    const uint32_t padw = pad_word(pad_byte);
    const uint32_t *in = (const uint32_t *)wire;
    uint32_t *out = (uint32_t *)packed_out;
    for (int64_t tile = 0; tile < lay.n_tiles; tile++)
        for (int lane = 0; lane < 32; lane++) {
            uint32_t *pb = out + (tile * Q * 32 + lane) * 4;
            for (int wi = 0; wi < 4 * Q; wi++) {
                uint32_t word = 0;
                if (wi < 2 * W) {
                    const int m = wi >= W, t = wi - m * W, nb = std::min(32, 2 * L - 32 * t), bit = m * 2 * L + 32 * t;
                    word = padw;                            // nb <= 0: a word past the read (W is rounded up for long reads)
                    if (nb > 0) {
                        const int a = bit >> 5, sh = bit & 31;
                        const uint32_t lo = in[(tile * NW + a) * 32 + lane], hi = a + 1 < NW ? in[(tile * NW + a + 1) * 32 + lane] : 0u;
                        const uint32_t v = af_funnel_r(lo, hi, sh), mask = nb >= 32 ? 0xFFFFFFFFu : (1u << nb) - 1u;
                        word = (v & mask) | (padw & ~mask);
                    }
                }
                pb[(wi >> 2) * 128 + (wi & 3)] = word;
            }
        }
    return AF_OK;
}

struct SeqRef { const char *p; int32_t len; };

// One mate's words of every pair of a batch (mate 0 also writes the zero words that pad a pair to
// whole quads).  The two mates touch disjoint 32-bit words, so the two sides of a FASTQ pair can
// be packed by two threads at once.  N positions and the tail past a read keep the pad pattern.
struct PackSide {
    std::vector<uint32_t> nids, nmask;   // read ids (2p+m) holding an N, and their 256-bit masks
    int32_t ulen = -1;                   // common length, -2 once lengths differ
    int rc = AF_OK;
    std::string err;
};

static const uint8_t *code_lut() {
    // built once, thread-safely (function-local static): the two mates are packed by two threads
    static const std::array<uint8_t, 256> lut = [] {
        std::array<uint8_t, 256> t{};
        for (int c = 0; c < 256; c++) t[(size_t)c] = af_code_of((char)c);
        return t;
    }();
    return lut.data();
}

// 16 bases -> one packed word.  ACGT (either case) map to 0..3 through ((c >> 1) ^ (c >> 2)) & 3; everything
// else is N: the word keeps the pad pattern there and the position is reported in nbits.
static inline uint32_t pack16_scalar(const char *seq, int cnt, uint32_t padw, const uint8_t *lut, uint32_t *nbits) {
    uint32_t word = padw, nb = 0;
    for (int i = 0; i < cnt; i++) {
        const uint8_t c = lut[(uint8_t)seq[i]];
        if (c == 4) { nb |= 1u << i; continue; }
        word = (word & ~(3u << (2 * i))) | ((uint32_t)c << (2 * i));
    }
    *nbits = nb;
    return word;
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define AF_HAVE_X86_PACK 1
// the same for 16 whole bases with SSSE3 + BMI2: codes by shifts, validity by looking the code's letter up
// again (pshufb) and comparing with the upper-cased input, 2-bit compaction with pext
__attribute__((target("ssse3,bmi2,sse4.1"))) static inline uint32_t pack16_x86(const char *seq, uint32_t padw, uint32_t *nbits) {
    const __m128i c = _mm_loadu_si128((const __m128i *)seq);
    const __m128i up = _mm_and_si128(c, _mm_set1_epi8((char)0xDF));
    const __m128i s1 = _mm_and_si128(_mm_srli_epi16(c, 1), _mm_set1_epi8(0x7F));
    const __m128i s2 = _mm_and_si128(_mm_srli_epi16(c, 2), _mm_set1_epi8(0x3F));
    const __m128i code = _mm_and_si128(_mm_xor_si128(s1, s2), _mm_set1_epi8(3));
    const __m128i letters = _mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m128i expect = _mm_shuffle_epi8(letters, code);
    const uint32_t valid = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(expect, up));      // bit i: base i is ACGT
    const uint64_t lo = (uint64_t)_mm_cvtsi128_si64(code), hi = (uint64_t)_mm_extract_epi64(code, 1);
    const uint32_t packed = (uint32_t)_pext_u64(lo, 0x0303030303030303ull) | ((uint32_t)_pext_u64(hi, 0x0303030303030303ull) << 16);
    const uint32_t vm = (uint32_t)_pdep_u32(valid, 0x55555555u) * 3u;                    // 2 mask bits per base
    *nbits = ~valid & 0xFFFFu;
    return (packed & vm) | (padw & ~vm);
}
static bool x86_pack_ok() {
    static const bool ok = __builtin_cpu_supports("ssse3") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("sse4.1");
    return ok;
}
#endif

// One mate's words of pairs [p0, p1) of a batch of n_pairs (mate 0 also writes the zero words that pad a
// pair to whole quads); r[0 .. p1 - p0) are the reads of those pairs.  p0 must be a multiple of 32 and
// p1 either a multiple of 32 or n_pairs: a range then owns whole tiles (the pairs past n_pairs in the last
// tile are written as pad), so ranges -- and the two mates, which touch disjoint 32-bit words -- can be
// packed by different threads at once.  N positions and the tail past a read keep the pad pattern.
void af_pack_range(const SeqRef *r, int m, int64_t p0, int64_t p1, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte,
                   void *packed_out, uint16_t *lens_out, PackSide &st) {
    af_layout_t lay;
    st.rc = af_layout(max_read_len, n_pairs, &lay);
    if (st.rc) return;
    const int W = lay.words_per_read, Q = lay.quads_per_pair;
    const uint8_t *lut = code_lut();
    uint32_t *out = (uint32_t *)packed_out;
    uint32_t padw = 0;  // 16 bases of the period-4 pad pattern
    for (int i = 0; i < 16; i++) padw |= (uint32_t)((pad_byte >> (2 * (i & 3))) & 3) << (2 * i);
#ifdef AF_HAVE_X86_PACK
    const bool fast = x86_pack_ok();
#endif
    const int64_t p_end = p1 == n_pairs ? lay.n_tiles * AF_TILE_PAIRS : p1;
    for (int64_t p = p0; p < p_end; p++) {
        const int64_t tile = p >> 5, lane = p & 31;
        uint32_t *pair_base = out + (tile * Q * 32 + lane) * 4;       // word wi of the pair at pair_base[(wi>>2)*128 + (wi&3)]
        if (m == 0) for (int wi = 2 * W; wi < 4 * Q; wi++) pair_base[(wi >> 2) * 128 + (wi & 3)] = 0u;
        int32_t len = 0;
        const char *seq = nullptr;
        if (p < n_pairs) {
            len = r[p - p0].len; seq = r[p - p0].p;
            if (len > max_read_len || len < 0) {
                char buf[160];
                snprintf(buf, sizeof(buf), "af_pack: read %lld/%d has %d bases, max_read_len is %d", (long long)p, m + 1, len, max_read_len);
                st.err = buf; st.rc = AF_ERR_ARG;
                return;
            }
            if (st.ulen == -1) st.ulen = len; else if (st.ulen != len) st.ulen = -2;
            if (lens_out) lens_out[2 * p + m] = (uint16_t)len;
        }
        uint32_t nm[AF_NMASK_WORDS] = {0};
        uint32_t anyn = 0;
        for (int t = 0; t < W; t++) {
            const int i0 = 16 * t, cnt = len - i0 < 16 ? (len - i0 > 0 ? len - i0 : 0) : 16;
            uint32_t word = padw, nb = 0;
#ifdef AF_HAVE_X86_PACK
            if (cnt == 16 && fast) word = pack16_x86(seq + i0, padw, &nb);
            else
#endif
            if (cnt) word = pack16_scalar(seq + i0, cnt, padw, lut, &nb);
            if (nb) { anyn = 1; nm[i0 >> 5] |= nb << (i0 & 31); }
            const int wi = m * W + t;
            pair_base[(wi >> 2) * 128 + (wi & 3)] = word;
        }
        if (anyn) {
            st.nids.push_back((uint32_t)(2 * p + m));
            st.nmask.insert(st.nmask.end(), nm, nm + AF_NMASK_WORDS);
        }
    }
}

void af_pack_side(const SeqRef *r, int m, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                  uint16_t *lens_out, PackSide &st) {
    af_pack_range(r, m, 0, n_pairs, n_pairs, max_read_len, pad_byte, packed_out, lens_out, st);
}

// merge the two sides' N lists (each sorted by read id) into the caller's arrays
int af_pack_finish(PackSide &a, PackSide &b, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                   int64_t *n_nreads_out, int32_t *uniform_len_out) {
    for (PackSide *s : {&a, &b}) if (s->rc) { af_set_error("%s", s->err.empty() ? "af_pack: bad layout" : s->err.c_str()); return s->rc; }
    const int64_t nn = (int64_t)(a.nids.size() + b.nids.size());
    if (nn > 0) {
        if (nn > ncap || !nread_ids_out || !nmask_out) { af_set_error("af_pack: N-read list capacity %lld exceeded", (long long)ncap); return AF_ERR_CAPACITY; }
        size_t ia = 0, ib = 0;
        int64_t k = 0;
        while (ia < a.nids.size() || ib < b.nids.size()) {
            const bool take_a = ib >= b.nids.size() || (ia < a.nids.size() && a.nids[ia] < b.nids[ib]);
            PackSide &s = take_a ? a : b;
            size_t &i = take_a ? ia : ib;
            nread_ids_out[k] = s.nids[i];
            memcpy(nmask_out + k * AF_NMASK_WORDS, s.nmask.data() + i * AF_NMASK_WORDS, AF_NMASK_WORDS * 4);
            i++; k++;
        }
    }
    if (n_nreads_out) *n_nreads_out = nn;
    if (uniform_len_out) {
        int32_t u = 0;
        if (a.ulen >= 0 && b.ulen >= 0 && a.ulen == b.ulen) u = a.ulen;
        *uniform_len_out = u > 0 ? u : 0;
    }
    return AF_OK;
}

// shared by af_pack_pairs and the FASTQ reader (af_fastq.cpp)
int af_pack_core(const SeqRef *r1, const SeqRef *r2, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte,
                 void *packed_out, uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                 int64_t *n_nreads_out, int32_t *uniform_len_out) {
    PackSide a, b;
    af_pack_side(r1, 0, n_pairs, max_read_len, pad_byte, packed_out, lens_out, a);
    af_pack_side(r2, 1, n_pairs, max_read_len, pad_byte, packed_out, lens_out, b);
    return af_pack_finish(a, b, nread_ids_out, nmask_out, ncap, n_nreads_out, uniform_len_out);
}

extern "C" int af_pack_pairs(const char *seq1, const int64_t *off1, const char *seq2, const int64_t *off2,
                             int64_t n_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                             uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                             int64_t *n_nreads_out, int32_t *uniform_len_out) {
    if (n_pairs < 0 || (n_pairs && (!seq1 || !seq2 || !off1 || !off2)) || !packed_out) { af_set_error("af_pack_pairs: null argument"); return AF_ERR_ARG; }
    std::vector<SeqRef> r1((size_t)n_pairs), r2((size_t)n_pairs);
    for (int64_t i = 0; i < n_pairs; i++) {
        r1[(size_t)i] = {seq1 + off1[i], (int32_t)(off1[i + 1] - off1[i])};
        r2[(size_t)i] = {seq2 + off2[i], (int32_t)(off2[i + 1] - off2[i])};
    }
    return af_pack_core(r1.data(), r2.data(), n_pairs, max_read_len, pad_byte, packed_out, lens_out, nread_ids_out,
                        nmask_out, ncap, n_nreads_out, uniform_len_out);
}

extern "C" int af_unpack_read(const void *packed, int32_t max_read_len, int64_t read_id, int32_t len, uint8_t *codes_out) {
    af_layout_t lay;
    int rc = af_layout(max_read_len, 0, &lay);
    if (rc) return rc;
    if (!packed || !codes_out || len < 0 || len > max_read_len) { af_set_error("af_unpack_read: bad argument"); return AF_ERR_ARG; }
    const uint32_t *in = (const uint32_t *)packed;
    int64_t p = read_id >> 1, tile = p >> 5, lane = p & 31;
    int m = (int)(read_id & 1), W = lay.words_per_read, Q = lay.quads_per_pair;
    for (int32_t i = 0; i < len; i++) {
        int wi = m * W + (i >> 4);
        uint32_t w = in[((tile * Q + (wi >> 2)) * 32 + lane) * 4 + (wi & 3)];
        codes_out[i] = (uint8_t)((w >> (2 * (i & 15))) & 3);
    }
    return AF_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic reads, host twin of the device generator (the reference's own
// recipe, utils/simulate_reads.py:20, shells out to wgsim, which is absent)
// ------------------------------------------------------------------------------------------
static int check_synth(const af_synth_t *s) {
    if (!s || s->ref_len < 4096 || s->anchor_len <= 0 || s->anchor_start < 0 ||
        s->anchor_start + s->anchor_len > s->ref_len || s->read_len <= 0 || s->read_len > AF_MAX_READ_LEN ||
        s->frag_mean < s->read_len || s->frag_sd < 0) {
        af_set_error("af_synth: bad generator description");
        return AF_ERR_ARG;
    }
    return AF_OK;
}

extern "C" int af_synth_anchor(const af_synth_t *s, char *ascii_out) {
    int rc = check_synth(s);
    if (rc) return rc;
    for (int32_t i = 0; i < s->anchor_len; i++) ascii_out[i] = "ACGT"[af_ref_base(s->seed, s->anchor_start + i)];
    return AF_OK;
}

extern "C" int af_synth_pairs_host(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, uint8_t *mate1, uint8_t *mate2) {
    int rc = check_synth(s);
    if (rc) return rc;
    const int L = s->read_len;
    for (int64_t p = 0; p < n_pairs; p++) {
        af_frag f = af_make_frag(*s, first_pair + p);
        for (int i = 0; i < L; i++) {
            mate1[p * L + i] = (uint8_t)af_read_base(*s, f, 0, i);
            mate2[p * L + i] = (uint8_t)af_read_base(*s, f, 1, i);
        }
    }
    return AF_OK;
}

// ------------------------------------------------------------------------------------------
// host twin of the seed-scan probe sequence: runs the SAME template the kernel runs
// (af_scan_read in af_common.h) on one pair's packed words.  For tests only -- it lets the CPU
// suite check the kernel's probe logic for every W without a GPU; no product path calls it.
// ------------------------------------------------------------------------------------------
template <int W, int KP>
static void scan_pair_host(const af_index *idx, const uint32_t *words, int nprobe, bool refine, int *f1, int *f2) {
    uint32_t w[2 * W];
    for (int i = 0; i < 2 * W; i++) w[i] = words[i];
    const uint32_t fm = idx->fmul;
    if (idx->bloom) {
        *f1 = af_scan_read<W, KP, 0, 2 * W, false, true>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
        *f2 = af_scan_read<W, KP, W, 2 * W, false, true>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
    } else if (refine) {
        *f1 = af_scan_read<W, KP, 0, 2 * W, true>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
        *f2 = af_scan_read<W, KP, W, 2 * W, true>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
    } else {
        *f1 = af_scan_read<W, KP, 0, 2 * W, false>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
        *f2 = af_scan_read<W, KP, W, 2 * W, false>(w, nprobe, idx->filter.data(), fm, idx->nb) != 0;
    }
}

#define AF_HOST_SCAN_CASE(WW) \
    case WW: if (idx->kp == 12) scan_pair_host<WW, 12>(idx, words, nprobe, refine, flag1, flag2); else scan_pair_host<WW, 13>(idx, words, nprobe, refine, flag1, flag2); return AF_OK;

extern "C" int af_debug_scan_pair(const af_index_t *idx, const uint32_t *words, int32_t words_per_read, int32_t read_len,
                                  int32_t with_neighbour_test, int32_t *flag1, int32_t *flag2) {
    const bool refine = with_neighbour_test != 0;
    if (!idx || !words || !flag1 || !flag2 || (idx->kp != 12 && idx->kp != 13) || idx->P.k != 19) { af_set_error("af_debug_scan_pair: bad argument"); return AF_ERR_ARG; }
    const int nprobe = af_nsamples(read_len, idx->kp);
    switch (words_per_read) {
        AF_HOST_SCAN_CASE(1) AF_HOST_SCAN_CASE(2) AF_HOST_SCAN_CASE(3) AF_HOST_SCAN_CASE(4) AF_HOST_SCAN_CASE(5)
        AF_HOST_SCAN_CASE(6) AF_HOST_SCAN_CASE(7) AF_HOST_SCAN_CASE(8) AF_HOST_SCAN_CASE(9) AF_HOST_SCAN_CASE(10)
        AF_HOST_SCAN_CASE(11) AF_HOST_SCAN_CASE(12) AF_HOST_SCAN_CASE(13) AF_HOST_SCAN_CASE(14) AF_HOST_SCAN_CASE(15)
        AF_HOST_SCAN_CASE(16) AF_HOST_SCAN_CASE(20) AF_HOST_SCAN_CASE(24) AF_HOST_SCAN_CASE(28) AF_HOST_SCAN_CASE(32)
    }
    af_set_error("af_debug_scan_pair: words_per_read %d", words_per_read);
    return AF_ERR_ARG;
}
