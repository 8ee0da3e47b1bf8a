// af_inflate.h -- raw DEFLATE (RFC 1951) decoder for the FASTQ.gz ingest, written for this path.
//
// zlib's inflate() runs at ~0.3 GB/s per core on FASTQ text, which made gzread the wall of the whole
// drop-in (round 1: 1.2 M pairs/s).  This decoder keeps a 64-bit bit buffer that is refilled once per
// length/distance pair, decodes through an 11-bit literal/length table (sub-tables for longer codes)
// whose entries carry base value, extra-bit count and code length in one word, emits two literals
// per table round trip when it can, and copies matches eight bytes at a time.  It is resumable at
// symbol boundaries, so one gzip member of any size can be decoded into a sequence of output
// segments: every segment is preceded in memory by the last 32 KB of the one before (the window).
//
// Never reads outside [in, in_end) and never writes outside [out_begin, out_hard_end); corrupt
// input ends in an error code, not in undefined behaviour (fuzzed under ASan, tools/fuzz_host.cpp).
#pragma once
#include <stdint.h>
#include <string.h>

namespace afz {

enum Status { OK_DONE = 0, NEED_OUTPUT = 1, ERR_DATA = -1, ERR_TRUNCATED = -2 };

static const int LL_BITS = 11, D_BITS = 8, PRE_BITS = 7;
// table entry: bits 0..4 code length to consume at this level; bit 5 LITERAL, bit 6 SPECIAL (end of
// block or invalid), bit 7 SUBTABLE; bits 8..12 extra bits (or sub-table index bits); bits 16..31 value
// A literal entry of the main literal/length table may carry TWO literals (second one in bits 24..31)
// when both codes fit the table index: bits 14..15 hold the literal count (1 or 2) and bits 0..4 the
// combined code length.  FASTQ text is mostly literals with 2-5 bit codes, so most lookups yield two bytes.
static const uint32_t E_LIT = 1u << 5, E_SPECIAL = 1u << 6, E_SUB = 1u << 7, E_EOB = 1u << 13;
static const int E_CNT_SHIFT = 14;

static inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void copy8(uint8_t *d, const uint8_t *s) { uint64_t v; memcpy(&v, s, 8); memcpy(d, &v, 8); }

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

struct Inflater {
    // input
    const uint8_t *in = nullptr, *in_end = nullptr;
    uint64_t bitbuf = 0;
    int bitcnt = 0;          // valid bits in bitbuf
    // block state
    enum { ST_HEADER, ST_STORED, ST_HUFF, ST_DONE } st = ST_HEADER;
    bool last_block = false;
    uint32_t stored_left = 0;
    uint32_t ll[(1 << LL_BITS) + 2048];   // main table + sub-tables (worst case well below this bound)
    uint32_t dt[(1 << D_BITS) + 1024];
    bool fixed_ready = false;

    void reset(const uint8_t *p, const uint8_t *e) {
        in = p; in_end = e; bitbuf = 0; bitcnt = 0; st = ST_HEADER; last_block = false; stored_left = 0; fixed_ready = false;
    }
    // bytes of input consumed so far, counting only whole bytes behind the bit cursor
    const uint8_t *byte_pos() const { return in - (bitcnt >> 3); }

    // ---- bit input ------------------------------------------------------------------------
    inline void refill_fast() {      // needs in + 8 <= in_end
        bitbuf |= load64(in) << bitcnt;
        in += (63 - bitcnt) >> 3;
        bitcnt |= 56;
    }
    inline void refill_safe() {
        while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; }
    }
    inline bool need(int n) {        // make n bits available (n <= 32); false when the input ends first
        if (bitcnt < n) { refill_safe(); if (bitcnt < n) return false; }
        return true;
    }
    inline uint32_t peek(int n) const { return (uint32_t)(bitbuf & ((1ull << n) - 1ull)); }
    inline void drop(int n) { bitbuf >>= n; bitcnt -= n; }

    // ---- canonical Huffman table construction -------------------------------------------------
    // lens[0..n): code lengths (0 = unused).  kind 0 = literal/length, 1 = distance, 2 = precode.
    // Returns false for an over-subscribed code, or an incomplete one (a single-code distance tree and
    // the all-zero distance tree excepted, as zlib accepts them).
    static uint32_t make_entry(int kind, int sym, int len) {
        if (kind == 2) return ((uint32_t)sym << 16) | (uint32_t)len;
        if (kind == 1) {
            if (sym >= 30) return E_SPECIAL | (uint32_t)len;                       // invalid distance symbol
            return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 8) | (uint32_t)len;
        }
        if (sym < 256) return ((uint32_t)sym << 16) | E_LIT | (1u << E_CNT_SHIFT) | (uint32_t)len;
        if (sym == 256) return E_SPECIAL | E_EOB | (uint32_t)len;
        if (sym >= 286) return E_SPECIAL | (uint32_t)len;                          // invalid length symbol
        return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 8) | (uint32_t)len;
    }
    static bool build(uint32_t *table, int table_cap, int main_bits, const uint8_t *lens, int n, int kind) {
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        int maxlen = 15;
        while (maxlen > 0 && count[maxlen] == 0) maxlen--;
        const int main_size = 1 << main_bits;
        if (maxlen == 0) {                       // no codes at all: every lookup is invalid
            for (int i = 0; i < main_size; i++) table[i] = E_SPECIAL | 1u;
            return kind == 1;                    // allowed for the distance tree of an all-literal block
        }
        // completeness
        int left = 1;
        for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
        if (left > 0 && !(kind == 1 && count[0] + 1 == n && maxlen == 1)) return false;
        uint16_t next_code[16];
        { int code = 0; count[0] = 0; for (int l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next_code[l] = (uint16_t)code; } }
        for (int i = 0; i < main_size; i++) table[i] = E_SPECIAL | 1u;            // invalid until filled (incomplete codes)
        // first pass: codes that fit the main table; remember, per main-table prefix, the longest long code
        uint8_t sub_bits[1 << LL_BITS];
        if (maxlen > main_bits) memset(sub_bits, 0, (size_t)main_size);
        uint16_t codes[320];
        for (int sym = 0; sym < n; sym++) {
            const int len = lens[sym];
            if (!len) continue;
            uint32_t c = next_code[len]++, r = 0;
            for (int b = 0; b < len; b++) r |= ((c >> b) & 1u) << (len - 1 - b);   // bit-reversed: codes are read LSB first
            codes[sym] = (uint16_t)r;
            if (len <= main_bits) {
                const uint32_t e = make_entry(kind, sym, len);
                for (uint32_t i = r; i < (uint32_t)main_size; i += 1u << len) table[i] = e;
            } else {
                const uint32_t pre = r & (uint32_t)(main_size - 1);
                if (len - main_bits > sub_bits[pre]) sub_bits[pre] = (uint8_t)(len - main_bits);
            }
        }
        if (maxlen <= main_bits) { if (kind == 0) pair_literals(table, main_bits); return true; }
        // second pass: allocate the sub-tables, then fill them
        int used = main_size;
        for (int pre = 0; pre < main_size; pre++) {
            if (!sub_bits[pre]) continue;
            const int sz = 1 << sub_bits[pre];
            if (used + sz > table_cap) return false;
            table[pre] = ((uint32_t)used << 16) | ((uint32_t)sub_bits[pre] << 8) | E_SUB | (uint32_t)main_bits;
            for (int i = 0; i < sz; i++) table[used + i] = E_SPECIAL | 1u;
            used += sz;
        }
        for (int sym = 0; sym < n; sym++) {
            const int len = lens[sym];
            if (len <= main_bits) continue;
            const uint32_t r = codes[sym], pre = r & (uint32_t)(main_size - 1);
            const uint32_t base = table[pre] >> 16, sb = (table[pre] >> 8) & 31u;
            const uint32_t e = make_entry(kind, sym, len - main_bits);
            for (uint32_t i = r >> main_bits; i < (1u << sb); i += 1u << (len - main_bits)) table[base + i] = e;
        }
        if (kind == 0) pair_literals(table, main_bits);
        return true;
    }
    // main-table entries whose index bits hold two complete literal codes get both literals
    static void pair_literals(uint32_t *table, int main_bits) {
        const uint32_t size = 1u << main_bits;
        uint32_t single[1 << LL_BITS];
        memcpy(single, table, size * 4);
        for (uint32_t i = 0; i < size; i++) {
            const uint32_t e1 = single[i];
            if (!(e1 & E_LIT)) continue;
            const uint32_t l1 = e1 & 31u;
            if (l1 >= (uint32_t)main_bits) continue;
            const uint32_t e2 = single[i >> l1];                 // the bits above the first code, zero-extended
            if (!(e2 & E_LIT)) continue;
            const uint32_t l2 = e2 & 31u;
            if (l1 + l2 > (uint32_t)main_bits) continue;          // second code not fully inside the index
            table[i] = (e1 & 0x00FF0000u) | ((e2 & 0x00FF0000u) << 8) | E_LIT | (2u << E_CNT_SHIFT) | (l1 + l2);
        }
    }

    bool build_fixed() {
        if (fixed_ready) return true;
        uint8_t l[288];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        if (!build(ll, (int)(sizeof(ll) / 4), LL_BITS, l, 288, 0)) return false;
        uint8_t d[32];
        for (int i = 0; i < 32; i++) d[i] = 5;
        if (!build(dt, (int)(sizeof(dt) / 4), D_BITS, d, 32, 1)) return false;
        fixed_ready = true;
        return true;
    }

    int read_dynamic_header() {
        if (!need(14)) return ERR_TRUNCATED;
        const int hlit = (int)peek(5) + 257; drop(5);
        const int hdist = (int)peek(5) + 1; drop(5);
        const int hclen = (int)peek(4) + 4; drop(4);
        if (hlit > 286 || hdist > 30) return ERR_DATA;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pl[19] = {0};
        for (int i = 0; i < hclen; i++) { if (!need(3)) return ERR_TRUNCATED; pl[order[i]] = (uint8_t)peek(3); drop(3); }
        uint32_t pre[1 << PRE_BITS];
        if (!build(pre, 1 << PRE_BITS, PRE_BITS, pl, 19, 2)) return ERR_DATA;
        uint8_t lens[286 + 30 + 140];
        int i = 0;
        const int total = hlit + hdist;
        while (i < total) {
            if (!need(PRE_BITS + 7)) { refill_safe(); if (bitcnt < 1) return ERR_TRUNCATED; }
            const uint32_t e = pre[peek(PRE_BITS)];
            if (e & E_SPECIAL) return ERR_DATA;
            const int clen = (int)(e & 31u), sym = (int)(e >> 16);
            if (clen > bitcnt) return ERR_TRUNCATED;
            drop(clen);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep, val = 0;
            if (sym == 16) {
                if (i == 0) return ERR_DATA;
                if (bitcnt < 2) return ERR_TRUNCATED;
                val = lens[i - 1]; rep = 3 + (int)peek(2); drop(2);
            } else if (sym == 17) {
                if (bitcnt < 3) return ERR_TRUNCATED;
                rep = 3 + (int)peek(3); drop(3);
            } else {
                if (bitcnt < 7) return ERR_TRUNCATED;
                rep = 11 + (int)peek(7); drop(7);
            }
            if (i + rep > total) return ERR_DATA;
            memset(lens + i, val, (size_t)rep);
            i += rep;
        }
        if (lens[256] == 0) return ERR_DATA;                                       // no end-of-block code
        fixed_ready = false;
        if (!build(ll, (int)(sizeof(ll) / 4), LL_BITS, lens, hlit, 0)) return ERR_DATA;
        if (!build(dt, (int)(sizeof(dt) / 4), D_BITS, lens + hlit, hdist, 1)) return ERR_DATA;
        return OK_DONE;
    }

    // Decode until the stream ends (OK_DONE), the output cursor reaches soft_end (NEED_OUTPUT; always at a
    // symbol boundary), or an error.  History: bytes [out_begin, *out) -- a back-reference may reach
    // down to out_begin, not below.  Writes stay below hard_end (soft_end <= hard_end - 258 lets every
    // symbol finish; for an exact-size one-shot decode pass soft_end = hard_end + 1: then running out of
    // room is ERR_DATA).
    int run(uint8_t *out_begin, uint8_t **outp, uint8_t *soft_end, uint8_t *hard_end) {
#if defined(__x86_64__) && defined(__GNUC__)
        static const bool bmi2 = __builtin_cpu_supports("bmi2");
        if (bmi2) return run_bmi2(out_begin, outp, soft_end, hard_end);
#endif
        return run_body(out_begin, outp, soft_end, hard_end);
    }
#if defined(__x86_64__) && defined(__GNUC__)
    // the same code compiled for BMI2 (shrx / bzhi: variable shifts and masks in one micro-op each)
    __attribute__((target("bmi2"))) int run_bmi2(uint8_t *out_begin, uint8_t **outp, uint8_t *soft_end, uint8_t *hard_end) {
        return run_body(out_begin, outp, soft_end, hard_end);
    }
#endif
    __attribute__((always_inline)) inline int run_body(uint8_t *out_begin, uint8_t **outp, uint8_t *soft_end, uint8_t *hard_end) {
        uint8_t *out = *outp;
        int rc = OK_DONE;
        for (;;) {
            if (st == ST_DONE) { rc = OK_DONE; break; }
            if (st == ST_HEADER) {
                if (!need(3)) { rc = ERR_TRUNCATED; break; }
                last_block = peek(1); drop(1);
                const uint32_t type = peek(2); drop(2);
                if (type == 0) {
                    drop(bitcnt & 7);                                              // to the byte boundary
                    if (!need(32)) { rc = ERR_TRUNCATED; break; }
                    const uint32_t len = peek(16); drop(16);
                    const uint32_t nlen = peek(16); drop(16);
                    if ((len ^ nlen) != 0xFFFFu) { rc = ERR_DATA; break; }
                    stored_left = len;
                    st = ST_STORED;
                } else if (type == 1) {
                    if (!build_fixed()) { rc = ERR_DATA; break; }
                    st = ST_HUFF;
                } else if (type == 2) {
                    rc = read_dynamic_header();
                    if (rc) break;
                    st = ST_HUFF;
                } else { rc = ERR_DATA; break; }
            }
            if (st == ST_STORED) {
                // whole bytes still sitting in the bit buffer go first
                while (stored_left && bitcnt >= 8) {
                    if (out >= hard_end) { rc = soft_end <= hard_end ? NEED_OUTPUT : ERR_DATA; goto done; }
                    *out++ = (uint8_t)peek(8); drop(8); stored_left--;
                }
                if (stored_left) {
                    size_t n = stored_left;
                    if ((size_t)(in_end - in) < n) n = (size_t)(in_end - in);
                    const size_t room = (size_t)(hard_end - out);
                    const bool short_out = room < n;
                    if (short_out) n = room;
                    memcpy(out, in, n);
                    out += n; in += n; stored_left -= (uint32_t)n;
                    if (stored_left) {
                        if (short_out) { rc = soft_end <= hard_end ? NEED_OUTPUT : ERR_DATA; goto done; }
                        rc = ERR_TRUNCATED; goto done;
                    }
                }
                st = last_block ? ST_DONE : ST_HEADER;
                if (out >= soft_end && st != ST_DONE) { rc = NEED_OUTPUT; break; }
                continue;
            }
            // ---- Huffman block -----------------------------------------------------------------
            {
                const uint32_t ll_mask = (1u << LL_BITS) - 1u, d_mask = (1u << D_BITS) - 1u;
                bool block_done = false;
                // fast loop: >= 32 input bytes and >= 320 output bytes of slack, no per-symbol bound checks.
                // The entry of the NEXT symbol is looked up before the bit buffer is refilled whenever
                // enough bits are left (>= 15: any code, or any pair of literals, is fully determined), so
                // the refill's load is off the table-lookup dependency chain, and the match copy overlaps both.
#define AFZ_FAST_OK() (in_end - in >= 32 && hard_end - out >= 320 && out < soft_end)
#define AFZ_EMIT_LIT(e) do { const uint16_t two_ = (uint16_t)((e) >> 16); memcpy(out, &two_, 2); out += ((e) >> E_CNT_SHIFT) & 3u; drop((e) & 31u); } while (0)
                if (AFZ_FAST_OK()) {
                    refill_fast();
                    uint32_t e = ll[bitbuf & ll_mask];
                    for (;;) {                             // e: entry at the cursor, not consumed yet; bitcnt >= 56
                        if (e & E_LIT) {                   // up to three literal entries (<= 33 bits, <= 6 bytes) per refill
                            AFZ_EMIT_LIT(e);
                            e = ll[bitbuf & ll_mask];
                            if (e & E_LIT) {
                                AFZ_EMIT_LIT(e);
                                e = ll[bitbuf & ll_mask];
                                if (e & E_LIT) {
                                    AFZ_EMIT_LIT(e);
                                    e = ll[bitbuf & ll_mask];       // >= 23 valid bits: e is exact
                                    refill_fast();
                                    if (AFZ_FAST_OK()) continue;
                                    break;
                                }
                            }
                            refill_fast();                 // e stays valid: a refill only adds bits above bitcnt
                        }
                        if (e & E_SUB) {
                            drop(e & 31u);
                            e = ll[(e >> 16) + (uint32_t)(bitbuf & ((1u << ((e >> 8) & 31u)) - 1u))];
                            if (e & E_LIT) {
                                drop(e & 31u); *out++ = (uint8_t)(e >> 16);
                                refill_fast();
                                e = ll[bitbuf & ll_mask];
                                if (AFZ_FAST_OK()) continue;
                                break;
                            }
                        }
                        if (e & E_SPECIAL) {
                            if (e & E_EOB) { drop(e & 31u); block_done = true; break; }
                            rc = ERR_DATA; goto done;
                        }
                        drop(e & 31u);
                        const uint32_t xb = (e >> 8) & 31u;
                        const uint32_t len = (e >> 16) + (uint32_t)(bitbuf & ((1u << xb) - 1u));
                        drop(xb);
                        uint32_t d = dt[bitbuf & d_mask];
                        if (d & E_SUB) {
                            drop(d & 31u);
                            d = dt[(d >> 16) + (uint32_t)(bitbuf & ((1u << ((d >> 8) & 31u)) - 1u))];
                        }
                        if (d & E_SPECIAL) { rc = ERR_DATA; goto done; }
                        drop(d & 31u);
                        const uint32_t dxb = (d >> 8) & 31u;
                        const uint32_t dist = (d >> 16) + (uint32_t)(bitbuf & ((1ull << dxb) - 1ull));
                        drop(dxb);
                        if (bitcnt >= 15) { e = ll[bitbuf & ll_mask]; refill_fast(); }
                        else { refill_fast(); e = ll[bitbuf & ll_mask]; }
                        if (dist > (size_t)(out - out_begin)) { rc = ERR_DATA; goto done; }
                        const uint8_t *src = out - dist;
                        uint8_t *dst = out;
                        out += len;
                        if (dist >= 8) {
                            copy8(dst, src); copy8(dst + 8, src + 8);
                            if (len > 16) { dst += 16; src += 16; do { copy8(dst, src); copy8(dst + 8, src + 8); dst += 16; src += 16; } while (dst < out); }
                        } else if (dist == 1) {
                            uint64_t v = 0x0101010101010101ull * src[0];
                            do { memcpy(dst, &v, 8); dst += 8; } while (dst < out);
                        } else {
                            do { *dst++ = *src++; } while (dst < out);
                        }
                        if (!AFZ_FAST_OK()) break;
                    }
                }
                // careful loop: every read and write checked
                while (!block_done) {
                    if (out >= soft_end) { rc = NEED_OUTPUT; goto done; }
                    if (in_end - in >= 32 && hard_end - out >= 320) break;      // back to the fast loop
                    refill_safe();
                    uint32_t e = ll[bitbuf & ll_mask];
                    int used = (int)(e & 31u);
                    if (e & E_SUB) {
                        if (bitcnt < used) { rc = ERR_TRUNCATED; goto done; }
                        drop(used);
                        e = ll[(e >> 16) + (uint32_t)(bitbuf & ((1u << ((e >> 8) & 31u)) - 1u))];
                        used = (int)(e & 31u);
                    }
                    if (bitcnt < used) { rc = ERR_TRUNCATED; goto done; }
                    if (e & E_SPECIAL) {
                        if (e & E_EOB) { drop(used); block_done = true; break; }
                        rc = ERR_DATA; goto done;
                    }
                    drop(used);
                    if (e & E_LIT) {
                        const uint32_t cnt = (e >> E_CNT_SHIFT) & 3u;
                        if ((size_t)(hard_end - out) < cnt) { rc = ERR_DATA; goto done; }
                        *out++ = (uint8_t)(e >> 16);
                        if (cnt == 2) *out++ = (uint8_t)(e >> 24);
                        continue;
                    }
                    const int xb = (int)((e >> 8) & 31u);
                    if (bitcnt < xb) { rc = ERR_TRUNCATED; goto done; }
                    const uint32_t len = (e >> 16) + (uint32_t)(bitbuf & ((1u << xb) - 1u));
                    drop(xb);
                    refill_safe();
                    uint32_t d = dt[bitbuf & d_mask];
                    used = (int)(d & 31u);
                    if (d & E_SUB) {
                        if (bitcnt < used) { rc = ERR_TRUNCATED; goto done; }
                        drop(used);
                        d = dt[(d >> 16) + (uint32_t)(bitbuf & ((1u << ((d >> 8) & 31u)) - 1u))];
                        used = (int)(d & 31u);
                    }
                    if (bitcnt < used) { rc = ERR_TRUNCATED; goto done; }
                    if (d & E_SPECIAL) { rc = ERR_DATA; goto done; }
                    drop(used);
                    const int dxb = (int)((d >> 8) & 31u);
                    if (bitcnt < dxb) { rc = ERR_TRUNCATED; goto done; }
                    const uint32_t dist = (d >> 16) + (uint32_t)(bitbuf & ((1ull << dxb) - 1ull));
                    drop(dxb);
                    if (dist > (size_t)(out - out_begin) || len > (size_t)(hard_end - out)) { rc = ERR_DATA; goto done; }
                    const uint8_t *src = out - dist;
                    for (uint32_t i = 0; i < len; i++) out[i] = src[i];
                    out += len;
                }
                if (block_done) {
                    st = last_block ? ST_DONE : ST_HEADER;
                    if (out >= soft_end && st != ST_DONE) { rc = NEED_OUTPUT; break; }
                }
            }
        }
    done:
        *outp = out;
        return rc;
    }
};

// ---- gzip member framing (RFC 1952) ---------------------------------------------------------
// Parses the member header at p.  On success *hdr_len is the header's size and, for a BGZF block
// (extra subfield 'B','C'), *bgzf_block_size its total size (else 0).  Returns false when p does not
// start a well-formed gzip member inside [p, end).
static inline bool gzip_header(const uint8_t *p, const uint8_t *end, size_t *hdr_len, uint32_t *bgzf_block_size) {
    *bgzf_block_size = 0;
    if (end - p < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return false;
    const uint8_t flg = p[3];
    if (flg & 0xE0) return false;
    const uint8_t *q = p + 10;
    if (flg & 4) {                                   // FEXTRA
        if (end - q < 2) return false;
        const size_t xlen = (size_t)q[0] | ((size_t)q[1] << 8);
        q += 2;
        if ((size_t)(end - q) < xlen) return false;
        const uint8_t *x = q, *xe = q + xlen;
        while (xe - x >= 4) {
            const size_t sl = (size_t)x[2] | ((size_t)x[3] << 8);
            if ((size_t)(xe - x - 4) < sl) break;
            if (x[0] == 'B' && x[1] == 'C' && sl == 2) *bgzf_block_size = ((uint32_t)x[4] | ((uint32_t)x[5] << 8)) + 1u;
            x += 4 + sl;
        }
        q += xlen;
    }
    if (flg & 8) { while (q < end && *q) q++; if (q >= end) return false; q++; }     // FNAME
    if (flg & 16) { while (q < end && *q) q++; if (q >= end) return false; q++; }    // FCOMMENT
    if (flg & 2) { if (end - q < 2) return false; q += 2; }                          // FHCRC
    *hdr_len = (size_t)(q - p);
    return true;
}

}  // namespace afz
