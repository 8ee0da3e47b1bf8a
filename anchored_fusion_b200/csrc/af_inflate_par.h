// af_inflate_par.h -- decoding ONE deflate stream on several threads (single-member .fastq.gz, the common case).
//
// A deflate stream is serial twice over: a block can start at any bit, and every back-reference may reach 32 KB
// into what came before.  The two-pass scheme of pugz / rapidgzip (Kerbiriou & Chikhi 2019; Knespel & Brunst 2023)
// removes both for text:
//   1. the compressed bytes are cut into chunks; a worker FINDS a block start after its chunk's first byte by trying
//      bit positions until a dynamic-Huffman header parses, its code is complete, the block decodes to plausible
//      FASTQ text and the block after it parses and decodes as well;
//   2. from there it decodes SYMBOLICALLY into 16-bit cells: a byte it knows, or a marker "byte i of the 32 KB window
//      in front of my chunk" -- references into the unknown window just copy markers along;
//   3. chunks are then closed in order: a chunk must start exactly where the one before it stopped (else it is decoded
//      again from the right bit), and once the window in front of it is known its markers are replaced -- slice by
//      slice, in parallel -- by the window's bytes.
// The member's CRC-32 is checked over the resolved text as usual, so a wrong guess anywhere cannot go unnoticed.
// Never reads outside [base, end); corrupt input ends in an error code (fuzzed with the reader, tools/fuzz_host.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <stdlib.h>

#include <vector>

#include "af_inflate.h"

namespace afz {

static const uint16_t SYM_MARK = 0x8000u;     // cell = SYM_MARK | index into the 32 KB window before the chunk
static const size_t WINDOW = 32768;

// growable array of cells without value-initialisation (a std::vector would zero-fill every growth: at 2 bytes per
// decoded byte that was a third of the decode time)
struct CellBuf {
    uint16_t *p = nullptr;
    size_t n = 0, cap = 0;
    CellBuf() {}
    CellBuf(const CellBuf &) = delete;
    CellBuf &operator=(const CellBuf &) = delete;
    CellBuf(CellBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    CellBuf &operator=(CellBuf &&o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
    ~CellBuf() { free(p); }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        uint16_t *q = (uint16_t *)realloc(p, want * sizeof(uint16_t));
        if (!q) return false;
        p = q; cap = want;
        return true;
    }
    size_t size() const { return n; }
    const uint16_t *data() const { return p; }
    void clear() { n = 0; }
};

struct SymResult {
    CellBuf sym;                // decoded cells
    uint64_t start_bit = 0;     // bit position (from `base`) of the block the decode started at
    uint64_t end_bit = 0;       // bit position where it stopped: a block start, or the end of the final block
    bool stream_end = false;    // stopped because the final block ended
    bool uses_window = false;   // at least one marker was written
    int status = OK_DONE;       // OK_DONE, ERR_DATA, ERR_TRUNCATED
};

struct SymDecoder {
    Inflater inf;
    const uint8_t *base = nullptr;

    uint64_t bitpos() const { return (uint64_t)(inf.in - base) * 8u - (uint64_t)inf.bitcnt; }

    void seek(const uint8_t *b, const uint8_t *end, uint64_t bit) {
        base = b;
        inf.reset(b + (bit >> 3), end);
        inf.refill_safe();
        inf.drop((int)(bit & 7));       // (a position past the end leaves bitcnt < 0 .. 7: need() fails later)
    }

    // Decodes whole blocks from the current position (which must be a block start) until a block boundary at or
    // after stop_bit, the end of the stream, max_out cells, or an error.  text_only: a literal outside FASTQ's
    // alphabet (LF, CR, TAB, 32..126) is an error -- used while looking for a block start.
    // have_window = false: the decode starts at the very beginning of the stream, a reference before it is an error.
    int run(CellBuf &out, uint64_t stop_bit, size_t max_out, bool text_only, bool have_window, bool *stream_end,
            bool *uses_window, int max_blocks = 1 << 30) {
        *stream_end = false;
        const uint32_t ll_mask = (1u << LL_BITS) - 1u, d_mask = (1u << D_BITS) - 1u;
        size_t n = out.n;
        int blocks = 0;
        for (;;) {
            // ---- block header ----
            if (bitpos() >= stop_bit || blocks >= max_blocks) { out.n = n; return OK_DONE; }
            if (!inf.need(3)) return ERR_TRUNCATED;
            const bool last = inf.peek(1); inf.drop(1);
            const uint32_t type = inf.peek(2); inf.drop(2);
            if (type == 0) {
                inf.drop(inf.bitcnt & 7);
                if (!inf.need(32)) return ERR_TRUNCATED;
                const uint32_t len = inf.peek(16); inf.drop(16);
                const uint32_t nlen = inf.peek(16); inf.drop(16);
                if ((len ^ nlen) != 0xFFFFu) return ERR_DATA;
                if (n + len > max_out) return ERR_DATA;
                if (out.cap < n + len && !out.reserve(n + len + (n >> 1) + 4096)) return ERR_DATA;
                for (uint32_t i = 0; i < len; i++) {
                    if (!inf.need(8)) return ERR_TRUNCATED;
                    const uint32_t c = inf.peek(8); inf.drop(8);
                    if (text_only && !(c == 10 || c == 13 || c == 9 || (c >= 32 && c < 127))) return ERR_DATA;
                    out.p[n++] = (uint16_t)c;
                }
            } else {
                if (type == 1) { if (!inf.build_fixed()) return ERR_DATA; }
                else if (type == 2) { const int rc = inf.read_dynamic_header(); if (rc) return rc; }
                else return ERR_DATA;
                if (out.cap < n + 1024 && !out.reserve(n + (n >> 1) + 65536)) return ERR_DATA;
                uint16_t *o = out.p;
                size_t room = out.cap - 600;       // n <= room: a refill's worth of symbols fits (3 x 2 literals, or 258)
                // the bit cursor lives in locals inside the symbol loop (the stores to the cells would otherwise force it
                // through memory on every step)
                uint64_t bb = inf.bitbuf;
                int bc = inf.bitcnt;
                const uint8_t *ip = inf.in, *const ie = inf.in_end;
#define SYM_RET(x) do { inf.bitbuf = bb; inf.bitcnt = bc; inf.in = ip; return (x); } while (0)
#define SYM_DROP(k) do { bb >>= (k); bc -= (int)(k); } while (0)
                for (;;) {
                    if (ie - ip >= 8) { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; }
                    else { while (bc <= 56 && ip < ie) { bb |= (uint64_t)*ip++ << bc; bc += 8; } }
                    if (n > room) { if (!out.reserve(n + (n >> 1) + 65536)) SYM_RET(ERR_DATA); o = out.p; room = out.cap - 600; }
                    uint32_t e = inf.ll[bb & ll_mask];
                    // up to three literal entries per refill (<= 33 of the >= 56 bits a fast refill leaves)
                    if ((e & E_LIT) && bc >= 48 && !text_only) {
                        o[n] = (uint16_t)((e >> 16) & 255u); o[n + 1] = (uint16_t)(e >> 24); n += (e >> E_CNT_SHIFT) & 3u; SYM_DROP((int)(e & 31u));
                        e = inf.ll[bb & ll_mask];
                        if (e & E_LIT) {
                            o[n] = (uint16_t)((e >> 16) & 255u); o[n + 1] = (uint16_t)(e >> 24); n += (e >> E_CNT_SHIFT) & 3u; SYM_DROP((int)(e & 31u));
                            e = inf.ll[bb & ll_mask];
                            if (e & E_LIT) {
                                o[n] = (uint16_t)((e >> 16) & 255u); o[n + 1] = (uint16_t)(e >> 24); n += (e >> E_CNT_SHIFT) & 3u; SYM_DROP((int)(e & 31u));
                                if (n > max_out) SYM_RET(ERR_DATA);
                                continue;
                            }
                        }
                        if (bc < 48) continue;                    // not enough bits left for a whole match: refill first
                    }
                    int used = (int)(e & 31u);
                    if (e & E_SUB) {
                        if (bc < used) SYM_RET(ERR_TRUNCATED);
                        SYM_DROP(used);
                        e = inf.ll[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 31u)) - 1u))];
                        used = (int)(e & 31u);
                    }
                    if (bc < used) SYM_RET(ERR_TRUNCATED);
                    if (e & E_SPECIAL) {
                        if (e & E_EOB) { SYM_DROP(used); break; }
                        SYM_RET(ERR_DATA);
                    }
                    SYM_DROP(used);
                    if (e & E_LIT) {
                        const uint32_t cnt = (e >> E_CNT_SHIFT) & 3u, c0 = (e >> 16) & 255u, c1 = e >> 24;
                        if (text_only && (!(c0 == 10 || c0 == 13 || c0 == 9 || (c0 >= 32 && c0 < 127)) ||
                                          (cnt == 2 && !(c1 == 10 || c1 == 13 || c1 == 9 || (c1 >= 32 && c1 < 127))))) SYM_RET(ERR_DATA);
                        o[n++] = (uint16_t)c0;
                        if (cnt == 2) o[n++] = (uint16_t)c1;
                        if (n > max_out) SYM_RET(ERR_DATA);
                        continue;
                    }
                    const int xb = (int)((e >> 8) & 31u);
                    if (bc < xb) SYM_RET(ERR_TRUNCATED);
                    const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << xb) - 1u));
                    SYM_DROP(xb);
                    uint32_t d = inf.dt[bb & d_mask];
                    used = (int)(d & 31u);
                    if (d & E_SUB) {
                        if (bc < used) SYM_RET(ERR_TRUNCATED);
                        SYM_DROP(used);
                        d = inf.dt[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 31u)) - 1u))];
                        used = (int)(d & 31u);
                    }
                    if (bc < used) SYM_RET(ERR_TRUNCATED);
                    if (d & E_SPECIAL) SYM_RET(ERR_DATA);
                    SYM_DROP(used);
                    const int dxb = (int)((d >> 8) & 31u);
                    if (bc < dxb) SYM_RET(ERR_TRUNCATED);
                    const uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1ull << dxb) - 1ull));
                    SYM_DROP(dxb);
                    if (n + len > max_out) SYM_RET(ERR_DATA);
                    uint16_t *dst = o + n;
                    if (dist <= n) {                                    // inside what this decode has produced
                        const uint16_t *src = dst - dist;
                        uint16_t *const de = dst + len;
                        if (dist >= 8) {                                // 8 cells per step; the slack behind n takes the overshoot
                            do { memcpy(dst, src, 16); dst += 8; src += 8; } while (dst < de);
                        } else if (dist == 1) {                         // a run (binned qualities, poly-A)
                            const uint16_t v = src[0];
                            do { dst[0] = dst[1] = dst[2] = dst[3] = dst[4] = dst[5] = dst[6] = dst[7] = v; dst += 8; } while (dst < de);
                        } else {
                            do { *dst++ = *src++; } while (dst < de);
                        }
                    } else {
                        if (!have_window || dist > WINDOW) SYM_RET(ERR_DATA);
                        *uses_window = true;
                        for (uint32_t i = 0; i < len; i++) {
                            const size_t at = n + i;                    // position of the cell being written
                            dst[i] = at >= dist ? o[at - dist] : (uint16_t)(SYM_MARK | (uint16_t)(WINDOW - (dist - at)));
                        }
                    }
                    n += len;
                }
                inf.bitbuf = bb; inf.bitcnt = bc; inf.in = ip;
#undef SYM_RET
#undef SYM_DROP
            }
            blocks++;
            out.n = n;
            if (last) { *stream_end = true; return OK_DONE; }
        }
    }
};

// First block start at a bit position >= from_bit (and < limit_bit) that survives the tests above, or ~0ull.
// Only dynamic-Huffman, non-final blocks are looked for: that is what a compressor emits in the middle of a text.
static inline uint64_t find_block_start(const uint8_t *base, const uint8_t *end, uint64_t from_bit, uint64_t limit_bit) {
    SymDecoder dec;
    CellBuf scratch;
    const uint64_t last_bit = (uint64_t)(end - base) * 8u;
    if (limit_bit > last_bit) limit_bit = last_bit;
    for (uint64_t b = from_bit; b + 64 < limit_bit; b++) {
        // cheap tests on the raw bits first: BFINAL = 0, BTYPE = 10, HLIT <= 29, HDIST <= 29
        const uint8_t *p = base + (b >> 3);
        if (end - p < 8) break;
        const uint64_t v = load64(p) >> (b & 7);
        if ((v & 7u) != 4u) continue;
        if (((v >> 3) & 31u) > 29u || ((v >> 8) & 31u) > 29u) continue;
        dec.seek(base, end, b);
        scratch.clear();
        bool se = false, uw = false;
        // two consecutive blocks must parse and decode to text (the second may be final)
        int rc = dec.run(scratch, ~0ull, (size_t)8 << 20, true, true, &se, &uw, 1);
        if (rc != OK_DONE || scratch.size() < 256) continue;
        if (!se) {
            const size_t n1 = scratch.size();
            rc = dec.run(scratch, ~0ull, (size_t)16 << 20, true, true, &se, &uw, 1);
            if (rc != OK_DONE || scratch.size() < n1 + 256) continue;
        }
        return b;
    }
    return ~0ull;
}

// cells -> bytes; window = the WINDOW bytes in front of the chunk the cells belong to
static inline void resolve_cells(const uint16_t *sym, size_t n, const uint8_t *window, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        const uint16_t c = sym[i];
        out[i] = (c & SYM_MARK) ? window[c & (SYM_MARK - 1)] : (uint8_t)c;
    }
}

}  // namespace afz

namespace afz {

// One chunk of a deflate stream, decoded symbolically.  start_bit = ~0: the chunk's first block is searched from
// nominal_bit on; else the decode starts at start_bit (known to be a block start).  Stops at the first block boundary
// at or after stop_bit (or at the end of the stream).
static inline void decode_chunk(const uint8_t *base, const uint8_t *end, uint64_t start_bit, uint64_t nominal_bit, uint64_t stop_bit,
                                bool have_window, SymResult &r) {
    r.sym.clear(); r.start_bit = r.end_bit = 0; r.stream_end = r.uses_window = false; r.status = OK_DONE;
    if (start_bit == ~0ull) {
        start_bit = find_block_start(base, end, nominal_bit, stop_bit);
        if (start_bit == ~0ull) { r.status = ERR_DATA; return; }       // nothing found: the chunk before runs on through here
    }
    r.start_bit = start_bit;
    SymDecoder dec;
    dec.seek(base, end, start_bit);
    const uint64_t last_bit = (uint64_t)(end - base) * 8u, upto = stop_bit < last_bit ? stop_bit : last_bit;
    r.sym.reserve((size_t)((upto > start_bit ? upto - start_bit : 0) / 8u) * 5u + 65536u);
    // (a chunk of FASTQ text expands 4-5x; the caller sends members that compress far better than that down the serial path,
    // so the bound only stops a damaged or hostile stream from asking for tens of gigabytes)
    const size_t max_cells = (size_t)((upto > start_bit ? upto - start_bit : 0) / 8u) * 64u + ((size_t)64 << 20);
    r.status = dec.run(r.sym, stop_bit, max_cells, false, have_window, &r.stream_end, &r.uses_window);
    r.end_bit = dec.bitpos();
}

}  // namespace afz
