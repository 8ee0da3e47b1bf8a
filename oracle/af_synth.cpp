// af_synth.cpp -- the synthetic read-pair generator for the CPU arm (bench.py --impl reference, cpu_baseline).
//
// TEST / MEASUREMENT INFRASTRUCTURE, not product code.  It exists so that the reference arm of bench.py
// loads no product library: the generator is a pure function of (seed, pair index) defined inline in
// anchored_fusion_b200/csrc/af_common.h (af_make_frag / af_read_base / af_ref_base), shared with the
// device generator k_synth_pairs, so both arms see the same pairs.  Stands in for the reference's
// utils/simulate_reads.py:20 (wgsim -d 200 -1 101 -2 101; wgsim is absent and the configs ask for 2x150).
#include <cstdarg>
#include <cstdint>
#include "../anchored_fusion_b200/csrc/af_common.h"

void af_set_error(const char *, ...) {}

extern "C" int afo_synth_anchor(const af_synth_t *s, char *ascii_out) {
    if (!s || !ascii_out || s->anchor_len <= 0) return -1;
    for (int32_t i = 0; i < s->anchor_len; i++) ascii_out[i] = "ACGT"[af_ref_base(s->seed, s->anchor_start + i)];
    return 0;
}

// codes 0..4, reads interleaved: row 2p = mate 1 of pair first_pair+p, row 2p+1 = mate 2; `stride` bytes per row
extern "C" int afo_synth_reads(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, uint8_t *reads, int64_t stride,
                               int threads) {
    if (!s || !reads || n_pairs < 0 || stride < s->read_len) return -1;
    const int L = s->read_len;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t p = 0; p < n_pairs; p++) {
        af_frag f = af_make_frag(*s, first_pair + p);
        uint8_t *r1 = reads + (2 * p) * stride, *r2 = r1 + stride;
        for (int i = 0; i < L; i++) {
            r1[i] = (uint8_t)af_read_base(*s, f, 0, i);
            r2[i] = (uint8_t)af_read_base(*s, f, 1, i);
        }
    }
    return 0;
}

// ---- the same pairs as FASTQ files (measurement input for the ingest path) -------------------------------
// Illumina-style names (both mates share the name up to the blank), binned qualities the way current
// instruments write them (mostly 'F', ~8 % of the bases in one of three lower bins, in short runs).
// format 0: plain text, 1: one gzip member (zlib), 2: BGZF (bgzip: independent <= 64 KB blocks).
#include <zlib.h>
#include <algorithm>
#include <cstring>

#include <cstdio>
#include <string>
#include <vector>

static void fastq_text(const af_synth_t *s, int64_t first_pair, int64_t n, int mate, std::string &out) {
    const int L = s->read_len;
    out.clear();
    out.reserve((size_t)n * (size_t)(2 * L + 64));
    char name[96];
    for (int64_t p = 0; p < n; p++) {
        const int64_t g = first_pair + p;
        af_frag f = af_make_frag(*s, g);
        const int nl = snprintf(name, sizeof(name), "@A00123:45:HXXXXXX:%d:%d:%d:%d %d:N:0:ACGTACGT\n", 1 + (int)((g >> 24) & 3),
                                1101 + (int)((g >> 17) % 600), (int)((g * 7) % 32768), (int)(g % 100000), mate + 1);
        out.append(name, (size_t)nl);
        for (int i = 0; i < L; i++) out += "ACGTN"[af_read_base(*s, f, mate, i)];
        out += "\n+\n";
        uint32_t h = af_mix32((uint32_t)g * 2654435761u + (uint32_t)mate);
        char q = 'F';
        for (int i = 0; i < L; i++) {
            if ((i & 7) == 0) h = af_mix32(h + (uint32_t)i);
            const uint32_t r = (h >> ((i & 7) * 4)) & 15u;
            if (r == 0) q = ":,#F"[(h >> 28) & 3];       // a new run starts: ~1 base in 16
            else if (r < 6) q = 'F';
            out += q;
        }
        out += '\n';
    }
}

static bool write_bgzf(FILE *fh, const std::string &t, int level) {
    std::vector<unsigned char> comp(70000);
    for (size_t i = 0; i < t.size(); i += 0xFF00) {
        const size_t n = std::min<size_t>(0xFF00, t.size() - i);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = (Bytef *)t.data() + i; zs.avail_in = (uInt)n;
        zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
        deflate(&zs, Z_FINISH);
        const unsigned clen = (unsigned)zs.total_out, bsize = clen + 25;
        deflateEnd(&zs);
        const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 255), (unsigned char)(bsize >> 8)};
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)t.data() + i, (uInt)n), isz = (uint32_t)n;
        unsigned char tail[8];
        for (int k = 0; k < 4; k++) { tail[k] = (unsigned char)(crc >> (8 * k)); tail[4 + k] = (unsigned char)(isz >> (8 * k)); }
        if (fwrite(hdr, 1, 18, fh) != 18 || fwrite(comp.data(), 1, clen, fh) != clen || fwrite(tail, 1, 8, fh) != 8) return false;
    }
    static const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    return fwrite(eof, 1, 28, fh) == 28;
}

static int write_one(const af_synth_t *s, int64_t first_pair, int64_t n, int mate, const char *path, int format, int level) {
    std::string t;
    fastq_text(s, first_pair, n, mate, t);
    if (format == 1) {
        char mode[8];
        snprintf(mode, sizeof(mode), "wb%d", level);
        gzFile g = gzopen(path, mode);
        if (!g) return -1;
        size_t off = 0;
        while (off < t.size()) { const int w = gzwrite(g, t.data() + off, (unsigned)std::min<size_t>(t.size() - off, 1u << 30)); if (w <= 0) { gzclose(g); return -1; } off += (size_t)w; }
        return gzclose(g) == Z_OK ? 0 : -1;
    }
    FILE *fh = fopen(path, "wb");
    if (!fh) return -1;
    bool ok = format == 2 ? write_bgzf(fh, t, level) : fwrite(t.data(), 1, t.size(), fh) == t.size();
    return (fclose(fh) == 0 && ok) ? 0 : -1;
}

// One large file written by many threads: the pair range is cut into chunks, every chunk's text is generated and
// compressed on its own (BGZF blocks are independent anyway; for a single gzip member every chunk is a raw deflate
// stream closed with a sync flush -- an empty stored block on a byte boundary -- and the last one with Z_FINISH, which
// concatenate into ONE valid deflate stream, the way pigz writes them), then the pieces are written in order.
static int write_one_chunked(const af_synth_t *s, int64_t first_pair, int64_t n, int mate, const char *path, int format, int level, int threads) {
    const int64_t CH = 1 << 16;
    const int64_t n_chunks = std::max<int64_t>(1, (n + CH - 1) / CH);
    std::vector<std::string> pieces((size_t)n_chunks);
    std::vector<uint32_t> crcs((size_t)n_chunks, 0);
    std::vector<size_t> lens((size_t)n_chunks, 0);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1) reduction(+ : bad)
    for (int64_t c = 0; c < n_chunks; c++) {
        std::string t;
        const int64_t p0 = c * CH, cnt = std::min<int64_t>(CH, n - p0);
        fastq_text(s, first_pair + p0, cnt > 0 ? cnt : 0, mate, t);
        std::string &out = pieces[(size_t)c];
        if (format == 0) { out.swap(t); continue; }
        if (format == 2) {                                   // BGZF blocks of this chunk (no EOF block)
            std::vector<unsigned char> comp(70000);
            for (size_t i = 0; i < t.size(); i += 0xFF00) {
                const size_t m = std::min<size_t>(0xFF00, t.size() - i);
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad++; break; }
                zs.next_in = (Bytef *)t.data() + i; zs.avail_in = (uInt)m;
                zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
                deflate(&zs, Z_FINISH);
                const unsigned clen = (unsigned)zs.total_out, bsize = clen + 25;
                deflateEnd(&zs);
                const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 255), (unsigned char)(bsize >> 8)};
                const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)t.data() + i, (uInt)m), isz = (uint32_t)m;
                unsigned char tail[8];
                for (int k = 0; k < 4; k++) { tail[k] = (unsigned char)(crc >> (8 * k)); tail[4 + k] = (unsigned char)(isz >> (8 * k)); }
                out.append((const char *)hdr, 18); out.append((const char *)comp.data(), clen); out.append((const char *)tail, 8);
            }
            continue;
        }
        // format 1: a piece of one raw deflate stream
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad++; continue; }
        out.resize(deflateBound(&zs, (uLong)t.size()) + 64);
        zs.next_in = (Bytef *)t.data(); zs.avail_in = (uInt)t.size();
        zs.next_out = (Bytef *)&out[0]; zs.avail_out = (uInt)out.size();
        const int rc = deflate(&zs, c + 1 == n_chunks ? Z_FINISH : Z_SYNC_FLUSH);
        if ((c + 1 == n_chunks && rc != Z_STREAM_END) || (c + 1 != n_chunks && (rc != Z_OK || zs.avail_in))) bad++;
        out.resize(zs.total_out);
        deflateEnd(&zs);
        crcs[(size_t)c] = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)t.data(), (uInt)t.size());
        lens[(size_t)c] = t.size();
    }
    if (bad) return -1;
    FILE *fh = fopen(path, "wb");
    if (!fh) return -1;
    bool ok = true;
    if (format == 1) {
        const unsigned char hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
        ok = fwrite(hdr, 1, 10, fh) == 10;
    }
    uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
    uint64_t total = 0;
    for (int64_t c = 0; c < n_chunks && ok; c++) {
        ok = fwrite(pieces[(size_t)c].data(), 1, pieces[(size_t)c].size(), fh) == pieces[(size_t)c].size();
        if (format == 1) { crc = (uint32_t)crc32_combine(crc, crcs[(size_t)c], (z_off_t)lens[(size_t)c]); total += lens[(size_t)c]; }
    }
    if (ok && format == 1) {
        unsigned char tail[8];
        for (int k = 0; k < 4; k++) { tail[k] = (unsigned char)(crc >> (8 * k)); tail[4 + k] = (unsigned char)((uint32_t)total >> (8 * k)); }
        ok = fwrite(tail, 1, 8, fh) == 8;
    }
    if (ok && format == 2) {
        static const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        ok = fwrite(eof, 1, 28, fh) == 28;
    }
    return (fclose(fh) == 0 && ok) ? 0 : -1;
}

// n_files file pairs, file i holding pairs [first_pair + i * pairs_per_file, ... + pairs_per_file); paths1[i] / paths2[i]
extern "C" int afo_synth_fastq(const af_synth_t *s, int64_t first_pair, int64_t pairs_per_file, int32_t n_files,
                               const char *const *paths1, const char *const *paths2, int32_t format, int32_t level, int threads) {
    if (!s || !paths1 || !paths2 || n_files <= 0 || pairs_per_file < 0) return -1;
    int bad = 0;
    if (2 * n_files < threads && pairs_per_file > (1 << 16)) {          // few large files: the threads share each file
        for (int k = 0; k < 2 * n_files; k++) {
            const int i = k >> 1, mate = k & 1;
            if (write_one_chunked(s, first_pair + (int64_t)i * pairs_per_file, pairs_per_file, mate, mate ? paths2[i] : paths1[i], format, level, threads)) bad++;
        }
        return bad ? -1 : 0;
    }
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1) reduction(+ : bad)
    for (int k = 0; k < 2 * n_files; k++) {
        const int i = k >> 1, mate = k & 1;
        if (write_one(s, first_pair + (int64_t)i * pairs_per_file, pairs_per_file, mate, mate ? paths2[i] : paths1[i], format, level)) bad++;
    }
    return bad ? -1 : 0;
}
