// af_genome_host.cpp -- host side of the genome pass (af_genome.cu): FASTA (plain / gzip) or in-memory contigs ->
// ONE 2-bit packed sequence [256 N] contig [256 N] contig ... [256 N] plus a 1-bit-per-base N map.  No CUDA here, so
// the parser of foreign files runs under AddressSanitizer / UBSan in tools/fuzz_host.cpp.
#include <zlib.h>

#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "af_common.h"
#include "af_crc32.h"

// an N is stored as a position-dependent pseudo-random base (and a set bit in the N bitmap), so the long N runs
// of an assembly look like random sequence to the filter instead of 150 M copies of one 12-mer
static inline uint32_t n_word(int64_t word_index) { return af_mix32((uint32_t)word_index * 0x9E3779B1u + 0x5BD1E995u); }

struct GenomeBuilder {
    std::vector<uint32_t> pk, nm;
    int64_t n = 0;
    uint8_t lut[256];
    GenomeBuilder() {
        for (int c = 0; c < 256; c++) lut[c] = af_code_of((char)c);
        lut[(int)' '] = lut[(int)'\t'] = lut[(int)'\r'] = lut[(int)'\n'] = 5;   // skipped
    }
    void room(int64_t more) {
        const size_t w = (size_t)((n + more + 15) >> 4) + 1, b = (size_t)((n + more + 31) >> 5) + 1;
        if (pk.size() < w) { if (pk.capacity() < w) pk.reserve(w + w / 2); pk.resize(w, 0u); }
        if (nm.size() < b) { if (nm.capacity() < b) nm.reserve(b + b / 2); nm.resize(b, 0u); }
    }
    inline void put(uint32_t c) {
        const int64_t x = n++;
        if (c == 4) { nm[(size_t)(x >> 5)] |= 1u << (x & 31); c = (n_word(x >> 4) >> (2 * (x & 15))) & 3u; }
        pk[(size_t)(x >> 4)] |= c << (2 * (x & 15));
    }
    void push_n(int64_t count) { room(count); for (int64_t i = 0; i < count; i++) put(4); }
    void push(const char *s, int64_t len) {
        room(len);
        for (int64_t i = 0; i < len; i++) { const uint32_t c = lut[(uint8_t)s[i]]; if (c != 5) put(c); }
    }
};

struct ContigList {
    GenomeBuilder b;
    std::vector<std::string> names;
    std::vector<int64_t> starts, lens;
    void begin(const std::string &name) {
        if (names.empty()) b.push_n(AF_GENOME_SEP);
        names.push_back(name); starts.push_back(b.n); lens.push_back(0);
    }
    void end() { lens.back() = b.n - starts.back(); b.push_n(AF_GENOME_SEP); }
};

static void finish(ContigList &c, af_genome_host &out) {
    out.n = c.b.n;
    out.pk.swap(c.b.pk); out.nm.swap(c.b.nm);
    out.names.swap(c.names); out.starts.swap(c.starts); out.lens.swap(c.lens);
}

int af_genome_host_from_contigs(const char *const *names, const char *const *seqs, const int64_t *lens, int32_t n, af_genome_host &out) {
    if (!names || !seqs || !lens || n <= 0) { af_set_error("af_genome_from_contigs: bad argument"); return AF_ERR_ARG; }
    ContigList c;
    for (int32_t i = 0; i < n; i++) {
        if (!names[i] || !seqs[i] || lens[i] < 0) { af_set_error("af_genome_from_contigs: contig %d is null", i); return AF_ERR_ARG; }
        c.begin(names[i]);
        c.b.push(seqs[i], lens[i]);
        c.end();
    }
    finish(c, out);
    return AF_OK;
}

static int parse_fasta(const char *path, af_genome_host &out) {
    gzFile f = gzopen(path, "rb");
    if (!f) { af_set_error("af_genome_from_fasta: cannot open %s", path); return AF_ERR_IO; }
    gzbuffer(f, 1 << 20);
    ContigList c;
    std::vector<char> buf((size_t)4 << 20);
    bool in_header = false, line_start = true, name_done = false, open = false;
    std::string name;
    int n;
    while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
        const char *p = buf.data(), *e = p + n;
        while (p < e) {
            if (in_header) {
                const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
                const char *stop = nl ? nl : e;
                for (const char *q = p; q < stop && !name_done; q++) {
                    if (*q == ' ' || *q == '\t' || *q == '\r') name_done = true; else name.push_back(*q);
                }
                if (nl) { in_header = false; line_start = true; c.begin(name); open = true; p = nl + 1; } else p = e;
                continue;
            }
            if (line_start && *p == '>') {
                if (open) { c.end(); open = false; }
                in_header = true; name.clear(); name_done = false; line_start = false; p++;
                continue;
            }
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *stop = nl ? nl : e;
            if (stop > p) {
                if (!open) { gzclose(f); af_set_error("af_genome_from_fasta: %s does not start with a '>' header", path); return AF_ERR_IO; }
                c.b.push(p, stop - p);
            }
            line_start = nl != nullptr;
            p = nl ? nl + 1 : e;
        }
    }
    int zerr = 0;
    const char *zmsg = gzerror(f, &zerr);
    if (n < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) {
        af_set_error("af_genome_from_fasta: %s: %s", path, zmsg ? zmsg : "read error");
        gzclose(f);
        return AF_ERR_IO;
    }
    gzclose(f);
    if (in_header) { c.begin(name); open = true; }
    if (open) c.end();
    if (c.names.empty()) { af_set_error("af_genome_from_fasta: %s holds no sequence", path); return AF_ERR_IO; }
    finish(c, out);
    return AF_OK;
}

// ---- packed-genome cache: <fasta>.af2bit beside the FASTA --------------------------------------------------------
// Parsing and packing a human genome takes ~9 s per process (330 Mbp/s); the packed form is 1.2 GB and loads at disk
// speed.  The cache is the host structure as it is: header {magic, version, FASTA size, FASTA mtime, bases, contigs},
// contig table, packed words, N words.  It is used only when its header matches the FASTA it sits next to and every
// size in it is consistent with the file's own length; anything else falls back to the FASTA (and rewrites the cache).
// AF_GENOME_CACHE=0 switches it off; an unwritable directory just means no cache.
static const uint64_t CACHE_MAGIC = 0x3154494232464101ull;   // "\x01AF2BIT1"
struct CacheHeader { uint64_t magic, version, fasta_size, fasta_mtime, n_bases, n_contigs, names_bytes, checksum; };

static uint64_t cache_checksum(const std::vector<int64_t> &tab, const std::string &names, const std::vector<uint32_t> &pk, const std::vector<uint32_t> &nm) {
    const uint64_t a = af_crc32((const uint8_t *)pk.data(), pk.size() * 4), b = af_crc32((const uint8_t *)nm.data(), nm.size() * 4);
    const uint64_t c = af_crc32((const uint8_t *)tab.data(), tab.size() * 8), d = af_crc32((const uint8_t *)names.data(), names.size());
    return (a | (b << 32)) ^ (c << 13) ^ (d << 29);
}

static bool fasta_stamp(const char *path, uint64_t *size, uint64_t *mtime) {
    struct stat st;
    if (stat(path, &st) != 0) return false;
    *size = (uint64_t)st.st_size; *mtime = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    return true;
}

static bool cache_load(const std::string &cpath, uint64_t fsize, uint64_t fmtime, af_genome_host &out) {
    FILE *f = fopen(cpath.c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    CacheHeader h;
    struct stat st;
    if (fread(&h, sizeof(h), 1, f) == 1 && h.magic == CACHE_MAGIC && h.version == 1 && h.fasta_size == fsize && h.fasta_mtime == fmtime &&
        h.n_bases < (1ull << 35) && h.n_contigs >= 1 && h.n_contigs < (1ull << 24) && h.names_bytes < (1ull << 30) && fstat(fileno(f), &st) == 0) {
        const uint64_t pkw = (h.n_bases + 15) / 16 + 1, nmw = (h.n_bases + 31) / 32 + 1;
        const uint64_t want = sizeof(h) + h.n_contigs * 16 + h.names_bytes + (pkw + nmw) * 4;
        if ((uint64_t)st.st_size == want) {
            std::vector<int64_t> tab((size_t)h.n_contigs * 2);
            std::string names((size_t)h.names_bytes, '\0');
            out.pk.resize((size_t)pkw); out.nm.resize((size_t)nmw);
            ok = fread(tab.data(), 16, (size_t)h.n_contigs, f) == h.n_contigs && (h.names_bytes == 0 || fread(&names[0], 1, names.size(), f) == names.size()) &&
                 fread(out.pk.data(), 4, (size_t)pkw, f) == pkw && fread(out.nm.data(), 4, (size_t)nmw, f) == nmw;
            if (ok && cache_checksum(tab, names, out.pk, out.nm) != h.checksum) ok = false;      // damaged body
            if (ok) {
                out.n = (int64_t)h.n_bases;
                out.names.clear(); out.starts.clear(); out.lens.clear();
                size_t at = 0;
                for (uint64_t i = 0; i < h.n_contigs && ok; i++) {
                    const size_t e = names.find('\0', at);
                    const int64_t start = tab[(size_t)i * 2], len = tab[(size_t)i * 2 + 1];
                    if (e == std::string::npos || start < AF_GENOME_SEP || len < 0 || start + len + AF_GENOME_SEP > out.n) { ok = false; break; }
                    out.names.push_back(names.substr(at, e - at)); out.starts.push_back(start); out.lens.push_back(len);
                    at = e + 1;
                }
            }
        }
    }
    fclose(f);
    return ok;
}

static void cache_save(const std::string &cpath, uint64_t fsize, uint64_t fmtime, const af_genome_host &g) {
    const std::string tmp = cpath + ".partial";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                           // read-only directory: no cache
    std::string names;
    std::vector<int64_t> tab;
    for (size_t i = 0; i < g.names.size(); i++) { names += g.names[i]; names.push_back('\0'); tab.push_back(g.starts[i]); tab.push_back(g.lens[i]); }
    const uint64_t pkw = (uint64_t)(g.n + 15) / 16 + 1, nmw = (uint64_t)(g.n + 31) / 32 + 1;
    std::vector<uint32_t> pk(g.pk), nm(g.nm);
    pk.resize((size_t)pkw, 0u); nm.resize((size_t)nmw, 0u);
    CacheHeader h = {CACHE_MAGIC, 1, fsize, fmtime, (uint64_t)g.n, (uint64_t)g.names.size(), (uint64_t)names.size(), cache_checksum(tab, names, pk, nm)};
    const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(tab.data(), 16, g.names.size(), f) == g.names.size() &&
                    (names.empty() || fwrite(names.data(), 1, names.size(), f) == names.size()) && fwrite(pk.data(), 4, (size_t)pkw, f) == pkw &&
                    fwrite(nm.data(), 4, (size_t)nmw, f) == nmw;
    if (fclose(f) == 0 && ok) rename(tmp.c_str(), cpath.c_str()); else remove(tmp.c_str());
}

int af_genome_host_from_fasta(const char *path, af_genome_host &out) {
    if (!path) { af_set_error("af_genome_from_fasta: null"); return AF_ERR_ARG; }
    const char *ce = getenv("AF_GENOME_CACHE");
    const bool use_cache = !(ce && ce[0] == '0');
    uint64_t fsize = 0, fmtime = 0;
    const std::string cpath = std::string(path) + ".af2bit";
    if (use_cache && fasta_stamp(path, &fsize, &fmtime) && cache_load(cpath, fsize, fmtime, out)) return AF_OK;
    out = af_genome_host();
    const int rc = parse_fasta(path, out);
    if (rc == AF_OK && use_cache && fsize) cache_save(cpath, fsize, fmtime, out);
    return rc;
}

// test hook: parse + pack on the host only; FNV-1a over (code 0..4 of every base of the concatenation) lets a test
// compare the packed genome with a sequence it built itself
extern "C" int af_debug_genome_fasta(const char *path, int64_t *total_len, int32_t *n_contigs, uint64_t *checksum) {
    af_genome_host h;
    const int rc = af_genome_host_from_fasta(path, h);
    if (rc) return rc;
    if (total_len) *total_len = h.n;
    if (n_contigs) *n_contigs = (int32_t)h.names.size();
    if (checksum) {
        uint64_t x = 1469598103934665603ull;
        for (int64_t i = 0; i < h.n; i++) {
            const uint32_t isn = (h.nm[(size_t)(i >> 5)] >> (i & 31)) & 1u;
            const uint32_t c = isn ? 4u : (h.pk[(size_t)(i >> 4)] >> (2 * (i & 15))) & 3u;
            x = (x ^ c) * 1099511628211ull;
        }
        *checksum = x;
    }
    return AF_OK;
}

