// af_genome_host.cpp -- host side of the genome pass (af_genome.cu): FASTA (plain / gzip) or in-memory contigs ->
// ONE 2-bit packed sequence [256 N] contig [256 N] contig ... [256 N] plus a 1-bit-per-base N map.  No CUDA here, so
// the parser of foreign files runs under AddressSanitizer / UBSan in tools/fuzz_host.cpp.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "af_common.h"

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

int af_genome_host_from_fasta(const char *path, af_genome_host &out) {
    if (!path) { af_set_error("af_genome_from_fasta: null"); return AF_ERR_ARG; }
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

