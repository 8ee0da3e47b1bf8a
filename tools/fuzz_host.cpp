// fuzz_host.cpp -- AddressSanitizer / UBSan run of the HOST side of the library (index build, packer,
// FASTQ reader), which parses files it did not write.  compute-sanitizer is closed on the GPU pool, so
// this covers the part of the code that can be sanitised here.  Built and run by tests/test_host_asan.py:
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined -Iinclude \
//       tools/fuzz_host.cpp anchored_fusion_b200/csrc/af_host.cpp anchored_fusion_b200/csrc/af_fastq.cpp \
//       anchored_fusion_b200/csrc/af_genome_host.cpp -lz -lpthread
// (af_fastq.cpp pulls in af_inflate.h, the DEFLATE decoder: damaged gzip / BGZF bytes go through it)
//   ./a.out <scratch dir> <iterations> <seed>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "anchored_fusion.h"

static std::mt19937_64 rng;
static int rnd(int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); }

static std::vector<std::string> g_expect[2];   // bases of every record of the last two generated files
static std::string random_fastq(int n, int mate, bool crlf) {
    std::string s;
    g_expect[mate - 1].clear();
    const char *nlr = crlf ? "\r\n" : "\n";
    for (int i = 0; i < n; i++) {
        int L = rnd(1, 180);
        s += "@r" + std::to_string(i) + (rnd(0, 1) ? "/" + std::to_string(mate) : "") + (rnd(0, 3) ? "" : " comment here") + nlr;
        std::string bases;
        for (int j = 0; j < L; j++) bases += "ACGTNacgtn"[rnd(0, 9)];
        g_expect[mate - 1].push_back(bases);
        s += bases;
        s += nlr;
        s += rnd(0, 4) ? "+" : "+again";
        s += nlr;
        for (int j = 0; j < L; j++) s += (char)rnd(33, 73);
        s += nlr;
        if (!rnd(0, 9)) s += nlr;
    }
    return s;
}

static void mutate(std::string &s) {
    int kind = rnd(0, 5);
    if (s.empty() || kind == 0) return;
    size_t p = (size_t)(rng() % s.size());
    if (kind == 1) s.resize(p);                                            // truncation
    else if (kind == 2) s[p] = (char)rnd(0, 255);                          // byte flip
    else if (kind == 3) s.insert(p, std::string((size_t)rnd(1, 5), '\n')); // blank lines
    else if (kind == 4) s.erase(p, (size_t)rnd(1, 300));                   // a hole
    else s.insert(p, std::string((size_t)rnd(1, 400), "@+ACGT\n"[rnd(0, 6)]));
}

// the text as one gzip member (zlib), as several members, or as BGZF blocks (raw deflate + 'BC' extra field)
static std::string gz_member(const std::string &t, int level) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, 31, 8, rnd(0, 4) ? Z_DEFAULT_STRATEGY : (rnd(0, 1) ? Z_FIXED : Z_HUFFMAN_ONLY));
    std::string out(deflateBound(&zs, (uLong)t.size()) + 64, '\0');
    zs.next_in = (Bytef *)t.data(); zs.avail_in = (uInt)t.size();
    zs.next_out = (Bytef *)&out[0]; zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}
static std::string bgzf(const std::string &t) {
    std::string out;
    const size_t step = (size_t)rnd(200, 0xFF00);
    for (size_t i = 0; i < t.size() || i == 0; i += step) {
        const std::string blk = t.substr(i, step);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        deflateInit2(&zs, rnd(0, 9), Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        std::string comp(deflateBound(&zs, (uLong)blk.size()) + 64, '\0');
        zs.next_in = (Bytef *)blk.data(); zs.avail_in = (uInt)blk.size();
        zs.next_out = (Bytef *)&comp[0]; zs.avail_out = (uInt)comp.size();
        deflate(&zs, Z_FINISH);
        comp.resize(zs.total_out);
        deflateEnd(&zs);
        const unsigned bsize = (unsigned)comp.size() + 25;
        const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 255), (unsigned char)(bsize >> 8)};
        out.append((const char *)hdr, 18);
        out += comp;
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)blk.data(), (uInt)blk.size()), isz = (uint32_t)blk.size();
        for (int k = 0; k < 4; k++) out += (char)(crc >> (8 * k));
        for (int k = 0; k < 4; k++) out += (char)(isz >> (8 * k));
        if (t.empty()) break;
    }
    return out;
}
static std::string encode(const std::string &t, int kind) {
    if (kind == 1) return gz_member(t, rnd(0, 9));
    if (kind == 2) { std::string o; for (size_t i = 0; i < t.size() || i == 0; i += 3000) { o += gz_member(t.substr(i, 3000), rnd(1, 9)); if (t.empty()) break; } return o; }
    if (kind == 3) return bgzf(t);
    return t;
}

static void write_file(const std::string &path, const std::string &s) {
    FILE *f = fopen(path.c_str(), "wb");
    fwrite(s.data(), 1, s.size(), f);
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <dir> <iterations> <seed>\n", argv[0]); return 2; }
    const std::string dir = argv[1];
    const int iters = atoi(argv[2]);
    rng.seed((uint64_t)atoll(argv[3]));
    long ok = 0, failed = 0, records = 0;
    for (int it = 0; it < iters; it++) {
        // ---- index build on random anchors (N runs, tiny and odd lengths) ----
        {
            int G = rnd(1, 4000);
            std::string a((size_t)G, 'A');
            for (auto &c : a) c = "ACGTN"[rnd(0, 19) ? rnd(0, 3) : 4];
            af_index_t *idx = nullptr;
            int rc = af_index_build(a.data(), G, nullptr, rnd(0, 1) ? 12 : 13, &idx);
            if (rc == AF_OK) { af_index_info_t info; af_index_info(idx, &info); af_index_free(idx); }
        }
        // ---- genome FASTA loader (host half of af_genome_from_fasta) on valid and damaged files ----
        {
            const int nc = rnd(0, 6);
            const bool crlf_fa = !rnd(0, 3);
            std::string fa;
            int64_t want_len = AF_GENOME_SEP;
            for (int c = 0; c < nc; c++) {
                fa += ">ctg" + std::to_string(c) + (rnd(0, 2) ? "" : " some description") + (crlf_fa ? "\r\n" : "\n");
                const int len = rnd(0, 3) ? rnd(0, 3000) : 0, width = rnd(1, 120);
                for (int i = 0; i < len; i++) {
                    fa += "ACGTNacgtnRYKM*-"[rnd(0, 15)];
                    if ((i + 1) % width == 0 || i + 1 == len) fa += crlf_fa ? "\r\n" : "\n";
                }
                want_len += len + AF_GENOME_SEP;
            }
            bool intact = true;
            if (!rnd(0, 2)) { mutate(fa); intact = false; }
            if (rnd(0, 1)) { fa = encode(fa, rnd(1, 2)); if (!rnd(0, 2)) { mutate(fa); intact = false; } }
            const std::string pf = dir + "/g.fa";
            write_file(pf, fa);
            int64_t total = 0; int32_t ncon = 0; uint64_t sum = 0;
            setenv("AF_GENOME_CACHE", rnd(0, 2) ? "0" : "1", 1);
            remove((pf + ".af2bit").c_str());
            int rc = af_debug_genome_fasta(pf.c_str(), &total, &ncon, &sum);
            if (rc == AF_OK && !rnd(0, 1)) {                           // a second load through a (possibly damaged) cache file
                FILE *cf = fopen((pf + ".af2bit").c_str(), "rb");
                if (cf) {
                    std::string cb;
                    char tmp[4096]; size_t got;
                    while ((got = fread(tmp, 1, sizeof(tmp), cf)) > 0) cb.append(tmp, got);
                    fclose(cf);
                    const bool damage = rnd(0, 1);
                    if (damage) { mutate(cb); write_file(pf + ".af2bit", cb); }
                    int64_t total2 = 0; int32_t ncon2 = 0; uint64_t sum2 = 0;
                    const int rc2 = af_debug_genome_fasta(pf.c_str(), &total2, &ncon2, &sum2);
                    if (rc2 != AF_OK || total2 != total || ncon2 != ncon || sum2 != sum) {
                        fprintf(stderr, "fuzz_host: genome cache (damaged: %d) changed the result in iteration %d\n", (int)damage, it);
                        return 1;
                    }
                }
            }
            if (intact && nc > 0 && (rc != AF_OK || ncon != nc || total != want_len)) {
                fprintf(stderr, "fuzz_host: FASTA of %d contigs / %lld bases came back as rc %d, %d contigs, %lld bases (iteration %d)\n",
                        nc, (long long)want_len, rc, ncon, (long long)total, it);
                return 1;
            }
        }
        // ---- FASTQ reader on valid and damaged files ----
        const int n = rnd(0, 400);
        const bool crlf = !rnd(0, 4);
        std::string f1 = random_fastq(n, 1, crlf), f2 = random_fastq(n, 2, crlf);
        bool pristine = true;
        if (rnd(0, 2)) { mutate(f1); pristine = false; }
        if (!rnd(0, 3)) { mutate(f2); pristine = false; }
        // plain text, gzip, several gzip members or BGZF; then sometimes damage the COMPRESSED bytes
        // (deflate bit stream, block sizes in the BGZF extra field, CRCs, truncation)
        f1 = encode(f1, rnd(0, 3));
        f2 = encode(f2, rnd(0, 3));
        if (!rnd(0, 2)) { mutate(f1); pristine = false; }
        if (!rnd(0, 5)) { mutate(f2); pristine = false; }
        const std::string p1 = dir + "/f_1.fastq", p2 = dir + "/f_2.fastq";
        write_file(p1, f1);
        write_file(p2, f2);
        int32_t peek = 0;
        af_fastq_peek(p1.c_str(), rnd(1, 50), &peek);
        af_fastq_t *fq = nullptr;
        int orc, copies = 1;
        if (rnd(0, 3)) orc = af_fastq_open_threads(p1.c_str(), p2.c_str(), rnd(1, 4), &fq);
        else {                                                            // the same pair three times: the cell layout
            const char *a[3] = {p1.c_str(), p1.c_str(), p1.c_str()}, *b[3] = {p2.c_str(), p2.c_str(), p2.c_str()};
            orc = af_fastq_open_multi(a, b, 3, rnd(1, 4), &fq);
            copies = 3;
        }
        if (orc != AF_OK) { failed++; continue; }
        const int mrl = rnd(0, 3) ? 192 : rnd(16, 256) / 16 * 16;
        const int64_t batch = rnd(1, 300);
        af_layout_t lay;
        af_layout(mrl, batch, &lay);
        std::vector<char> packed((size_t)lay.packed_bytes + 64);
        std::vector<uint16_t> lens((size_t)(2 * batch));
        std::vector<uint32_t> nids((size_t)(2 * batch)), nmask((size_t)(2 * batch) * AF_NMASK_WORDS);
        int64_t seen = 0;
        bool complete = false;
        for (;;) {
            int64_t n_n = 0, got = 0;
            int32_t ulen = 0;
            int rc = rnd(0, 9) ? af_fastq_next(fq, batch, mrl, 0xE4, packed.data(), lens.data(), nids.data(), nmask.data(), (int64_t)nids.size(),
                                               &n_n, &ulen, &got)
                               : af_fastq_skip(fq, batch, &got);
            if (rc != AF_OK) { failed++; break; }
            if (got == 0) { ok++; complete = true; break; }
            if (pristine) {                                                // undamaged input: every record must come back
                for (int64_t rid = 0; rid < 2 * got; rid++) {
                    const char *name, *seq, *qual;
                    int32_t nl = 0, len = 0;
                    const std::vector<std::string> &ex = g_expect[rid & 1];
                    const size_t at = (size_t)((seen + (rid >> 1)) % (int64_t)ex.size());
                    if (af_fastq_record(fq, rid, &name, &nl, &seq, &qual, &len) != AF_OK || std::string(seq, (size_t)len) != ex[at]) {
                        fprintf(stderr, "fuzz_host: WRONG RECORD %lld of iteration %d\n", (long long)(seen * 2 + rid), it);
                        return 1;
                    }
                }
            }
            seen += got;
            records += got;
            for (int k = 0; k < 4; k++) {
                const char *name, *seq, *qual;
                int32_t nl = 0, len = 0;
                if (af_fastq_record(fq, (int64_t)(rng() % (uint64_t)(2 * got)), &name, &nl, &seq, &qual, &len) == AF_OK) {
                    volatile char sink = 0;
                    for (int j = 0; j < nl; j++) sink ^= name[j];
                    for (int j = 0; j < len; j++) sink ^= (char)(seq[j] ^ qual[j]);
                    (void)sink;
                }
            }
            if (!rnd(0, 15)) break;                                        // close mid-file
        }
        if (pristine && complete && seen != (int64_t)copies * n) { fprintf(stderr, "fuzz_host: %lld of %d pairs read in iteration %d\n", (long long)seen, copies * n, it); return 1; }
        af_fastq_close(fq);
    }
    printf("fuzz_host: %d iterations, %ld clean, %ld rejected, %ld record pairs read\n", iters, ok, failed, records);
    return 0;
}
