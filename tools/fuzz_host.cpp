// fuzz_host.cpp -- AddressSanitizer / UBSan run of the HOST side of the library (index build, packer,
// FASTQ reader), which parses files it did not write.  compute-sanitizer is closed on the GPU pool, so
// this covers the part of the code that can be sanitised here.  Built and run by tests/test_host_asan.py:
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined -Iinclude \
//       tools/fuzz_host.cpp anchored_fusion_b200/csrc/af_host.cpp anchored_fusion_b200/csrc/af_fastq.cpp -lz -lpthread
//   ./a.out <scratch dir> <iterations> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "anchored_fusion.h"

static std::mt19937_64 rng;
static int rnd(int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); }

static std::string random_fastq(int n, int mate, bool crlf) {
    std::string s;
    const char *nlr = crlf ? "\r\n" : "\n";
    for (int i = 0; i < n; i++) {
        int L = rnd(1, 180);
        s += "@r" + std::to_string(i) + (rnd(0, 1) ? "/" + std::to_string(mate) : "") + (rnd(0, 3) ? "" : " comment here") + nlr;
        for (int j = 0; j < L; j++) s += "ACGTNacgtn"[rnd(0, 9)];
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
        // ---- FASTQ reader on valid and damaged files ----
        const int n = rnd(0, 400);
        const bool crlf = !rnd(0, 4);
        std::string f1 = random_fastq(n, 1, crlf), f2 = random_fastq(n, 2, crlf);
        if (rnd(0, 2)) mutate(f1);
        if (!rnd(0, 3)) mutate(f2);
        const std::string p1 = dir + "/f_1.fastq", p2 = dir + "/f_2.fastq";
        write_file(p1, f1);
        write_file(p2, f2);
        int32_t peek = 0;
        af_fastq_peek(p1.c_str(), rnd(1, 50), &peek);
        af_fastq_t *fq = nullptr;
        if (af_fastq_open(p1.c_str(), p2.c_str(), &fq) != AF_OK) { failed++; continue; }
        const int mrl = rnd(0, 3) ? 192 : rnd(16, 256) / 16 * 16;
        const int64_t batch = rnd(1, 300);
        af_layout_t lay;
        af_layout(mrl, batch, &lay);
        std::vector<char> packed((size_t)lay.packed_bytes + 64);
        std::vector<uint16_t> lens((size_t)(2 * batch));
        std::vector<uint32_t> nids((size_t)(2 * batch)), nmask((size_t)(2 * batch) * AF_NMASK_WORDS);
        for (;;) {
            int64_t n_n = 0, got = 0;
            int32_t ulen = 0;
            int rc = af_fastq_next(fq, batch, mrl, 0xE4, packed.data(), lens.data(), nids.data(), nmask.data(), (int64_t)nids.size(),
                                   &n_n, &ulen, &got);
            if (rc != AF_OK) { failed++; break; }
            if (got == 0) { ok++; break; }
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
        af_fastq_close(fq);
    }
    printf("fuzz_host: %d iterations, %ld clean, %ld rejected, %ld record pairs read\n", iters, ok, failed, records);
    return 0;
}
