// af_fastq.cpp -- paired FASTQ / FASTQ.gz reader feeding the 2-bit packer.
// Replaces the kseq/zlib ingest that happens inside `bwa mem ... fastq1 fastq2`
// (Anchored_Fusion.py:182).  One zlib decode thread per file; the text of the current batch
// is kept so that names / bases / qualities of the (few) anchored reads can be written out.
#include <zlib.h>

#include <cstring>
#include <thread>

#include "af_common.h"

struct SeqRef { const char *p; int32_t len; };
struct PackSide {
    std::vector<uint32_t> nids, nmask;
    int32_t ulen = -1;
    int rc = AF_OK;
    std::string err;
};
void af_pack_side(const SeqRef *r, int m, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                  uint16_t *lens_out, PackSide &st);
int af_pack_finish(PackSide &a, PackSide &b, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                   int64_t *n_nreads_out, int32_t *uniform_len_out);

namespace {

struct Rec { int64_t name_off, seq_off, qual_off; int32_t name_len, len; };

struct Side {
    gzFile gz = nullptr;
    std::vector<char> in;      // decode buffer
    size_t in_pos = 0, in_end = 0;
    bool eof = false;
    std::vector<char> text;    // records of the current batch (names, bases, quals)
    std::vector<Rec> recs;
    std::string err;

    bool fill() {
        if (eof) return false;
        if (in_pos > 0) { memmove(in.data(), in.data() + in_pos, in_end - in_pos); in_end -= in_pos; in_pos = 0; }
        if (in_end == in.size()) in.resize(in.size() * 2);
        int n = gzread(gz, in.data() + in_end, (unsigned)(in.size() - in_end));
        if (n < 0) { int e; err = gzerror(gz, &e); eof = true; return false; }
        if (n == 0) { eof = true; return false; }
        in_end += (size_t)n;
        return true;
    }
    // next line without its terminator; false at EOF with nothing left
    bool line(const char *&p, size_t &len) {
        for (;;) {
            char *s = in.data() + in_pos;
            char *nl = (char *)memchr(s, '\n', in_end - in_pos);
            if (nl) {
                len = (size_t)(nl - s);
                in_pos += len + 1;
                if (len && s[len - 1] == '\r') len--;
                p = s;
                return true;
            }
            if (!fill()) {
                if (in_pos < in_end) { p = in.data() + in_pos; len = in_end - in_pos; in_pos = in_end; return true; }
                return false;
            }
        }
    }
    void read_batch(int64_t max_pairs) {
        text.clear();
        recs.clear();
        const char *p;
        size_t n;
        while ((int64_t)recs.size() < max_pairs) {
            if (!line(p, n)) break;
            if (n == 0) continue;  // stray blank line between records
            if (p[0] != '@') { err = "FASTQ record does not start with '@'"; return; }
            Rec r;
            // name: up to the first blank; a trailing /1 or /2 is dropped, as bwa does
            size_t nl = 1;
            while (nl < n && p[nl] != ' ' && p[nl] != '\t') nl++;
            size_t name_len = nl - 1;
            if (name_len >= 2 && p[nl - 2] == '/' && (p[nl - 1] == '1' || p[nl - 1] == '2')) name_len -= 2;
            r.name_off = (int64_t)text.size();
            r.name_len = (int32_t)name_len;
            text.insert(text.end(), p + 1, p + 1 + name_len);
            if (!line(p, n)) { err = "truncated FASTQ record"; return; }
            r.seq_off = (int64_t)text.size();
            r.len = (int32_t)n;
            text.insert(text.end(), p, p + n);
            if (!line(p, n) || n == 0 || p[0] != '+') { err = "FASTQ record lacks its '+' line"; return; }
            if (!line(p, n)) { err = "truncated FASTQ record"; return; }
            if ((int32_t)n != r.len) { err = "FASTQ quality length differs from sequence length"; return; }
            r.qual_off = (int64_t)text.size();
            text.insert(text.end(), p, p + n);
            recs.push_back(r);
        }
    }
};

}  // namespace

struct af_fastq {
    Side side[2];
    int64_t n_cur = 0;
};

extern "C" int af_fastq_open(const char *path1, const char *path2, af_fastq_t **out) {
    if (!path1 || !path2 || !out) { af_set_error("af_fastq_open: null argument"); return AF_ERR_ARG; }
    af_fastq *fq = new af_fastq();
    const char *paths[2] = {path1, path2};
    for (int i = 0; i < 2; i++) {
        fq->side[i].gz = gzopen(paths[i], "rb");  // transparently reads plain text too
        if (!fq->side[i].gz) {
            af_set_error("af_fastq_open: cannot open %s", paths[i]);
            for (int j = 0; j < i; j++) gzclose(fq->side[j].gz);
            delete fq;
            return AF_ERR_IO;
        }
        gzbuffer(fq->side[i].gz, 1 << 20);
        fq->side[i].in.resize(4 << 20);
    }
    *out = fq;
    return AF_OK;
}

extern "C" void af_fastq_close(af_fastq_t *fq) {
    if (!fq) return;
    for (int i = 0; i < 2; i++) if (fq->side[i].gz) gzclose(fq->side[i].gz);
    delete fq;
}

extern "C" int af_fastq_next(af_fastq_t *fq, int64_t max_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                             uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                             int64_t *n_nreads_out, int32_t *uniform_len_out, int64_t *n_pairs_out) {
    if (!fq || !n_pairs_out || max_pairs <= 0) { af_set_error("af_fastq_next: bad argument"); return AF_ERR_ARG; }
    // each side: inflate + parse its file, then pack its mate's words (the mates own disjoint words)
    PackSide ps[2];
    auto work = [&](int m) {
        Side &sd = fq->side[m];
        sd.read_batch(max_pairs);
        if (!sd.err.empty() || !packed_out) return;
        std::vector<SeqRef> refs(sd.recs.size());
        for (size_t i = 0; i < sd.recs.size(); i++) refs[i] = {sd.text.data() + sd.recs[i].seq_off, sd.recs[i].len};
        af_pack_side(refs.data(), m, (int64_t)refs.size(), max_read_len, pad_byte, packed_out, lens_out, ps[m]);
    };
    std::thread t1(work, 1);
    work(0);
    t1.join();
    for (int i = 0; i < 2; i++)
        if (!fq->side[i].err.empty()) { af_set_error("af_fastq_next: file %d: %s", i + 1, fq->side[i].err.c_str()); return AF_ERR_IO; }
    if (fq->side[0].recs.size() != fq->side[1].recs.size()) {
        af_set_error("af_fastq_next: the two FASTQ files are out of step (%zu vs %zu records)", fq->side[0].recs.size(), fq->side[1].recs.size());
        return AF_ERR_IO;
    }
    int64_t n = (int64_t)fq->side[0].recs.size();
    fq->n_cur = n;
    *n_pairs_out = n;
    if (n == 0) { if (n_nreads_out) *n_nreads_out = 0; if (uniform_len_out) *uniform_len_out = 0; return AF_OK; }
    if (!packed_out) { af_set_error("af_fastq_next: packed_out is null"); return AF_ERR_ARG; }
    return af_pack_finish(ps[0], ps[1], nread_ids_out, nmask_out, ncap, n_nreads_out, uniform_len_out);
}

extern "C" int af_fastq_record(const af_fastq_t *fq, int64_t read_id, const char **name, int32_t *name_len,
                               const char **seq, const char **qual, int32_t *len) {
    if (!fq || read_id < 0 || (read_id >> 1) >= fq->n_cur) { af_set_error("af_fastq_record: read_id out of range"); return AF_ERR_ARG; }
    const Side &s = fq->side[read_id & 1];
    const Rec &r = s.recs[(size_t)(read_id >> 1)];
    if (name) *name = s.text.data() + r.name_off;
    if (name_len) *name_len = r.name_len;
    if (seq) *seq = s.text.data() + r.seq_off;
    if (qual) *qual = s.text.data() + r.qual_off;
    if (len) *len = r.len;
    return AF_OK;
}
