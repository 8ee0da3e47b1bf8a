// af_fastq.cpp -- paired FASTQ / FASTQ.gz reader feeding the 2-bit packer.
// Replaces the kseq/zlib ingest that happens inside `bwa mem ... fastq1 fastq2`
// (Anchored_Fusion.py:182).  Per file one inflate thread runs ahead of the consumer through a
// bounded queue of decoded blocks, and one parse+pack thread per file works inside af_fastq_next,
// so inflate (the wall for .gz input), record parsing and 2-bit packing overlap, and decoding goes
// on while the caller is busy with the previous batch.  The text of the current batch is kept so
// that names / bases / qualities of the (few) anchored reads can be written out.
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
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

static const size_t BLOCK_BYTES = 4u << 20;   // decoded text per queue entry (the first entries are smaller:
static const size_t FIRST_BLOCK = 256u << 10; //   a single-cell file of a few thousand reads should not pay for 4 MB blocks)
static const size_t QUEUE_BLOCKS = 16;        // read-ahead bound per file (64 MB of text)

struct Side {
    gzFile gz = nullptr;
    std::vector<char> in;      // parse buffer: decoded text not yet consumed
    size_t in_pos = 0, in_end = 0;
    bool eof = false;
    std::vector<char> text;    // records of the current batch (names, bases, quals)
    std::vector<Rec> recs;
    std::string err;

    // inflate thread -> consumer
    std::thread inflater;
    std::mutex mu;
    std::condition_variable cv_data, cv_room;
    std::deque<std::vector<char>> ready;
    std::vector<std::vector<char>> spare;   // recycled blocks
    bool done = false, quit = false;
    std::string inflate_err;

    void inflate_loop() {
        size_t want = FIRST_BLOCK;
        for (;;) {
            std::vector<char> blk;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_room.wait(lk, [&] { return quit || ready.size() < QUEUE_BLOCKS; });
                if (quit) return;
                if (!spare.empty()) { blk.swap(spare.back()); spare.pop_back(); }
            }
            if (blk.size() < want) blk.resize(want);
            int n = gzread(gz, blk.data(), (unsigned)want);
            if (want < BLOCK_BYTES) want *= 2;
            std::lock_guard<std::mutex> lk(mu);
            if (n <= 0) {
                if (n < 0) { int e; inflate_err = gzerror(gz, &e); }
                done = true;
                cv_data.notify_all();
                return;
            }
            blk.resize((size_t)n);
            ready.push_back(std::move(blk));
            cv_data.notify_one();
        }
    }
    void start() { inflater = std::thread([this] { inflate_loop(); }); }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv_room.notify_all();
        if (inflater.joinable()) inflater.join();
    }

    bool fill() {
        if (eof) return false;
        std::vector<char> blk;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_data.wait(lk, [&] { return done || !ready.empty(); });
            if (ready.empty()) {
                if (!inflate_err.empty()) err = inflate_err;
                eof = true;
                return false;
            }
            blk.swap(ready.front());
            ready.pop_front();
        }
        cv_room.notify_one();
        if (in_pos > 0) { memmove(in.data(), in.data() + in_pos, in_end - in_pos); in_end -= in_pos; in_pos = 0; }
        if (in_end + blk.size() > in.size()) in.resize(std::max(in.size() * 2, in_end + blk.size()));
        memcpy(in.data() + in_end, blk.data(), blk.size());
        in_end += blk.size();
        {
            std::lock_guard<std::mutex> lk(mu);
            if (spare.size() < QUEUE_BLOCKS) spare.push_back(std::move(blk));
        }
        return true;
    }
    // Finds the next record (4 lines, terminators stripped) in the parse buffer, pulling decoded blocks
    // as needed; b[k] / e[k] delimit line k, `next` is where the following record starts.  Returns
    // false at the end of the input (err is set when a record is cut short).
    bool next_record(size_t b[4], size_t e[4], size_t &next) {
        for (;;) {
            size_t cur = in_pos;
            int k = 0;
            bool need_more = false;
            while (k < 4) {
                const char *nl = cur < in_end ? (const char *)memchr(in.data() + cur, '\n', in_end - cur) : nullptr;
                size_t end;
                if (nl) end = (size_t)(nl - in.data());
                else if (eof && cur < in_end) end = in_end;           // last line without a terminator
                else { need_more = true; break; }
                const size_t le = (end > cur && in[end - 1] == '\r') ? end - 1 : end;
                const size_t after = end < in_end ? end + 1 : in_end;
                if (k == 0 && le == cur) { in_pos = cur = after; continue; }   // stray blank line between records
                b[k] = cur; e[k] = le; k++;
                cur = after;
            }
            if (!need_more) { next = cur; return true; }
            if (eof) {
                if (k > 0) err = "truncated FASTQ record";
                return false;
            }
            fill();   // compacts the buffer (in_pos -> 0) and appends a block, or sets eof; rescan either way
        }
    }
    void read_batch(int64_t max_pairs) {
        const size_t last = text.size();
        text.clear();
        text.reserve(last);
        recs.clear();
        // records are parsed in place in the parse buffer, then name, bases and qualities are copied
        // back to back into `text`
        size_t b[4], e[4], next;
        while ((int64_t)recs.size() < max_pairs && next_record(b, e, next)) {
            const char *p = in.data();
            if (p[b[0]] != '@') { err = "FASTQ record does not start with '@'"; return; }
            if (e[2] == b[2] || p[b[2]] != '+') { err = "FASTQ record lacks its '+' line"; return; }
            const size_t slen = e[1] - b[1];
            if (e[3] - b[3] != slen) { err = "FASTQ quality length differs from sequence length"; return; }
            // name: up to the first blank; a trailing /1 or /2 is dropped, as bwa does
            size_t nl = b[0] + 1;
            while (nl < e[0] && p[nl] != ' ' && p[nl] != '\t') nl++;
            size_t name_len = nl - (b[0] + 1);
            if (name_len >= 2 && p[nl - 2] == '/' && (p[nl - 1] == '1' || p[nl - 1] == '2')) name_len -= 2;
            Rec r;
            r.name_off = (int64_t)text.size();
            r.name_len = (int32_t)name_len;
            r.len = (int32_t)slen;
            r.seq_off = r.name_off + (int64_t)name_len;
            r.qual_off = r.seq_off + (int64_t)slen;
            const size_t at = text.size();
            text.resize(at + name_len + 2 * slen);
            char *dst = text.data() + at;
            memcpy(dst, p + b[0] + 1, name_len);
            memcpy(dst + name_len, p + b[1], slen);
            memcpy(dst + name_len + slen, p + b[3], slen);
            recs.push_back(r);
            in_pos = next;
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
        fq->side[i].in.resize(2 * FIRST_BLOCK);
    }
    for (int i = 0; i < 2; i++) fq->side[i].start();
    *out = fq;
    return AF_OK;
}

// Longest read among the first n_records records of a FASTQ / FASTQ.gz file (the reader needs the
// packed width of a read before it starts; the Python gzip loop this replaces cost 5 ms per file,
// which is most of the per-cell time of a single-cell run).
extern "C" int af_fastq_peek(const char *path, int32_t n_records, int32_t *max_len_out) {
    if (!path || !max_len_out || n_records <= 0) { af_set_error("af_fastq_peek: bad argument"); return AF_ERR_ARG; }
    gzFile gz = gzopen(path, "rb");
    if (!gz) { af_set_error("af_fastq_peek: cannot open %s", path); return AF_ERR_IO; }
    // inflate in 256 KB steps and stop as soon as n_records records have been seen (at most 4 MB of text)
    const size_t STEP = 256u << 10, LIMIT = 4u << 20;
    std::vector<char> buf;
    size_t n = 0, pos = 0, line_no = 0;
    int32_t longest = 0, rec = 0;
    bool at_eof = false;
    while (rec < n_records) {
        const char *nl = pos < n ? (const char *)memchr(buf.data() + pos, '\n', n - pos) : nullptr;
        if (!nl && !at_eof) {                            // need more text
            if (n >= LIMIT) break;
            buf.resize(n + STEP);
            const int got = gzread(gz, buf.data() + n, (unsigned)STEP);
            if (got <= 0) at_eof = true; else n += (size_t)got;
            continue;
        }
        if (!nl && pos >= n) break;
        const size_t end = nl ? (size_t)(nl - buf.data()) : n;   // last line of the file may lack its terminator
        size_t len = end - pos;
        if (len && buf[end - 1] == '\r') len--;
        pos = end + 1;
        if (len == 0 && line_no % 4 == 0) continue;      // blank line between records
        if (line_no % 4 == 1) longest = std::max<int32_t>(longest, (int32_t)len);
        if (line_no % 4 == 3) rec++;
        line_no++;
    }
    gzclose(gz);
    *max_len_out = longest;
    return AF_OK;
}

extern "C" void af_fastq_close(af_fastq_t *fq) {
    if (!fq) return;
    for (int i = 0; i < 2; i++) fq->side[i].stop();
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
