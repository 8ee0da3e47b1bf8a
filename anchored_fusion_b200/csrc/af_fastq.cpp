// af_fastq.cpp -- paired FASTQ / FASTQ.gz / BGZF reader feeding the 2-bit packer.
// Replaces the kseq/zlib ingest that happens inside `bwa mem -t T ... fastq1 fastq2`
// (Anchored_Fusion.py:182; `--thread`, Anchored_Fusion.py:29, is the worker count here as it is bwa's).
//
// Round 1 read each file with gzread on one thread (1.2 M pairs/s, 13x below the CPU port's compute
// rate).  This reader is a small task-parallel runtime:
//
//   * every input file is mmap-ed; a DRIVER thread per mate walks its file(s) and cuts the decoded
//     text into SEGMENTS (contiguous buffers of ~8 MB, each preceded by a copy of the last 64 KB of
//     the previous one: the deflate window and the head of a record that straddles the cut);
//   * BGZF input (bgzip; block sizes are in the gzip extra field) is inflated block-group by
//     block-group on the worker POOL, plain text is copied the same way, a plain gzip member -- one
//     serial bit stream -- is inflated by the driver itself with the decoder in af_inflate.h
//     (2-2.5x zlib), and many small files (single-cell runs: one pair of files per cell) become one
//     task per file, so thousands of cells decode concurrently and land in the same packed batches;
//   * the same tasks index the newlines of what they decoded and check the members' CRC-32 slices;
//   * af_fastq_next assembles a batch from the finalised segments of both mates and fans the record
//     parse + 2-bit pack out over the pool (tile-aligned pair ranges; the two mates own disjoint words).
//
// Records are never copied: names / bases / qualities of the (few) anchored reads are looked up in the
// segments of the current batch, which stay alive until the next af_fastq_next.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "af_common.h"
#include "af_inflate.h"
#include "af_inflate_par.h"
#include "af_crc32.h"

struct SeqRef { const char *p; int32_t len; };
struct PackSide {
    std::vector<uint32_t> nids, nmask;
    int32_t ulen = -1;
    int rc = AF_OK;
    std::string err;
};
void af_pack_side(const SeqRef *r, int m, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                  uint16_t *lens_out, PackSide &st);
void af_pack_range(const SeqRef *r, int m, int64_t p0, int64_t p1, int64_t n_pairs, int32_t max_read_len, int32_t pad_byte,
                   void *packed_out, uint16_t *lens_out, PackSide &st);
int af_pack_finish(PackSide &a, PackSide &b, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                   int64_t *n_nreads_out, int32_t *uniform_len_out);

namespace {

static const size_t HEAD = 64u << 10;          // bytes of the previous segment kept in front of a segment's text
#ifdef AF_FASTQ_TEST_SIZES                     // sanitizer / fuzz builds: many segments and slices out of small files
static const size_t SEG_TEXT = 24u << 10, SLICE = 5000, SMALL_FILE = 2000;
#else
static const size_t SEG_TEXT = 8u << 20;       // decoded text per segment
static const size_t SLICE = 1u << 20;          // newline-index / CRC / copy granularity inside a segment
static const size_t SMALL_FILE = 3u << 20;     // files up to this size (on disk) are one task each
#endif
static const size_t AHEAD_BYTES = 192u << 20;  // decoded text a driver may hold ready ahead of the consumer
// one gzip member decoded by several workers (af_inflate_par.h): compressed bytes per chunk, and the size from which
// a member is worth it
#ifdef AF_FASTQ_TEST_SIZES
static const size_t PAR_CHUNK_DEFAULT = 6000;
static const int PAR_MIN_WORKERS = 2;
#else
static const size_t PAR_CHUNK_DEFAULT = 8u << 20;
static const int PAR_MIN_WORKERS = 6;              // below that the two serial inflate threads are as fast (measured)
#endif
// AF_GZIP_PAR_CHUNK=<bytes> (tests): chunk size, which also makes members from 4 chunks on and pools from 2 workers on eligible
static size_t par_chunk_env() { const char *v = getenv("AF_GZIP_PAR_CHUNK"); const long long x = v ? atoll(v) : 0; return x >= 2048 ? (size_t)x : 0; }

// cell buffers of the parallel gzip path, recycled (2 bytes per decoded byte: fresh pages would cost more than the decode)
struct CellCache {
    std::mutex mu;
    std::vector<afz::CellBuf> free_list;
    afz::CellBuf get() {
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.empty()) return afz::CellBuf();
        afz::CellBuf b = std::move(free_list.back());
        free_list.pop_back();
        return b;
    }
    void put(afz::CellBuf &&b) {
        b.clear();
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.size() < 24) free_list.push_back(std::move(b));
    }
    void trim() {                        // a reader was closed: hand the memory back (another open reader will allocate again)
        std::vector<afz::CellBuf> drop;
        { std::lock_guard<std::mutex> lk(mu); drop.swap(free_list); }
    }
};
static CellCache g_cells;

// ---- worker pool ---------------------------------------------------------------------------------
class Pool {
public:
    explicit Pool(int n) {
        if (n < 1) n = 1;
        for (int i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); quit_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void submit(std::function<void()> f) {
        { std::lock_guard<std::mutex> lk(mu_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
    int size() const { return (int)th_.size(); }

private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || !q_.empty(); });
                if (q_.empty()) return;            // quit, queue drained
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> th_;
    std::deque<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool quit_ = false;
};

// counts outstanding tasks; wait() blocks until all are done
struct Latch {
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    void add(int n = 1) { std::lock_guard<std::mutex> lk(mu); pending += n; }
    void done() { std::lock_guard<std::mutex> lk(mu); if (--pending == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return pending == 0; }); }
};

// ---- input files ---------------------------------------------------------------------------------
struct MappedFile {
    const uint8_t *p = nullptr;
    size_t n = 0;
    bool mapped = false;
    ~MappedFile() { if (mapped && p) munmap((void *)p, n); }
    bool open(const char *path, std::string &err) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) { err = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); err = std::string("cannot open ") + path + " (not a regular file)"; return false; }
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); err = std::string("cannot map ") + path; return false; }
            madvise(m, n, MADV_SEQUENTIAL);
            p = (const uint8_t *)m;
            mapped = true;
        }
        ::close(fd);
        return true;
    }
};

// Segment buffers are recycled: a fresh 8 MB malloc is an mmap whose 2 048 pages fault in one by one
// (and are unmapped again on free) -- a tenth of the per-byte cost of the whole reader.
struct BufCache {
    std::mutex mu;
    std::vector<std::pair<char *, size_t>> free_list;
    size_t held = 0;
    char *get(size_t want, size_t *got) {
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].second >= want && free_list[i].second <= 2 * want + (1u << 20)) {
                    char *p = free_list[i].first;
                    *got = free_list[i].second;
                    held -= free_list[i].second;
                    free_list[i] = free_list.back();
                    free_list.pop_back();
                    return p;
                }
        }
        *got = want;
        return (char *)malloc(want);
    }
    void put(char *p, size_t n) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (n >= (1u << 20) && held + n <= (512u << 20)) { free_list.push_back({p, n}); held += n; return; }
        }
        free(p);
    }
    ~BufCache() { for (auto &f : free_list) free(f.first); }
};
static BufCache g_bufs;

// CRC-32 of a stretch of one gzip member's output, to be combined in order by the consumer
struct CrcPiece { uint32_t crc; size_t len; bool member_end; uint32_t want_crc, want_size; };

struct Segment {
    char *buf = nullptr;              // HEAD + capacity + slack (malloc: not zero-filled)
    size_t buf_size = 0;
    size_t acct_bytes = 0;            // what the driver's read-ahead budget is charged (set before the segment is queued)
    char *text = nullptr;             // buf + HEAD
    size_t len = 0;                   // decoded bytes in text[0, len)
    int file_idx = 0;
    bool file_first = false, file_end = false;
    Latch latch;                      // tasks still writing / indexing this segment
    std::mutex err_mu;
    std::string err;
    // per slice: newline offsets (relative to text) and CRC pieces, filled by the tasks
    struct Slice { size_t b = 0, e = 0; std::vector<int32_t> nl; std::vector<CrcPiece> crc; };
    std::vector<std::unique_ptr<Slice>> slices;    // tasks hold Slice pointers: stable while the driver appends
    Slice *add_slice(size_t b, size_t e) { slices.emplace_back(new Slice()); slices.back()->b = b; slices.back()->e = e; return slices.back().get(); }
    // ---- set by the consumer's finalise step ----
    std::vector<int32_t> nl;          // all line ends of the logical text [lstart, rec_end), in order
    std::vector<uint32_t> first_line; // slow path only: first line of every record
    ptrdiff_t lstart = 0;             // logical start (negative: carried head of a record from the previous segment)
    ptrdiff_t rec_end = 0;            // end of the last complete record (== lstart when none completes here)
    int64_t n_recs = 0;

    ~Segment() { g_bufs.put(buf, buf_size); }
    Segment() {}
    Segment(const Segment &) = delete;
    Segment &operator=(const Segment &) = delete;
    bool alloc(size_t cap) {          // (re)allocates keeping the contents
        const size_t want = HEAD + cap + 320;
        if (buf && buf_size >= want) return true;
        size_t got = 0;
        char *nb = g_bufs.get(want, &got);
        if (!nb) return false;
        if (buf) { memcpy(nb, buf, buf_size); g_bufs.put(buf, buf_size); }
        buf = nb; buf_size = got; text = buf + HEAD;
        return true;
    }
    void fail(const std::string &m) { std::lock_guard<std::mutex> lk(err_mu); if (err.empty()) err = m; }
    // line i of the logical text
    inline const char *line_begin(int64_t i) const { return text + (i == 0 ? lstart : (ptrdiff_t)nl[(size_t)i - 1] + 1); }
    inline const char *line_end(int64_t i) const { return text + nl[(size_t)i]; }
    inline int64_t rec_line(int64_t r) const { return first_line.empty() ? 4 * r : (int64_t)first_line[(size_t)r]; }
};
typedef std::shared_ptr<Segment> SegP;

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
// newline offsets of [b, e), 32 bytes per step (memchr per ~80-byte line costs a call per line)
__attribute__((target("avx2"))) static void index_avx2(const char *t, size_t b, size_t e, std::vector<int32_t> &nl) {
    const __m256i k = _mm256_set1_epi8('\n');
    size_t pos = b;
    for (; pos + 32 <= e; pos += 32) {
        uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(t + pos)), k));
        while (m) { nl.push_back((int32_t)(pos + (size_t)__builtin_ctz(m))); m &= m - 1; }
    }
    for (; pos < e; pos++) if (t[pos] == '\n') nl.push_back((int32_t)pos);
}
#endif

static void index_slice(Segment &s, Segment::Slice &sl) {
    const char *t = s.text;
    sl.nl.reserve((sl.e - sl.b) / 40 + 16);
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { index_avx2(t, sl.b, sl.e, sl.nl); return; }
#endif
    size_t pos = sl.b;
    while (pos < sl.e) {
        const char *q = (const char *)memchr(t + pos, '\n', sl.e - pos);
        if (!q) break;
        sl.nl.push_back((int32_t)(q - t));
        pos = (size_t)(q - t) + 1;
    }
}

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// ---- one mate's stream of segments ---------------------------------------------------------------
struct Side {
    std::vector<std::string> paths;
    Pool *pool = nullptr;
    std::thread driver;
    std::mutex mu;
    std::condition_variable cv_data, cv_room;
    std::deque<SegP> ready;          // produced, not yet taken by the consumer (tasks may still be running)
    size_t ready_bytes = 0;
    bool done = false, quit = false;
    std::string driver_err;

    // consumer state
    std::deque<SegP> segs;           // finalised segments that still hold unconsumed records
    int64_t seg0_rec = 0;            // first unconsumed record of segs.front()
    int64_t avail = 0;               // unconsumed records in segs
    bool eof = false;
    SegP last_final;                 // the most recently finalised segment (source of the carried record head)
    uint32_t run_crc = 0;            // CRC-32 of the current gzip member so far
    size_t run_len = 0;
    int64_t recs_total = 0;          // records finalised so far
    std::vector<int64_t> file_first_rec;   // per file: index of its first record (filled as files start)
    std::string err;

    void push(const SegP &s) {
        std::unique_lock<std::mutex> lk(mu);
        ready.push_back(s);
        if (!s->acct_bytes) s->acct_bytes = s->buf_size ? s->buf_size : 1;
        ready_bytes += s->acct_bytes;
        cv_data.notify_all();
    }
    bool wait_room() {               // false when the reader is being closed
        std::unique_lock<std::mutex> lk(mu);
        cv_room.wait(lk, [&] { return quit || ready_bytes < AHEAD_BYTES; });
        return !quit;
    }
    void finish(const std::string &e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!e.empty() && driver_err.empty()) driver_err = e;
        done = true;
        cv_data.notify_all();
    }

    // ---- driver ----------------------------------------------------------------------------------
    void submit_index(const SegP &seg, size_t b, size_t e, bool with_crc_member, uint32_t want_crc, uint32_t want_size, bool member_end) {
        // one slice = [b, e): newline index (+ CRC piece of a streamed gzip member)
        Segment::Slice *sl = seg->add_slice(b, e);
        seg->latch.add();
        SegP keep = seg;
        pool->submit([keep, sl, with_crc_member, want_crc, want_size, member_end] {
            Segment *sp = keep.get();
            index_slice(*sp, *sl);
            if (with_crc_member) {
                CrcPiece c;
                c.crc = af_crc32((const uint8_t *)sp->text + sl->b, sl->e - sl->b);
                c.len = sl->e - sl->b; c.member_end = member_end; c.want_crc = want_crc; c.want_size = want_size;
                sl->crc.push_back(c);
            }
            sp->latch.done();
        });
    }

    // whole small file (any format) in one task: inflate / copy + index
    void small_file_task(const std::shared_ptr<MappedFile> &mf, int file_idx) {
        SegP seg = std::make_shared<Segment>();
        seg->file_idx = file_idx; seg->file_first = true; seg->file_end = true;
        Segment::Slice *sl0 = seg->add_slice(0, 0);
        seg->acct_bytes = mf->n * 5 + HEAD;              // the task allocates (and may grow) the buffer itself
        seg->latch.add();
        SegP keep0 = seg;
        pool->submit([keep0, sl0, mf] {
            Segment *sp = keep0.get();
            std::string e;
            const uint8_t *p = mf->p, *end = mf->p + mf->n;
            if (mf->n >= 2 && p[0] == 0x1f && p[1] == 0x8b) {
                size_t cap = mf->n * 5 + (64u << 10);
                sp->alloc(cap);
                size_t out_len = 0;
                static thread_local afz::Inflater inf;
                while (p < end) {
                    size_t hl; uint32_t bs;
                    if (!afz::gzip_header(p, end, &hl, &bs)) {
                        if (out_len) break;                                 // trailing garbage after a member: ignored, as gzip does
                        e = "not a gzip member"; break;
                    }
                    inf.reset(p + hl, end);
                    const size_t member_start = out_len;
                    int rc;
                    for (;;) {
                        uint8_t *o = (uint8_t *)sp->text + out_len;
                        rc = inf.run((uint8_t *)sp->text + member_start, &o, (uint8_t *)sp->text + cap - 258, (uint8_t *)sp->text + cap);
                        out_len = (size_t)(o - (uint8_t *)sp->text);
                        if (rc != afz::NEED_OUTPUT) break;
                        cap *= 2;                                           // grow; pointers into the old buffer are re-derived
                        if (!sp->alloc(cap)) { rc = afz::ERR_DATA; e = "out of memory"; break; }
                    }
                    if (rc < 0) { if (e.empty()) e = rc == afz::ERR_TRUNCATED ? "gzip stream is truncated" : "corrupt deflate data"; break; }
                    const uint8_t *t = inf.byte_pos();
                    if (end - t < 8) { e = "gzip stream is truncated (no trailer)"; break; }
                    const uint32_t crc = af_crc32((const uint8_t *)sp->text + member_start, out_len - member_start);
                    if (crc != rd32(t) || (uint32_t)(out_len - member_start) != rd32(t + 4)) { e = "gzip CRC or length check failed"; break; }
                    p = t + 8;
                }
                sp->len = out_len;
            } else {
                sp->alloc(mf->n + 1);
                if (mf->n) memcpy(sp->text, mf->p, mf->n);
                sp->len = mf->n;
            }
            if (!e.empty()) sp->fail(e);
            if (sp->len && sp->text[sp->len - 1] != '\n') sp->text[sp->len++] = '\n';    // last line without a terminator
            sl0->b = 0; sl0->e = sp->len;
            index_slice(*sp, *sl0);
            sp->latch.done();
        });
        push(seg);
    }

    // plain text: segments are slices of the file, copied and indexed by the pool
    bool plain_file(const std::shared_ptr<MappedFile> &mf, int file_idx) {
        size_t off = 0;
        bool first = true;
        do {
            if (!wait_room()) return false;
            const size_t n = std::min(SEG_TEXT, mf->n - off);
            SegP seg = std::make_shared<Segment>();
            seg->file_idx = file_idx; seg->file_first = first; seg->file_end = off + n == mf->n;
            seg->alloc(n + 1);
            seg->len = n;
            const bool add_nl = seg->file_end && n && mf->p[mf->n - 1] != '\n';
            if (add_nl) seg->len = n + 1;
            for (size_t b = 0; b < seg->len || b == 0; b += SLICE) {
                const size_t e = std::min(b + SLICE, seg->len);
                Segment::Slice *sl = seg->add_slice(b, e);
                seg->latch.add();
                SegP keep = seg;
                const uint8_t *src = mf->p + off;
                pool->submit([keep, mf, sl, src, n, add_nl] {
                    Segment *sp = keep.get();
                    const size_t ce = std::min(sl->e, n);
                    if (ce > sl->b) memcpy(sp->text + sl->b, src + sl->b, ce - sl->b);
                    if (add_nl && sl->e == n + 1) sp->text[n] = '\n';
                    index_slice(*sp, *sl);
                    sp->latch.done();
                });
                if (e >= seg->len) break;
            }
            push(seg);
            off += n;
            first = false;
        } while (off < mf->n);
        return true;
    }

    // BGZF: walk the block headers, group blocks into segments, inflate groups of blocks on the pool.
    // Returns the file offset it stopped at (== file size when everything was BGZF).
    bool bgzf_file(const std::shared_ptr<MappedFile> &mf, int file_idx, size_t &off, bool &first, std::string &err) {
        const uint8_t *base = mf->p, *end = mf->p + mf->n;
        struct Blk { size_t off, hdr; uint32_t size, isize; };
        while (off < mf->n) {
            // collect blocks for one segment
            std::vector<Blk> blks;
            size_t text = 0, o = off;
            while (o < mf->n && text < SEG_TEXT) {
                size_t hl; uint32_t bs;
                if (!afz::gzip_header(base + o, end, &hl, &bs) || bs == 0) break;   // not BGZF from here on
                if (bs < hl + 8 || o + bs > mf->n) { err = "BGZF block is truncated"; return false; }
                const uint32_t isize = rd32(base + o + bs - 4);
                if (isize > (1u << 16)) { err = "BGZF block larger than 64 KB"; return false; }
                blks.push_back({o, hl, bs, isize});
                text += isize;
                o += bs;
            }
            if (blks.empty()) return true;            // caller continues in stream mode at `off`
            if (!wait_room()) return false;
            SegP seg = std::make_shared<Segment>();
            seg->file_idx = file_idx; seg->file_first = first;
            seg->file_end = o == mf->n;
            seg->alloc(text + 1);
            seg->len = text;
            // groups of blocks of ~SLICE text each: one task inflates them in place and indexes what it wrote
            size_t gi = 0, tpos = 0;
            std::vector<std::pair<size_t, size_t>> groups;   // [first block, last block)
            while (gi < blks.size()) {
                size_t gj = gi, gt = 0;
                while (gj < blks.size() && (gt < SLICE || gj == gi)) { gt += blks[gj].isize; gj++; }
                groups.push_back({gi, gj});
                gi = gj;
            }
            auto shared_blks = std::make_shared<std::vector<Blk>>(std::move(blks));
            for (size_t g = 0; g < groups.size(); g++) {
                size_t gt = 0;
                for (size_t k = groups[g].first; k < groups[g].second; k++) gt += (*shared_blks)[k].isize;
                Segment::Slice *sl = seg->add_slice(tpos, tpos + gt);
                tpos += gt;
                seg->latch.add();
                SegP keep = seg;
                const size_t kb = groups[g].first, ke = groups[g].second;
                pool->submit([keep, mf, shared_blks, sl, kb, ke] {
                    static thread_local afz::Inflater inf;
                    Segment *sp = keep.get();
                    uint8_t *o = (uint8_t *)sp->text + sl->b;
                    for (size_t k = kb; k < ke; k++) {
                        const Blk &b = (*shared_blks)[k];
                        const uint8_t *bp = mf->p + b.off;
                        inf.reset(bp + b.hdr, bp + b.size - 8);
                        uint8_t *start = o;
                        const int rc = inf.run(start, &o, start + b.isize + 1, start + b.isize);
                        if (rc != afz::OK_DONE || (size_t)(o - start) != b.isize) { sp->fail("corrupt BGZF block"); o = start + b.isize; continue; }
                        if (b.isize && af_crc32(start, b.isize) != rd32(bp + b.size - 8)) sp->fail("BGZF block CRC check failed");
                    }
                    index_slice(*sp, *sl);
                    sp->latch.done();
                });
            }
            push(seg);                       // (a last line without a terminator is closed by the consumer)
            off = o;
            first = false;
        }
        return true;
    }

    // ONE large gzip member at `off`, decoded by several workers at once (af_inflate_par.h): the compressed bytes are cut
    // into chunks; waves of chunks are decoded symbolically on the pool, each from a block start its task found by
    // search; the driver then closes them in order -- a chunk must start where the one before stopped (else it is
    // decoded again from there), its markers are resolved against the 32 KB that precede it -- and cuts the text into
    // segments whose slices are resolved, indexed and CRC-ed on the pool.  Returns 1 when the member was handled (off
    // moves behind its trailer), 0 when the path does not apply or the start could not be trusted (nothing was
    // pushed: the serial path takes over at off), -1 on error / close.
    int member_parallel(const std::shared_ptr<MappedFile> &mf, int file_idx, size_t &off, bool &first, std::string &err) {
        if (getenv("AF_GZIP_SERIAL")) return 0;
        const uint8_t *p = mf->p + off, *end = mf->p + mf->n;
        size_t hl; uint32_t bs;
        const int workers = pool->size();
        const size_t env_chunk = par_chunk_env(), PAR_CHUNK = env_chunk ? env_chunk : PAR_CHUNK_DEFAULT, PAR_MIN = 4 * PAR_CHUNK;
        if ((size_t)(end - p) < PAR_MIN || workers < (env_chunk ? 2 : PAR_MIN_WORKERS) || !afz::gzip_header(p, end, &hl, &bs) || bs) return 0;
        const uint8_t *base = p + hl;
        // text that compresses more than 20:1 is not what this path is sized for (its cell buffers are 2 bytes per decoded
        // byte): ISIZE, the last four bytes of a one-member file, says so up front
        if ((uint64_t)rd32(end - 4) > (uint64_t)(end - base) * 20u) return 0;
        const uint64_t total_bits = (uint64_t)(end - base) * 8u;
        const size_t n_chunks = ((size_t)(end - base) + PAR_CHUNK - 1) / PAR_CHUNK;
        const size_t wave = (size_t)std::min(12, std::max(2, workers / 2));
        struct Chunk { afz::SymResult r; Latch done; };
        auto new_chunk = [] {
            return std::shared_ptr<Chunk>(new Chunk(), [](Chunk *c) { g_cells.put(std::move(c->r.sym)); delete c; });
        };
        uint64_t at = 0;                                  // bit position of the next block to decode (a known block start)
        bool member_done = false, pushed_any = false;
        auto window = std::make_shared<std::vector<uint8_t>>(afz::WINDOW, (uint8_t)0);
        // a rolling window of `wave` chunks is in flight on the pool: chunk c + wave is submitted when chunk c is taken
        std::deque<std::shared_ptr<Chunk>> inflight;
        size_t next_submit = 0;
        auto stop_of = [&](size_t c) { return c + 1 < n_chunks ? (uint64_t)(c + 1) * PAR_CHUNK * 8u : ~0ull; };
        auto submit_chunk = [&](size_t c) {
            std::shared_ptr<Chunk> ck = new_chunk();
            ck->r.sym = g_cells.get();
            ck->done.add();
            const uint64_t nominal = (uint64_t)c * PAR_CHUNK * 8u, stop = stop_of(c);
            pool->submit([ck, mf, base, end, nominal, stop, c] {      // (mf keeps the mapping alive if the driver has left)
                afz::decode_chunk(base, end, c == 0 ? 0 : ~0ull, nominal, stop, c != 0, ck->r);
                ck->done.done();
            });
            inflight.push_back(ck);
        };
        {
            for (size_t c = 0; c < n_chunks && !member_done; c++) {
                while (next_submit < n_chunks && next_submit < c + wave) submit_chunk(next_submit++);
                std::shared_ptr<Chunk> ck = inflight.front();
                inflight.pop_front();
                const uint64_t stop = stop_of(c);
                if (c > 0 && at >= stop) continue;        // the chunk before ran past this one entirely
                ck->done.wait();
                if (ck->r.status != afz::OK_DONE || ck->r.start_bit != at) {
                    // not where the stream really continues (no block start found, or a false one): decode it again from `at`
                    afz::decode_chunk(base, end, at, at, stop, c != 0, ck->r);
                    if (ck->r.status != afz::OK_DONE) {
                        if (!pushed_any) return 0;        // let the serial decoder say what is wrong with this stream
                        err = ck->r.status == afz::ERR_TRUNCATED ? "gzip stream is truncated" : "corrupt deflate data";
                        return -1;
                    }
                }
                const afz::SymResult &r = ck->r;
                member_done = r.stream_end;
                uint32_t want_crc = 0, want_size = 0;
                if (member_done) {
                    const uint8_t *t = base + (r.end_bit + 7) / 8;
                    if (end - t < 8) { if (!pushed_any) return 0; err = "gzip stream is truncated (no trailer)"; return -1; }
                    want_crc = rd32(t); want_size = rd32(t + 4);
                    off = (size_t)(t + 8 - mf->p);
                }
                // the text of this chunk, one segment per SEG_TEXT bytes; every slice resolves, indexes and CRCs itself
                const size_t n = r.sym.size();
                const bool file_ends = member_done && off >= mf->n;
                for (size_t b0 = 0; b0 < n || (b0 == 0 && member_done); b0 += SEG_TEXT) {
                    const size_t len = std::min(SEG_TEXT, n - b0);
                    if (!wait_room()) return -1;
                    SegP seg = std::make_shared<Segment>();
                    seg->file_idx = file_idx; seg->file_first = first; first = false;
                    if (!seg->alloc(len + 1)) { err = "out of memory"; return -1; }
                    seg->len = len;
                    const bool last_of_member = member_done && b0 + len >= n;
                    for (size_t b = 0; b < len || (b == 0 && last_of_member); b += SLICE) {
                        const size_t e = std::min(b + SLICE, len);
                        Segment::Slice *sl = seg->add_slice(b, e);
                        seg->latch.add();
                        SegP keep = seg;
                        const bool mend = last_of_member && e == len;
                        pool->submit([keep, sl, ck, window, b0, mend, want_crc, want_size] {
                            Segment *sp = keep.get();
                            afz::resolve_cells(ck->r.sym.data() + b0 + sl->b, sl->e - sl->b, window->data(), (uint8_t *)sp->text + sl->b);
                            index_slice(*sp, *sl);
                            CrcPiece cp;
                            cp.crc = af_crc32((const uint8_t *)sp->text + sl->b, sl->e - sl->b);
                            cp.len = sl->e - sl->b; cp.member_end = mend; cp.want_crc = want_crc; cp.want_size = want_size;
                            sl->crc.push_back(cp);
                            sp->latch.done();
                        });
                        if (e == len) break;
                    }
                    if (last_of_member && file_ends) {
                        // (a last line without a terminator is closed by the consumer through file_end; the byte is added
                        // once the slices are done -- the latch is waited for by the consumer before it reads the text)
                        seg->file_end = true;
                    }
                    push(seg);
                    pushed_any = true;
                    if (b0 + len >= n) break;
                }
                // the window in front of the next chunk: the last 32 KB of everything resolved so far
                auto nw = std::make_shared<std::vector<uint8_t>>(afz::WINDOW, (uint8_t)0);
                if (n >= afz::WINDOW) afz::resolve_cells(r.sym.data() + n - afz::WINDOW, afz::WINDOW, window->data(), nw->data());
                else {
                    memcpy(nw->data(), window->data() + n, afz::WINDOW - n);
                    afz::resolve_cells(r.sym.data(), n, window->data(), nw->data() + afz::WINDOW - n);
                }
                window = nw;
                at = r.end_bit;
            }
        }
        if (!member_done) { err = "gzip stream is truncated"; return -1; }
        (void)total_bits;
        return 1;
    }

    // a serial gzip stream (one or more plain members) from `off`: the driver inflates, the pool indexes + CRCs
    bool stream_file(const std::shared_ptr<MappedFile> &mf, int file_idx, size_t off, bool first, std::string &err) {
        const uint8_t *p = mf->p + off, *end = mf->p + mf->n;
        std::unique_ptr<afz::Inflater> inf(new afz::Inflater());
        SegP prev;
        SegP seg;
        size_t hist = 0;                      // bytes of this member's history in front of seg->text + seg->len .. (within HEAD)
        bool any_member = false;
        auto new_segment = [&]() -> bool {
            if (!wait_room()) return false;
            SegP s = std::make_shared<Segment>();
            s->file_idx = file_idx; s->file_first = first; first = false;
            s->alloc(SEG_TEXT);
            // the window: the HEAD bytes that end where the previous text ends (every segment owns HEAD bytes in
            // front of its text, so the source range is always inside the previous buffer)
            if (seg) memcpy(s->text - HEAD, seg->text + seg->len - HEAD, HEAD);
            prev = seg;
            seg = s;
            return true;
        };
        auto close_segment = [&](bool file_end) {
            seg->file_end = file_end;
            if (file_end && seg->len && seg->text[seg->len - 1] != '\n') seg->text[seg->len++] = '\n';
            push(seg);
        };
        if (!new_segment()) return false;
        size_t slice_b = 0;                   // start of the text not yet handed to an index task
        while (p < end) {
            size_t hl; uint32_t bs;
            if (!afz::gzip_header(p, end, &hl, &bs)) {
                if (any_member) break;        // trailing garbage after the last member: ignored, as gzip does
                err = "not a gzip file";
                return false;
            }
            any_member = true;
            inf->reset(p + hl, end);
            hist = 0;
            for (;;) {
                uint8_t *text = (uint8_t *)seg->text;
                uint8_t *o = text + seg->len;
                const size_t room_hist = std::min(hist, HEAD + seg->len);
                const int rc = inf->run(o - room_hist, &o, text + SEG_TEXT - 258, text + SEG_TEXT);
                const size_t got = (size_t)(o - (text + seg->len));
                seg->len += got;
                hist += got;
                if (rc < 0) { err = rc == afz::ERR_TRUNCATED ? "gzip stream is truncated" : "corrupt deflate data"; close_segment(true); return false; }
                if (rc == afz::OK_DONE) {
                    const uint8_t *t = inf->byte_pos();
                    if (end - t < 8) { err = "gzip stream is truncated (no trailer)"; close_segment(true); return false; }
                    // slices of this member up to here, the last one closes the member
                    for (size_t b = slice_b; b < seg->len || b == slice_b; b += SLICE) {
                        const size_t e = std::min(b + SLICE, seg->len);
                        submit_index(seg, b, e, true, rd32(t), rd32(t + 4), e == seg->len);
                        if (e == seg->len) break;
                    }
                    slice_b = seg->len;
                    p = t + 8;
                    break;
                }
                // NEED_OUTPUT: the segment is full
                for (size_t b = slice_b; b < seg->len; b += SLICE) submit_index(seg, b, std::min(b + SLICE, seg->len), true, 0, 0, false);
                close_segment(false);
                if (!new_segment()) return false;
                slice_b = 0;
            }
        }
        close_segment(true);
        return true;
    }

    void drive() {
        std::string err;
        for (size_t fi = 0; fi < paths.size(); fi++) {
            auto mf = std::make_shared<MappedFile>();
            if (!mf->open(paths[fi].c_str(), err)) break;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (quit) return;
            }
            const bool gz = mf->n >= 2 && mf->p[0] == 0x1f && mf->p[1] == 0x8b;
            if (mf->n <= SMALL_FILE) {
                if (!wait_room()) return;
                small_file_task(mf, (int)fi);
                continue;
            }
            if (!gz) { if (!plain_file(mf, (int)fi)) return; continue; }
            size_t off = 0;
            bool first = true;
            if (!bgzf_file(mf, (int)fi, off, first, err)) { if (err.empty()) return; break; }
            bool failed = false;
            while (off < mf->n) {                         // large plain members in parallel, whatever else serially
                const int pr = member_parallel(mf, (int)fi, off, first, err);
                if (pr < 0) { failed = true; break; }
                if (pr == 0) { if (!stream_file(mf, (int)fi, off, first, err)) failed = true; break; }
            }
            if (failed) { if (err.empty()) return; break; }
        }
        finish(err);
    }
    void start() { driver = std::thread([this] { drive(); }); }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv_room.notify_all();
        if (driver.joinable()) driver.join();
    }

    // ---- consumer --------------------------------------------------------------------------------
    // slow path: group lines into records the way the serial parser did (blank lines between records skipped)
    static void group_lines_tolerant(Segment &s, int64_t n_lines, int64_t &lines_used) {
        s.first_line.clear();
        int64_t i = 0;
        lines_used = 0;
        while (i < n_lines) {
            const char *b = s.line_begin(i), *e = s.line_end(i);
            if (e > b && e[-1] == '\r') e--;
            if (e == b) { i++; lines_used = i; continue; }        // stray blank line between records
            if (i + 4 > n_lines) break;
            s.first_line.push_back((uint32_t)i);
            i += 4;
            lines_used = i;
        }
    }

    // Takes the next produced segment, waits for its tasks, joins it to the previous one (carried record
    // head), lists its lines and counts its records.  false at the end of the stream or on error (err set).
    bool finalize_next() {
        SegP s;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_data.wait(lk, [&] { return done || !ready.empty(); });
            if (ready.empty()) {
                if (!driver_err.empty()) err = driver_err;
                else if (last_final && last_final->rec_end < (ptrdiff_t)last_final->len) err = "truncated FASTQ record";
                eof = true;
                return false;
            }
            s = ready.front();
            ready.pop_front();
            ready_bytes -= s->acct_bytes;
        }
        cv_room.notify_all();
        s->latch.wait();
        if (!s->err.empty()) { err = s->err; eof = true; return false; }
        // member CRCs of streamed gzip input, in order
        for (auto &sl : s->slices)
            for (auto &c : sl->crc) {
                run_crc = run_len ? (uint32_t)crc32_combine(run_crc, c.crc, (z_off_t)c.len) : c.crc;
                if (run_len == 0 && c.len == 0) run_crc = (uint32_t)crc32(0L, Z_NULL, 0);
                run_len += c.len;
                if (c.member_end) {
                    if (run_crc != c.want_crc || (uint32_t)run_len != c.want_size) { err = "gzip CRC or length check failed"; eof = true; return false; }
                    run_crc = 0; run_len = 0;
                }
            }
        if (s->file_end && s->len && s->text[s->len - 1] != '\n') {       // BGZF / plain: last line without a terminator
            s->text[s->len] = '\n';
            if (!s->slices.empty()) s->slices.back()->nl.push_back((int32_t)s->len);
            s->len++;
        }
        // the head of a record that straddles the cut
        size_t carry = 0;
        if (last_final && last_final->rec_end < (ptrdiff_t)last_final->len) {
            carry = (size_t)((ptrdiff_t)last_final->len - last_final->rec_end);
            if (s->file_first) { err = "truncated FASTQ record"; eof = true; return false; }
            if (carry > HEAD) { err = "a FASTQ record longer than 64 KB"; eof = true; return false; }
            // (a streamed gzip member's driver has put the same bytes there already -- its deflate window --
            // and may be reading them for the next segment: compare first, write only when they differ)
            if (memcmp(s->text - carry, last_final->text + last_final->rec_end, carry) != 0)
                memmove(s->text - carry, last_final->text + last_final->rec_end, carry);
        }
        s->lstart = -(ptrdiff_t)carry;
        if (s->file_first && (int)file_first_rec.size() <= s->file_idx) file_first_rec.resize((size_t)s->file_idx + 1, recs_total);
        // all line ends, in order: those inside the carried head first
        size_t total = 0;
        for (auto &sl : s->slices) total += sl->nl.size();
        s->nl.clear();
        s->nl.reserve(total + 4);
        for (ptrdiff_t i = s->lstart; i < 0; i++) if (s->text[i] == '\n') s->nl.push_back((int32_t)i);
        for (auto &sl : s->slices) {
            s->nl.insert(s->nl.end(), sl->nl.begin(), sl->nl.end());
            std::vector<int32_t>().swap(sl->nl);
        }
        const int64_t n_lines = (int64_t)s->nl.size();
        // a blank line anywhere (also "\r\n" alone) sends the segment to the tolerant grouping
        bool blank = false;
        {
            ptrdiff_t prev_end = s->lstart - 1;
            for (int64_t i = 0; i < n_lines; i++) {
                const ptrdiff_t e = s->nl[(size_t)i];
                const ptrdiff_t l = e - prev_end - 1;
                if (l == 0 || (l == 1 && s->text[e - 1] == '\r')) { blank = true; break; }
                prev_end = e;
            }
        }
        int64_t lines_used;
        if (blank) {
            group_lines_tolerant(*s, n_lines, lines_used);
            s->n_recs = (int64_t)s->first_line.size();
            if (s->n_recs == 0) s->first_line.push_back(0);      // keeps rec_line() on the table path; never read
        } else {
            s->n_recs = n_lines / 4;
            lines_used = 4 * s->n_recs;
        }
        s->rec_end = lines_used ? (ptrdiff_t)s->nl[(size_t)lines_used - 1] + 1 : s->lstart;
        if (s->file_end && s->rec_end < (ptrdiff_t)s->len) {
            for (ptrdiff_t i = s->rec_end; i < (ptrdiff_t)s->len; i++)           // only blank lines may remain
                if (s->text[i] != '\n' && s->text[i] != '\r') { err = "truncated FASTQ record"; eof = true; return false; }
            s->rec_end = (ptrdiff_t)s->len;
        }
        recs_total += s->n_recs;
        last_final = s;
        if (s->n_recs) { segs.push_back(s); avail += s->n_recs; }
        return true;
    }
};

}  // namespace

struct af_fastq {
    std::unique_ptr<Pool> pool;
    Side side[2];
    int n_files = 0;
    // the current batch: per side, spans of records inside segments
    struct Span { SegP seg; int64_t r0, r1, p0; };     // records [r0, r1) of seg are pairs [p0, p0 + r1 - r0)
    std::vector<Span> spans[2];
    int64_t n_cur = 0;
    int64_t pairs_done = 0;                              // pairs handed out by earlier batches
    std::string sticky_err;
};

namespace {

struct RecView { const char *name; int32_t name_len; const char *seq, *qual; int32_t len; };

// parse record r of a segment; returns nullptr or an error text
static const char *view_record(const Segment &s, int64_t r, RecView &v) {
    const int64_t l0 = s.rec_line(r);
    const char *b0 = s.line_begin(l0), *e0 = s.line_end(l0);
    const char *b1 = e0 + 1, *e1 = s.line_end(l0 + 1);
    const char *b2 = e1 + 1, *e2 = s.line_end(l0 + 2);
    const char *b3 = e2 + 1, *e3 = s.line_end(l0 + 3);
    if (e0 > b0 && e0[-1] == '\r') e0--;
    if (e1 > b1 && e1[-1] == '\r') e1--;
    if (e2 > b2 && e2[-1] == '\r') e2--;
    if (e3 > b3 && e3[-1] == '\r') e3--;
    if (e0 == b0 || *b0 != '@') return "FASTQ record does not start with '@'";
    if (e2 == b2 || *b2 != '+') return "FASTQ record lacks its '+' line";
    if (e3 - b3 != e1 - b1) return "FASTQ quality length differs from sequence length";
    // name: up to the first blank; a trailing /1 or /2 is dropped, as bwa does
    const char *n = b0 + 1, *ne = n;
    while (ne < e0 && *ne != ' ' && *ne != '\t') ne++;
    size_t nl = (size_t)(ne - n);
    if (nl >= 2 && ne[-2] == '/' && (ne[-1] == '1' || ne[-1] == '2')) nl -= 2;
    v.name = n; v.name_len = (int32_t)nl; v.seq = b1; v.qual = b3; v.len = (int32_t)(e1 - b1);
    return nullptr;
}

}  // namespace

static int fastq_open_impl(const char *const *paths1, const char *const *paths2, int n_files, int n_threads, af_fastq_t **out) {
    if (!paths1 || !paths2 || !out || n_files <= 0) { af_set_error("af_fastq_open: null argument"); return AF_ERR_ARG; }
    for (int i = 0; i < n_files; i++) {
        if (!paths1[i] || !paths2[i]) { af_set_error("af_fastq_open: null path"); return AF_ERR_ARG; }
        for (const char *p : {paths1[i], paths2[i]}) {
            if (access(p, R_OK) != 0) { af_set_error("af_fastq_open: cannot open %s", p); return AF_ERR_IO; }
        }
    }
    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 4;
        if (n_threads > 32) n_threads = 32;
    }
    af_fastq *fq = new af_fastq();
    fq->n_files = n_files;
    fq->pool.reset(new Pool(n_threads));
    for (int m = 0; m < 2; m++) {
        fq->side[m].pool = fq->pool.get();
        for (int i = 0; i < n_files; i++) fq->side[m].paths.push_back(m == 0 ? paths1[i] : paths2[i]);
    }
    for (int m = 0; m < 2; m++) fq->side[m].start();
    *out = fq;
    return AF_OK;
}

extern "C" int af_fastq_open(const char *path1, const char *path2, af_fastq_t **out) {
    if (!path1 || !path2) { af_set_error("af_fastq_open: null argument"); return AF_ERR_ARG; }
    return fastq_open_impl(&path1, &path2, 1, 0, out);
}

extern "C" int af_fastq_open_threads(const char *path1, const char *path2, int32_t n_threads, af_fastq_t **out) {
    if (!path1 || !path2) { af_set_error("af_fastq_open: null argument"); return AF_ERR_ARG; }
    return fastq_open_impl(&path1, &path2, 1, n_threads, out);
}

extern "C" int af_fastq_open_multi(const char *const *paths1, const char *const *paths2, int32_t n_files, int32_t n_threads,
                                   af_fastq_t **out) {
    return fastq_open_impl(paths1, paths2, n_files, n_threads, out);
}

extern "C" int af_fastq_threads(const af_fastq_t *fq) { return fq ? fq->pool->size() : 0; }

// Longest read among the first n_records records of a FASTQ / FASTQ.gz file (the reader needs the
// packed width of a read before it starts).
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
    fq->pool.reset();        // joins the workers; queued tasks still run (they hold their segments alive)
    delete fq;
    g_cells.trim();
}

// Gathers up to max_pairs records per mate; n = pairs available on both.  Errors are sticky.
static int gather(af_fastq *fq, int64_t max_pairs, int64_t *n_out) {
    if (!fq->sticky_err.empty()) { af_set_error("%s", fq->sticky_err.c_str()); return AF_ERR_IO; }
    for (int m = 0; m < 2; m++) {
        Side &sd = fq->side[m];
        while (sd.avail < max_pairs && !sd.eof) sd.finalize_next();
        if (!sd.err.empty()) {
            char buf[400];
            snprintf(buf, sizeof(buf), "af_fastq_next: file %d: %s", m + 1, sd.err.c_str());
            fq->sticky_err = buf;
            af_set_error("%s", buf);
            return AF_ERR_IO;
        }
    }
    int64_t n = std::min(std::min(fq->side[0].avail, fq->side[1].avail), max_pairs);
    // one mate's stream ended while the other still has records, or two files of one pair (cell) differ in size
    bool out_of_step = false;
    for (int m = 0; m < 2; m++) if (fq->side[m].eof && fq->side[m].avail == n && n < max_pairs && fq->side[m ^ 1].avail > n) out_of_step = true;
    const std::vector<int64_t> &fa = fq->side[0].file_first_rec, &fb = fq->side[1].file_first_rec;
    for (size_t i = 0; i < std::min(fa.size(), fb.size()); i++) if (fa[i] != fb[i]) out_of_step = true;
    if (out_of_step) {
        char buf[300];
        snprintf(buf, sizeof(buf), "af_fastq_next: the two FASTQ files are out of step (%lld vs %lld records so far)",
                 (long long)(fq->side[0].recs_total), (long long)(fq->side[1].recs_total));
        fq->sticky_err = buf;
        af_set_error("%s", buf);
        return AF_ERR_IO;
    }
    *n_out = n;
    return AF_OK;
}

// moves the batch's records out of the sides' queues into span lists
static void take_spans(af_fastq *fq, int64_t n) {
    for (int m = 0; m < 2; m++) {
        Side &sd = fq->side[m];
        fq->spans[m].clear();
        int64_t need = n, p0 = 0;
        while (need > 0) {
            SegP s = sd.segs.front();
            const int64_t r0 = sd.seg0_rec, take = std::min(need, s->n_recs - r0);
            fq->spans[m].push_back({s, r0, r0 + take, p0});
            p0 += take; need -= take;
            if (r0 + take == s->n_recs) { sd.segs.pop_front(); sd.seg0_rec = 0; }
            else sd.seg0_rec = r0 + take;
        }
        sd.avail -= n;
    }
    fq->n_cur = n;
}

extern "C" int af_fastq_next(af_fastq_t *fq, int64_t max_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                             uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                             int64_t *n_nreads_out, int32_t *uniform_len_out, int64_t *n_pairs_out) {
    if (!fq || !n_pairs_out || max_pairs <= 0) { af_set_error("af_fastq_next: bad argument"); return AF_ERR_ARG; }
    int64_t n = 0;
    int rc = gather(fq, max_pairs, &n);
    if (rc) return rc;
    fq->pairs_done += fq->n_cur;
    take_spans(fq, n);
    *n_pairs_out = n;
    if (n == 0) { if (n_nreads_out) *n_nreads_out = 0; if (uniform_len_out) *uniform_len_out = 0; return AF_OK; }
    if (!packed_out) { af_set_error("af_fastq_next: packed_out is null"); return AF_ERR_ARG; }
    af_layout_t lay;
    rc = af_layout(max_read_len, n, &lay);
    if (rc) return rc;
    // parse + pack, fanned out over the pool: tile-aligned pair ranges per mate
    const int T = fq->pool->size();
    int64_t per = (n + 2 * T - 1) / (2 * T);
    per = std::max<int64_t>(4096, (per + 31) / 32 * 32);
    struct Job { int m; int64_t p0, p1; PackSide ps; std::string err; };
    std::vector<Job> jobs;
    for (int m = 0; m < 2; m++)
        for (int64_t p0 = 0; p0 < n; p0 += per) { jobs.emplace_back(); jobs.back().m = m; jobs.back().p0 = p0; jobs.back().p1 = std::min(n, p0 + per); }
    Latch latch;
    latch.add((int)jobs.size());
    for (size_t j = 0; j < jobs.size(); j++) {
        Job *job = &jobs[j];
        fq->pool->submit([fq, job, n, max_read_len, pad_byte, packed_out, lens_out, &latch] {
            const std::vector<af_fastq::Span> &sp = fq->spans[job->m];
            std::vector<SeqRef> refs((size_t)(job->p1 - job->p0));
            size_t k = 0;
            while (k + 1 < sp.size() && sp[k + 1].p0 <= job->p0) k++;
            for (int64_t p = job->p0; p < job->p1; p++) {
                while (p >= sp[k].p0 + (sp[k].r1 - sp[k].r0)) k++;
                RecView v;
                const char *e = view_record(*sp[k].seg, sp[k].r0 + (p - sp[k].p0), v);
                if (e) { job->err = e; break; }
                refs[(size_t)(p - job->p0)] = {v.seq, v.len};
            }
            if (job->err.empty()) af_pack_range(refs.data(), job->m, job->p0, job->p1, n, max_read_len, pad_byte, packed_out, lens_out, job->ps);
            latch.done();
        });
    }
    latch.wait();
    PackSide ps[2];
    ps[0].ulen = ps[1].ulen = -1;
    for (Job &j : jobs) {
        if (!j.err.empty()) {
            char buf[300];
            snprintf(buf, sizeof(buf), "af_fastq_next: file %d: %s", j.m + 1, j.err.c_str());
            fq->sticky_err = buf;
            af_set_error("%s", buf);
            return AF_ERR_IO;
        }
        PackSide &d = ps[j.m];
        if (j.ps.rc && !d.rc) { d.rc = j.ps.rc; d.err = j.ps.err; }
        d.nids.insert(d.nids.end(), j.ps.nids.begin(), j.ps.nids.end());
        d.nmask.insert(d.nmask.end(), j.ps.nmask.begin(), j.ps.nmask.end());
        if (j.ps.ulen != -1) { if (d.ulen == -1) d.ulen = j.ps.ulen; else if (d.ulen != j.ps.ulen) d.ulen = -2; }
    }
    return af_pack_finish(ps[0], ps[1], nread_ids_out, nmask_out, ncap, n_nreads_out, uniform_len_out);
}

// Advance over up to max_pairs pairs without packing them (a rank of a multi-GPU job skipping batches
// that belong to other ranks).  Record text of the skipped batch is still available.
extern "C" int af_fastq_skip(af_fastq_t *fq, int64_t max_pairs, int64_t *n_pairs_out) {
    if (!fq || !n_pairs_out || max_pairs <= 0) { af_set_error("af_fastq_skip: bad argument"); return AF_ERR_ARG; }
    int64_t n = 0;
    int rc = gather(fq, max_pairs, &n);
    if (rc) return rc;
    fq->pairs_done += fq->n_cur;
    take_spans(fq, n);
    *n_pairs_out = n;
    return AF_OK;
}

extern "C" int af_fastq_record(const af_fastq_t *fq, int64_t read_id, const char **name, int32_t *name_len,
                               const char **seq, const char **qual, int32_t *len) {
    if (!fq || read_id < 0 || (read_id >> 1) >= fq->n_cur) { af_set_error("af_fastq_record: read_id out of range"); return AF_ERR_ARG; }
    const std::vector<af_fastq::Span> &sp = fq->spans[read_id & 1];
    const int64_t p = read_id >> 1;
    size_t lo = 0, hi = sp.size();
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (sp[mid].p0 <= p) lo = mid; else hi = mid; }
    RecView v;
    const char *e = view_record(*sp[lo].seg, sp[lo].r0 + (p - sp[lo].p0), v);
    if (e) { af_set_error("af_fastq_record: %s", e); return AF_ERR_IO; }
    if (name) *name = v.name;
    if (name_len) *name_len = v.name_len;
    if (seq) *seq = v.seq;
    if (qual) *qual = v.qual;
    if (len) *len = v.len;
    return AF_OK;
}

// Many records of the current batch at once: for read_ids[i], the offsets of its name / bases / qualities
// inside the caller's byte buffer `text` (filled back to back: name, bases, qualities) -- one call
// instead of three ctypes round trips per anchored read.  offs: 4 int64 per read (name, seq, qual, end).
extern "C" int af_fastq_records(const af_fastq_t *fq, const int64_t *read_ids, int64_t n, char *text, int64_t text_cap,
                                int64_t *offs, int64_t *text_used) {
    if (!fq || (n && (!read_ids || !offs)) || !text_used) { af_set_error("af_fastq_records: bad argument"); return AF_ERR_ARG; }
    int64_t at = 0;
    for (int64_t i = 0; i < n; i++) {
        const char *nm, *sq, *ql;
        int32_t nl, ln;
        int rc = af_fastq_record(fq, read_ids[i], &nm, &nl, &sq, &ql, &ln);
        if (rc) return rc;
        const int64_t need = (int64_t)nl + 2 * (int64_t)ln;
        if (text && at + need <= text_cap) {
            memcpy(text + at, nm, (size_t)nl);
            memcpy(text + at + nl, sq, (size_t)ln);
            memcpy(text + at + nl + ln, ql, (size_t)ln);
        }
        offs[4 * i] = at; offs[4 * i + 1] = at + nl; offs[4 * i + 2] = at + nl + ln; offs[4 * i + 3] = at + need;
        at += need;
    }
    *text_used = at;
    if (text && at > text_cap) { af_set_error("af_fastq_records: text buffer holds %lld bytes, %lld needed", (long long)text_cap, (long long)at); return AF_ERR_CAPACITY; }
    return AF_OK;
}

// Pair index (over the whole run) at which each input file starts, for the files started so far
// (single-cell runs: file i = cell i).  Returns the number of entries written.
extern "C" int af_fastq_file_starts(const af_fastq_t *fq, int64_t *first_pair_out, int32_t cap) {
    if (!fq || !first_pair_out) return 0;
    const std::vector<int64_t> &a = fq->side[0].file_first_rec, &b = fq->side[1].file_first_rec;
    const size_t n = std::min(std::min(a.size(), b.size()), (size_t)std::max(cap, 0));
    for (size_t i = 0; i < n; i++) first_pair_out[i] = a[i];
    return (int)n;
}

// pairs handed out before the current batch (global index of the current batch's first pair)
extern "C" int64_t af_fastq_batch_first_pair(const af_fastq_t *fq) { return fq ? fq->pairs_done : 0; }

// test hook: the reader's CRC-32 (carry-less multiplication where the CPU has it) of a host buffer
extern "C" uint32_t af_debug_crc32(const void *buf, int64_t len) { return buf && len > 0 ? af_crc32((const uint8_t *)buf, (size_t)len) : af_crc32((const uint8_t *)"", 0); }

// test hook: one gzip member decoded as the reader's parallel path decodes it -- the deflate payload cut into n_chunks
// pieces, each decoded symbolically by its own thread from a block start it found itself, then chained and resolved.
// *n_redone = chunks that had to be decoded again because they did not start where the chunk before them stopped.
extern "C" int af_debug_gunzip_chunks(const void *gz, int64_t n, int32_t n_chunks, void *out, int64_t cap, int64_t *out_len, int32_t *n_redone) {
    const uint8_t *p = (const uint8_t *)gz, *end = p + n;
    size_t hl; uint32_t bs;
    if (!gz || !out_len || n_chunks < 1 || !afz::gzip_header(p, end, &hl, &bs)) { af_set_error("af_debug_gunzip_chunks: not a gzip member"); return AF_ERR_ARG; }
    const uint8_t *base = p + hl;
    const uint64_t total_bits = (uint64_t)(end - base) * 8u;
    std::vector<afz::SymResult> res((size_t)n_chunks);
    std::vector<std::thread> th;
    for (int k = 0; k < n_chunks; k++)
        th.emplace_back([&, k] {
            const uint64_t nominal = total_bits / (uint64_t)n_chunks * (uint64_t)k / 8u * 8u, stop = k + 1 < n_chunks ? total_bits / (uint64_t)n_chunks * (uint64_t)(k + 1) / 8u * 8u : ~0ull;
            afz::decode_chunk(base, end, k == 0 ? 0 : ~0ull, nominal, stop, k != 0, res[(size_t)k]);
        });
    for (auto &t : th) t.join();
    // chain: chunk k must start where the last good chunk stopped
    int redone = 0;
    std::vector<uint8_t> window(afz::WINDOW, 0);
    uint8_t *o = (uint8_t *)out;
    int64_t len = 0;
    uint64_t at = 0;
    bool done = false;
    for (int k = 0; k < n_chunks && !done; k++) {
        afz::SymResult &r = res[(size_t)k];
        const uint64_t stop = k + 1 < n_chunks ? total_bits / (uint64_t)n_chunks * (uint64_t)(k + 1) / 8u * 8u : ~0ull;
        if (k > 0 && at >= stop) continue;                    // the chunk before ran past this one entirely
        if (r.status != afz::OK_DONE || r.start_bit != at) {
            afz::decode_chunk(base, end, at, at, stop, k != 0, r);
            redone += k != 0;
            if (r.status != afz::OK_DONE) { af_set_error("af_debug_gunzip_chunks: corrupt deflate data in chunk %d", k); return AF_ERR_IO; }
        }
        const int64_t m = (int64_t)r.sym.size();
        if (len + m > cap) { af_set_error("af_debug_gunzip_chunks: output buffer too small"); return AF_ERR_CAPACITY; }
        afz::resolve_cells(r.sym.data(), (size_t)m, window.data(), o + len);
        len += m;
        // the window of the next chunk: the last 32 KB of everything resolved so far
        if (len >= (int64_t)afz::WINDOW) memcpy(window.data(), o + len - afz::WINDOW, afz::WINDOW);
        else { memset(window.data(), 0, afz::WINDOW); memcpy(window.data() + afz::WINDOW - len, o, (size_t)len); }
        at = r.end_bit;
        done = r.stream_end;
    }
    if (!done) { af_set_error("af_debug_gunzip_chunks: the stream does not end"); return AF_ERR_IO; }
    const uint8_t *t = base + (at + 7) / 8;
    if (end - t < 8 || af_crc32(o, (size_t)len) != rd32(t) || (uint32_t)len != rd32(t + 4)) { af_set_error("af_debug_gunzip_chunks: CRC or length check failed"); return AF_ERR_IO; }
    *out_len = len;
    if (n_redone) *n_redone = redone;
    return AF_OK;
}
