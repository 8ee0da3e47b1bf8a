/*
 * anchored_fusion.h -- C ABI of the B200-native read-anchoring path (libafb200.so).
 *
 * The reference (ShenLab-Genomics/Anchored-Fusion) has NO FFI for this path: the pass is
 * three shell-outs in the driver scripts,
 *     bwa index <anchor.fa>                                  Anchored_Fusion.py:167-172
 *     bwa mem -M -t T <anchor.fa> fq1 fq2 | samtools ...     Anchored_Fusion.py:181-182
 *     samtools view -f 8 -F 260 / -f 4 -F 264 / -F 772       Anchored_Fusion.py:186,187,194
 * (same lines in Anchored_Fusion_singlecell.py:185-231), glued by files on disk.  The entry
 * points below are what a binding for that path would bind; each cites the reference step it
 * replaces.  INTEGRATION.md shows the ctypes stub and the file-level drop-in.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative af_status code, and af_last_error() gives the message (thread-local).  No
 * exceptions cross the ABI.  There is NO CPU fallback: device entry points fail with
 * AF_ERR_CUDA when no sm_100 device is usable.  Device entry points are stream-ordered and
 * re-entrant per (device, stream); `stream` is a cudaStream_t passed as void*.  The caller
 * owns every device buffer (e.g. torch tensors' data_ptr()); the library owns only the
 * objects it returns through af_*_build / af_*_create and frees them in af_*_free.
 */
#ifndef ANCHORED_FUSION_H
#define ANCHORED_FUSION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AF_ABI_VERSION 2 /* 2: AF_NMASK_WORDS 8 -> 16, AF_MAX_READ_LEN 256 -> 512, W rounded to a multiple of 4 beyond 16 */
#define AF_MAX_READ_LEN 512 /* bases; up to 32 packed words per read (reads beyond 256 bases take the long-read kernels) */
#define AF_TILE_PAIRS 32    /* pairs per packed tile == lanes per warp */
#define AF_NMASK_WORDS 16   /* 512-bit N mask per read that holds an N */
#define AF_GENOME_MAX_READ_LEN 256 /* af_genome_align (the contiguity filter's genome pass) takes reads up to this length */

enum af_status {
    AF_OK = 0,
    AF_ERR_ARG = -1,      /* bad argument */
    AF_ERR_CUDA = -2,     /* CUDA runtime error / no usable device */
    AF_ERR_CAPACITY = -3, /* caller-provided buffer too small (nothing is dropped silently) */
    AF_ERR_IO = -4,       /* file / zlib error */
    AF_ERR_NOMEM = -5
};

/* Alignment parameters == bwa-mem defaults the reference relies on implicitly
 * (Anchored_Fusion.py:182 passes only "-M -t"): -k 19 -A 1 -B 4 -L 5,5 -T 30 -d 100. */
typedef struct {
    int32_t k, A, B, clip5, clip3, T, X;
} af_params_t;

/* One anchored read == one primary mapped SAM record of `bwa mem -M | samtools view -F 772`
 * (Anchored_Fusion.py:194): CIGAR is clip_l S, m_len M, clip_r S; pos is SAM POS (1-based,
 * anchor-forward); strand 1 == FLAG 0x10 (SEQ reverse-complemented).  read_id = pair*2+mate.
 * These are the fields deal_cigar (functions.py:656) / contact_reads (functions.py:917-930)
 * turn into type SM/MS and the split point. */
typedef struct {
    uint32_t read_id;
    int32_t pos;
    uint16_t clip_l;
    uint16_t m_len;
    uint16_t clip_r;
    uint16_t score_strand; /* score*2 + strand */
} af_hit_t;

typedef struct af_index af_index_t;         /* host-side anchor index */
typedef struct af_dev_index af_dev_index_t; /* its copy in one GPU's HBM */
typedef struct af_fastq af_fastq_t;         /* paired FASTQ(.gz) reader */
typedef struct af_pipeline af_pipeline_t;   /* host->device streaming executor */

typedef struct {
    int32_t anchor_len;   /* G */
    int32_t k;            /* semantic seed length (19) */
    int32_t kp;           /* sampled k-mer length k' */
    int32_t stride;       /* read sampling stride s, k' + s - 1 <= k */
    int32_t n_keys;       /* distinct k'-mers over both strands */
    int32_t n_entries;    /* (k'-mer, strand, position) occurrences */
    int32_t n_buckets;    /* shared-memory filter buckets (4 B each) */
    int32_t n_overflow;   /* buckets marked always-hit */
    int32_t table_slots;  /* exact table slots (8 B each) */
    uint32_t filter_mul;  /* multiplier chosen for the filter hash */
    int32_t pad_byte;     /* 4-base pad pattern absent from the anchor's k'-mers */
} af_index_info_t;

/* Geometry of a packed batch.  Reads are 2 bit/base (A,C,G,T = 0..3, base i of a read in
 * bits [2i,2i+2) of its word i/16); a pair is W words of mate 1 then W words of mate 2,
 * zero-extended to Q 16-byte quads; tiles of 32 pairs are stored quad-interleaved:
 * quad q of pair (tile*32 + lane) lives at ((tile*Q + q)*32 + lane)*16 bytes, so a warp's
 * 128-bit loads are 512 contiguous bytes. */
typedef struct {
    int32_t max_read_len;
    int32_t words_per_read; /* W = ceil(max_read_len/16), rounded up to a multiple of 4 beyond 16 */
    int32_t quads_per_pair; /* Q = ceil(2W/4) */
    int32_t reserved;
    int64_t n_pairs;
    int64_t n_tiles;      /* ceil(n_pairs/32) */
    int64_t packed_bytes; /* n_tiles * Q * 512 */
} af_layout_t;

/* A batch of read pairs.  All pointers are DEVICE pointers for af_anchor_batch and friends,
 * HOST (ideally pinned) pointers for af_pipeline_run.  Replaces the FASTQ ingest inside
 * `bwa mem` (Anchored_Fusion.py:182). */
typedef struct {
    const void *packed;        /* packed_bytes of af_layout(max_read_len, n_pairs) */
    int64_t n_pairs;
    int32_t max_read_len;
    int32_t uniform_len;       /* > 0: every read has this length and lens is ignored */
    const uint16_t *lens;      /* [2*n_pairs] by read_id, or NULL when uniform_len > 0 */
    const uint32_t *nread_ids; /* sorted read_ids of reads holding N (packed as pad there) */
    const uint32_t *nmask;     /* [n_nreads][AF_NMASK_WORDS], bit i set = base i is N */
    int64_t n_nreads;
} af_batch_t;

/* counts written by af_anchor_batch (device memory, AF_N_COUNTS x uint32) */
enum { AF_CNT_FLAGGED = 0, AF_CNT_HITS = 1, AF_CNT_STATUS = 2, AF_CNT_SEEDED = 3, AF_CNT_SCRATCH = 4 /* internal */, AF_N_COUNTS = 8 };
/* bits of counts[AF_CNT_STATUS] */
#define AF_STATUS_CAND_OVERFLOW 1u
#define AF_STATUS_HIT_OVERFLOW 2u
#define AF_STATUS_LOG_OVERFLOW 4u /* a hit-exchange log region filled up (af_anchor_batch_exchange) */

const char *af_last_error(void);
int af_abi_version(void);
void af_default_params(af_params_t *p);

/* ---- anchor index: replaces `bwa index <anchor.fa>` (Anchored_Fusion.py:167-172) -------- */
/* anchor: ASCII bases (ACGT any case; anything else is N).  kp = 0 picks the default. */
int af_index_build(const char *anchor, int64_t len, const af_params_t *params, int32_t kp, af_index_t **out);
void af_index_free(af_index_t *idx);
int af_index_info(const af_index_t *idx, af_index_info_t *info);
/* 0: the filter words are fingerprint buckets; 1: they hold a blocked Bloom filter (anchors beyond ~12 kb, whose
 * k'-mers overflow the 3-slot buckets) */
int af_index_filter_kind(const af_index_t *idx);
/* raw views for tests (host memory owned by the index) */
const uint32_t *af_index_filter(const af_index_t *idx);
const uint32_t *af_index_table(const af_index_t *idx); /* table_slots x {key, value} */
int af_index_upload(const af_index_t *idx, int device, af_dev_index_t **out);
void af_dev_index_free(af_dev_index_t *d);
int af_dev_index_device(const af_dev_index_t *d);

/* ---- host ingest: replaces kseq/zlib inside bwa (Anchored_Fusion.py:182) ---------------- */
int af_layout(int32_t max_read_len, int64_t n_pairs, af_layout_t *out);
/* seq1/seq2: concatenated ASCII reads, off[i]..off[i+1] delimit read i (n_pairs+1 offsets).
 * lens_out may be NULL; n-lists are filled up to ncap reads (AF_ERR_CAPACITY beyond).
 * *uniform_len_out = common length, or 0 when lengths differ. */
int af_pack_pairs(const char *seq1, const int64_t *off1, const char *seq2, const int64_t *off2,
                  int64_t n_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                  uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                  int64_t *n_nreads_out, int32_t *uniform_len_out);
/* inverse, for tests: codes 0..3 (pad bases come back as their 2-bit value) */
int af_unpack_read(const void *packed, int32_t max_read_len, int64_t read_id, int32_t len, uint8_t *codes_out);

/* Paired FASTQ reader: plain text, gzip (any number of members) and BGZF (bgzip), detected per file.
 * n_threads workers (the reference forwards `--thread` to `bwa mem -t`, Anchored_Fusion.py:29,182;
 * 0 = one per host core) inflate BGZF blocks / copy text, index lines, check CRCs and pack 2-bit tiles in
 * parallel; a plain gzip member is one serial bit stream and is inflated by one thread per file.
 * Record text of the current batch stays valid until the next af_fastq_next / af_fastq_skip. */
int af_fastq_open(const char *path1, const char *path2, af_fastq_t **out);                 /* n_threads = 0 */
int af_fastq_open_threads(const char *path1, const char *path2, int32_t n_threads, af_fastq_t **out);
/* Several file pairs read back to back as ONE stream of pairs (Anchored_Fusion_singlecell.py:86-113: one
 * FASTQ pair per cell): small files are decoded whole, many at a time, and a batch may span many cells. */
int af_fastq_open_multi(const char *const *paths1, const char *const *paths2, int32_t n_files, int32_t n_threads,
                        af_fastq_t **out);
int af_fastq_threads(const af_fastq_t *fq);
void af_fastq_close(af_fastq_t *fq);
/* longest read among the first n_records records (the packed width must be known before reading) */
int af_fastq_peek(const char *path, int32_t n_records, int32_t *max_len_out);
/* Reads up to max_pairs pairs; packs them like af_pack_pairs.  *n_pairs_out = 0 at EOF. */
int af_fastq_next(af_fastq_t *fq, int64_t max_pairs, int32_t max_read_len, int32_t pad_byte, void *packed_out,
                  uint16_t *lens_out, uint32_t *nread_ids_out, uint32_t *nmask_out, int64_t ncap,
                  int64_t *n_nreads_out, int32_t *uniform_len_out, int64_t *n_pairs_out);
/* the same without packing: a rank of a multi-GPU job steps over the batches of the other ranks */
int af_fastq_skip(af_fastq_t *fq, int64_t max_pairs, int64_t *n_pairs_out);
/* record text of read_id of the current batch: name (up to first blank, /1 /2 stripped as bwa
 * does), bases, qualities.  Pointers are not NUL-terminated. */
int af_fastq_record(const af_fastq_t *fq, int64_t read_id, const char **name, int32_t *name_len,
                    const char **seq, const char **qual, int32_t *len);
/* many records at once: name, bases, qualities of read_ids[i] copied back to back into `text`;
 * offs[4i .. 4i+3] = start of name / bases / qualities / end.  AF_ERR_CAPACITY (with *text_used = bytes
 * needed) when text_cap is too small; text may be NULL to size the buffer. */
int af_fastq_records(const af_fastq_t *fq, const int64_t *read_ids, int64_t n, char *text, int64_t text_cap,
                     int64_t *offs, int64_t *text_used);
/* test hook: the CRC-32 the reader checks every inflated BGZF block / gzip member with (PCLMULQDQ folding where the
 * CPU has it, zlib otherwise); equals zlib's crc32(0, buf, len) */
uint32_t af_debug_crc32(const void *buf, int64_t len);
/* test hook: ONE gzip member decoded the way the reader's parallel single-member path decodes it (payload cut into
 * n_chunks pieces, block starts found by search, symbolic decode, chaining, marker resolution, CRC check) */
int af_debug_gunzip_chunks(const void *gz, int64_t n, int32_t n_chunks, void *out, int64_t cap, int64_t *out_len, int32_t *n_redone);
/* index (over the whole run) of the first pair of every input file started so far; returns the count */
int af_fastq_file_starts(const af_fastq_t *fq, int64_t *first_pair_out, int32_t cap);
/* index (over the whole run) of the current batch's first pair */
int64_t af_fastq_batch_first_pair(const af_fastq_t *fq);

/* ---- the hot path on one GPU: replaces `bwa mem -M | samtools view -F 772` -------------- */
/* bytes af_anchor_batch needs for a batch of n_pairs pairs of reads up to max_read_len bases with room for
 * cand_cap flagged reads; af_workspace_bytes assumes AF_MAX_READ_LEN (always sufficient) */
size_t af_workspace_bytes(int64_t n_pairs, int64_t cand_cap);
size_t af_workspace_bytes_len(int64_t n_pairs, int64_t cand_cap, int32_t max_read_len);
/* seed scan -> compaction -> verify -> compaction -> extend -> compaction: 6 kernels, all on `stream`,
 * no host sync.  counts[AF_CNT_FLAGGED] reads passed the scan's filter, counts[AF_CNT_SEEDED] were handed
 * to the extension (with the default verify stage: they hold a true >= k-base exact match),
 * counts[AF_CNT_HITS] are anchored (score >= T).
 * d_hits[0..counts[AF_CNT_HITS]) come back ordered by read_id. */
int af_anchor_batch(const af_dev_index_t *d, const af_batch_t *batch, void *workspace, size_t workspace_bytes,
                    int64_t cand_cap, af_hit_t *d_hits, int64_t hits_cap, uint32_t *d_counts, void *stream);
/* the stages, exposed for tests, profiling and the roofline measurement */
int af_seed_scan(const af_dev_index_t *d, const af_batch_t *batch, uint32_t *d_flags /*2 words per tile*/,
                 void *stream);
/* host twin of the kernel's probe sequence for one pair (2*words_per_read packed words, mate 1 then
 * mate 2): test hook, runs the same template code the kernel runs; not on any product path */
int af_debug_scan_pair(const af_index_t *idx, const uint32_t *words, int32_t words_per_read, int32_t read_len,
                       int32_t with_neighbour_test, int32_t *flag1, int32_t *flag2);
/* tuning knobs.  mode 0/3: threads_per_block of the scan (64..768, 0 = default).  Other modes ignore
 * threads_per_block: 4/5 fused scan+verify kernel on/off; what stands between scan and extension --
 * 7 (default) k_verify, 11 second look inside the scan (refine queue), 8 nothing; 9/10 k_verify_smem / k_verify. */
int af_seed_scan_config(int32_t threads_per_block, int32_t mode);
int64_t af_kernel_launches(void); /* kernels this library has launched since it was loaded */
/* per-stage device time of af_anchor_batch, CUDA events on the launching stream: begin, run,
 * synchronise the stream, end.  ms_out[5] / calls_out[5]: seed scan, flag compaction, verify
 * (+ its compaction), extension, hit compaction. */
void af_profile_begin(void);
int af_profile_end(double *ms_out, int64_t *calls_out);

/* ---- host->device streaming executor (pinned staging, cudaMemcpyAsync, N slots) --------- */
int af_pipeline_create(const af_dev_index_t *d, int64_t slot_pairs, int32_t max_read_len, int32_t n_slots,
                       af_pipeline_t **out);
void af_pipeline_free(af_pipeline_t *p);
/* Anchors a HOST batch of any size: splits it into slot_pairs chunks (tile-aligned), copies
 * each with cudaMemcpyAsync, runs the kernels, copies the hits back.  h_hits ordered by
 * read_id; read_ids are relative to the batch. */
int af_pipeline_run(af_pipeline_t *p, const af_batch_t *host_batch, af_hit_t *h_hits, int64_t hits_cap,
                    int64_t *n_hits_out, int64_t *n_flagged_out);
/* The same for several anchor indexes (all on the pipeline's device) in ONE pass over the host batch: each
 * chunk is copied once and scanned for every index while resident (the reference re-reads
 * both FASTQ files once per gene, Anchored_Fusion.py:126,182).  h_hits / hits_cap / n_hits_out / n_flagged_out
 * are arrays of n_indexes entries. */
int af_pipeline_run_multi(af_pipeline_t *p, int32_t n_indexes, const af_dev_index_t *const *indexes,
                          const af_batch_t *host_batch, af_hit_t *const *h_hits, const int64_t *hits_cap,
                          int64_t *n_hits_out, int64_t *n_flagged_out);
/* Wire format: the same batch without the padding of the tile layout -- per pair 4 * max_read_len bits (mate 1's
 * bases then mate 2's, 2 bit/base) rounded up to whole 32-bit words, e.g. 19 words = 76 bytes for 2 x 150 bp where a
 * tile spends 80; word w of pair (tile*32 + lane) at ((tile * NW + w) * 32 + lane) * 4 bytes.  It is what crosses PCIe
 * in af_pipeline_run_wire: each chunk is expanded into tiles on the device (k_wire_expand) before the kernels run, so
 * the host->device traffic is the algorithmic 76 bytes per pair.  Lengths and N lists travel as in af_batch_t. */
int64_t af_wire_bytes(int32_t max_read_len, int64_t n_pairs);
/* host conversions (whole tiles: the pad lanes of the last tile are carried along); to_packed is the host twin of the
 * device expansion and restores the pad pattern past max_read_len */
int af_wire_from_packed(const void *packed, int32_t max_read_len, int64_t n_pairs, void *wire_out);
int af_wire_to_packed(const void *wire, int32_t max_read_len, int64_t n_pairs, int32_t pad_byte, void *packed_out);
/* device expansion alone (both pointers in device memory), stream-ordered */
int af_wire_expand_device(const void *d_wire, int32_t max_read_len, int64_t n_pairs, int32_t pad_byte, void *d_packed, void *stream);
/* af_pipeline_run with host_batch->packed in wire format; pad_byte = the index's pad pattern (af_index_info) */
int af_pipeline_run_wire(af_pipeline_t *p, const af_batch_t *host_batch, int32_t pad_byte, af_hit_t *h_hits, int64_t hits_cap,
                         int64_t *n_hits_out, int64_t *n_flagged_out);
int64_t af_pipeline_launches(const af_pipeline_t *p); /* kernels launched so far */
int64_t af_pipeline_h2d_bytes(const af_pipeline_t *p); /* bytes copied host -> device so far */
void *af_host_alloc(size_t bytes);                    /* cudaHostAlloc (pinned) */
void af_host_free(void *p);

/* ---- multi-GPU hit exchange over NVLink peer memory ---------------------- *
 * One process per GPU, reads sharded over the ranks; the only exchange on the path is the small
 * list of hit records.  (The reference has no counterpart: it is one process, its "gather" is
 * `samtools view` writing one BAM, Anchored_Fusion.py:194.)  Instead of a collective after the
 * kernels, the last kernel of the path (hit compaction) stores every record straight into a
 * log region on EVERY rank -- plain 16-byte stores through NVLink into buffers opened with CUDA
 * IPC -- so there is no per-batch rendezvous between the ranks and nothing for NCCL to launch.
 *
 * Each rank's buffer holds world x n_slots regions; region (src, slot) is written only by rank
 * `src` from its stream of workspace slot `slot`:
 *     128-byte header { uint64 tail; uint32 status; uint32 n_batches; }
 *     af_hit_t log[log_cap]: per batch one marker { read_id = 0xFFFFFFFF, pos / clip_l|m_len =
 *     low / high 32 bits of pair_base, score_strand word = record count } then its records
 *     (ordered by read_id, read_ids relative to the batch).
 * A consumer reads after every rank has synchronised its streams and the ranks have met at a
 * host barrier (no device-side waiting anywhere). */
typedef struct af_exchange af_exchange_t;
#define AF_IPC_HANDLE_BYTES 64
#define AF_MAX_PEERS 16
#define AF_LOG_MARKER 0xFFFFFFFFu
int af_exchange_create(int device, int32_t rank, int32_t world, int32_t n_slots, int64_t log_cap, af_exchange_t **out);
void af_exchange_free(af_exchange_t *ex);
/* this rank's buffer as an IPC handle, to be passed to the other ranks by any host channel */
int af_exchange_handle(const af_exchange_t *ex, void *handle_out /* AF_IPC_HANDLE_BYTES */);
/* handles: world x AF_IPC_HANDLE_BYTES, entry `rank` is ignored; opens the peers' buffers */
int af_exchange_connect(af_exchange_t *ex, const void *handles);
/* empty this rank's logs on every rank (stream-ordered; call between host barriers) */
int af_exchange_reset(af_exchange_t *ex, void *stream);
/* af_anchor_batch + append of the batch's records to log (rank, slot) on every rank.  `stream`
 * must be the same for all calls with one `slot`.  pair_base: job-wide index of the batch's first pair. */
int af_anchor_batch_exchange(const af_dev_index_t *d, const af_batch_t *batch, void *workspace, size_t workspace_bytes,
                             int64_t cand_cap, af_hit_t *d_hits, int64_t hits_cap, uint32_t *d_counts,
                             af_exchange_t *ex, int32_t slot, int64_t pair_base, void *stream);
/* synchronous copy of log (src_rank, slot) of THIS rank's buffer to the host: markers + records */
int af_exchange_read(const af_exchange_t *ex, int32_t src_rank, int32_t slot, af_hit_t *h_out, int64_t cap,
                     int64_t *n_out, uint32_t *status_out, uint32_t *n_batches_out);

/* ---- genome pass of the contiguity filter ----------------------------------- *
 * Replaces `bwa mem -M -t T <genome> <w>_del_tmp.fa` inside del_too_many_reads (functions.py:716): the 2-op
 * anchored reads are aligned to the whole GENOME under the same anchoring spec (k = 19 exact seed, ungapped
 * X-drop extension, one primary record per read).  No genome index exists: the genome lives 2 bit/base in HBM
 * (human: 0.8 GB + 0.4 GB N bitmap) and every pass STREAMS it once past a shared-memory filter holding the
 * k'-mers of ~90 reads (both orientations); hits are verified against an exact table, every seeded diagonal is
 * extended by one warp, the best record per read is kept.
 * The genome is one sequence: [256 N] contig 0 [256 N] contig 1 ... [>= 256 N]; positions in records are 1-based
 * in that sequence, af_genome_contig gives the contigs' starts. */
typedef struct af_genome af_genome_t;
typedef struct {
    int64_t pos;           /* 1-based leftmost aligned base in the concatenated genome */
    uint32_t read_id;      /* index of the read in the call */
    uint16_t clip_l, m_len, clip_r;
    uint16_t score_strand; /* score*2 + strand; strand 1 == the read's reverse complement lies on the genome */
    uint32_t reserved;
} af_genome_hit_t;
typedef struct {
    int64_t genome_bases;  /* bases streamed per pass */
    int64_t n_candidates;  /* genome samples that passed the filter, all passes */
    int64_t n_seeds;       /* (read, strand, diagonal) runs of >= k matches handed to the extension */
    int32_t n_passes;
    int32_t n_retries;     /* passes repeated with larger buffers */
    double scan_ms;        /* device time of the genome scans (CUDA events), all passes */
    double total_ms;       /* device time of the whole call */
    double host_index_ms;  /* host time spent building the per-pass filters and tables */
} af_genome_stats_t;
#define AF_GENOME_SEP 256
/* FASTA (plain or gzip); contig name = header up to the first blank */
int af_genome_from_fasta(const char *path, int device, af_genome_t **out);
/* n contigs given as ASCII strings (ACGT any case; anything else is N) */
int af_genome_from_contigs(const char *const *names, const char *const *seqs, const int64_t *lens, int32_t n, int device,
                           af_genome_t **out);
/* test hook: the host half of af_genome_from_fasta alone (parse + pack, no GPU): length of the concatenation, number of
 * contigs and an FNV-1a checksum over the base codes 0..4 of the concatenation */
int af_debug_genome_fasta(const char *path, int64_t *total_len, int32_t *n_contigs, uint64_t *checksum);
/* measurement input: one contig "synth" whose base x is af_synth's random reference base f(seed, x), generated on the device */
int af_genome_synth(uint64_t seed, int64_t len, int device, af_genome_t **out);
void af_genome_free(af_genome_t *g);
int64_t af_genome_length(const af_genome_t *g); /* separators included */
int32_t af_genome_n_contigs(const af_genome_t *g);
int af_genome_contig(const af_genome_t *g, int32_t i, const char **name, int64_t *start /* 0-based in the concatenation */, int64_t *len);
/* reads: concatenated ASCII, offs[i]..offs[i+1] delimit read i (<= AF_GENOME_MAX_READ_LEN bases each).  hits_out holds up to
 * n_reads records, ordered by read_id; reads without a record are unaligned.  reads_per_pass = 0 picks the default.
 * Calls on one genome object are serialised (they share its device buffers); use one object per thread for concurrency. */
int af_genome_align(af_genome_t *g, const char *reads, const int64_t *offs, int64_t n_reads, const af_params_t *params,
                    int32_t reads_per_pass, af_genome_hit_t *hits_out, int64_t *n_hits_out, af_genome_stats_t *stats);

/* ---- seeded synthetic reads (measurement) --------------------------------- */
typedef struct {
    uint64_t seed;
    int64_t ref_len;       /* random reference, base x = f(seed, x) */
    int64_t anchor_start;  /* anchor = ref[anchor_start, anchor_start+anchor_len) */
    int32_t anchor_len;
    int32_t read_len;
    int32_t frag_mean, frag_sd;
    uint32_t sub_ppm;      /* per-base substitution rate, parts per million */
    uint32_t fusion_ppm;   /* fraction of fragments that are anchor|elsewhere chimeras */
    uint32_t n_ppm;        /* per-base N rate (host generator only) */
    uint32_t reserved;
} af_synth_t;
int af_synth_anchor(const af_synth_t *s, char *ascii_out);
/* codes (0..4), n_pairs x read_len each */
int af_synth_pairs_host(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, uint8_t *mate1, uint8_t *mate2);
/* packed tiles straight into HBM (n_ppm must be 0); pair p of the batch = global pair first_pair + p */
int af_synth_pairs_device(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, int32_t pad_byte,
                          void *d_packed, void *stream);

#ifdef __cplusplus
}
#endif
#endif
