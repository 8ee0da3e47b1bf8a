/*
 * af_oracle.c -- CPU ORACLE for the read-anchoring pass.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under anchored_fusion_b200/ may import, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / reported CPU baseline.
 *
 * What it restates.  In the reference the anchoring pass is a shell-out:
 *     bwa index <anchor.fa>                               Anchored_Fusion.py:172
 *     bwa mem -M -t T <anchor.fa> fq1 fq2 | samtools ...  Anchored_Fusion.py:182
 *     samtools view -F 772 ...  (mapped primary reads)    Anchored_Fusion.py:194
 * bwa (README.md:18, "bwa >= 0.7.17") is a third-party C program that is NOT
 * under /root/reference and is not installed here, and the reference holds no
 * expected output for this stage.  PARITY UNPINNED at the bwa boundary: this
 * oracle restates the published bwa-mem algorithm for the indel-free case --
 * exact seeds of length >= k (bwa-mem -k 19), ungapped extension with
 * bwa-mem's default scores (-A 1 -B 4), its z-drop (-d 100), its end-clipping
 * rule (-L 5: extend to the read end iff the end-to-end score is greater than
 * best local score - 5) and its report threshold (-T 30) -- as the north star
 * (BASELINE.json) specifies: "k-mer seed -> integer-scored ungapped/X-drop
 * extension".  What IS pinned by the reference is the interpretation of the
 * records this pass emits (functions.py:656-702 deal_cigar, :892-950
 * contact_reads); see oracle/ref_bridge.py and tests/golden/.  The positions
 * it reports on the reference's bundled sample are also checked against the
 * ground truth in the wgsim read names of that sample (tests/test_oracle.py),
 * and every record of that sample attains the optimum of an exhaustive
 * affine-gap DP under bwa-mem's scores and clip penalty (af_sw.c,
 * tests/test_oracle_vs_sw.py).
 *
 * Semantics (frozen; DESIGN.md "Anchoring spec v1"):
 *   anchor a[0..G), read r[0..L): base codes 0..3 = A,C,G,T; 4 = N/other.
 *   oriented read q_s:  q_0 = r;  q_1[i] = comp(r[L-1-i])  (comp(4)=4), i.e.
 *   strand 1 == the read's reverse complement lies on the anchor's forward
 *   strand (SAM FLAG 0x10; SEQ is printed reverse-complemented, cf.
 *   functions.py:498-504 for the alphabet).
 *   match(s,i,d) := q_s[i] < 4 and 0 <= i+d < G and q_s[i] == a[i+d].
 *   A diagonal (s,d) is SEEDED iff it holds k consecutive matches; its seed is
 *   the LEFTMOST such window [qb0, qb0+k).  The definition is semantic: it does
 *   not depend on how an implementation enumerates k-mers.
 *   Extension from the seed, left first, then right (bwa-mem order), scores
 *   h0 = k*A, +A per match, -B per mismatch (N = mismatch):
 *       cur += delta; if cur <= 0 stop; if cur > max {max = cur; off = j+1};
 *       if j+1 == qlen: gscore = cur; if max - cur > X stop.
 *   Clip rule: if gscore <= 0 or gscore <= max - clip: local end at off, score
 *   max; else extend to the read end, score gscore.  The right extension
 *   starts from the score the left extension returned.
 *   A read is ANCHORED iff its best diagonal scores >= T.  Best = highest
 *   score, then strand 0 before 1, then smaller d.
 *   Record: read_id, POS = qb + d + 1 (1-based, anchor-forward), clipL = qb,
 *   M = qe - qb, clipR = L - qe, strand, score  ->  CIGAR clipL S M M clipR S,
 *   which is what deal_cigar (functions.py:656) / contact_reads (:917-930) eat.
 *
 * Schedule: a sorted array of ALL forward-anchor k-mers (k = 19), every read
 * position probed on both orientations, every distinct seeded diagonal
 * evaluated.  Deliberately different from the GPU schedule (strided short
 * k-mers + verification) so that the two check one another.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t k;      /* min seed length        (bwa mem -k, 19) */
    int32_t A;      /* match score            (-A, 1)  */
    int32_t B;      /* mismatch penalty       (-B, 4)  */
    int32_t clip5;  /* 5' clip penalty        (-L, 5)  */
    int32_t clip3;  /* 3' clip penalty        (-L, 5)  */
    int32_t T;      /* min score to report    (-T, 30) */
    int32_t X;      /* z-drop                 (-d, 100) */
} afo_params;

typedef struct {
    uint32_t read_id;       /* index of the read in the input array */
    int32_t pos;            /* 1-based leftmost aligned anchor base */
    uint16_t clip_l;
    uint16_t m_len;
    uint16_t clip_r;
    uint16_t score_strand;  /* score*2 + strand */
} afo_hit;

typedef struct { uint64_t kmer; int32_t pos; } afo_ent;

static int ent_cmp(const void *x, const void *y) {
    const afo_ent *a = (const afo_ent *)x, *b = (const afo_ent *)y;
    if (a->kmer != b->kmer) return a->kmer < b->kmer ? -1 : 1;
    return (a->pos > b->pos) - (a->pos < b->pos);
}

/* every forward anchor k-mer without N, sorted by code */
static afo_ent *build_index(const uint8_t *a, int32_t G, int k, int64_t *n_out) {
    afo_ent *e = (afo_ent *)malloc(sizeof(afo_ent) * (size_t)(G > 0 ? G : 1));
    int64_t n = 0;
    uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1), km = 0;
    int run = 0;
    for (int32_t i = 0; i < G; i++) {
        if (a[i] < 4) { km = ((km << 2) | a[i]) & mask; run++; } else { run = 0; km = 0; }
        if (run >= k) { e[n].kmer = km; e[n].pos = i - k + 1; n++; }
    }
    qsort(e, (size_t)n, sizeof(afo_ent), ent_cmp);
    *n_out = n;
    return e;
}

/* A presence bitmap over a hash of the anchor's k-mers: a read window whose bit is clear cannot be in
 * the sorted index, so the binary search is skipped.  Purely a shortcut of the lookup -- the set of
 * (window, anchor position) matches is unchanged -- that makes the full-size parity runs affordable. */
#define AFO_PRESENT_BITS (1u << 22)
static inline uint32_t present_slot(uint64_t km) { return (uint32_t)((km * 0x9E3779B97F4A7C15ULL) >> 42); }
static uint64_t *build_present(const afo_ent *e, int64_t n) {
    uint64_t *b = (uint64_t *)calloc(AFO_PRESENT_BITS / 64, sizeof(uint64_t));
    for (int64_t i = 0; i < n; i++) { uint32_t s = present_slot(e[i].kmer); b[s >> 6] |= 1ULL << (s & 63); }
    return b;
}

static int64_t lower_bound(const afo_ent *e, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (e[mid].kmer < key) lo = mid + 1; else hi = mid; }
    return lo;
}

/* one extension step loop; q/t walk with stride dq (+1 right, -1 left) */
static void extend(const uint8_t *q, const uint8_t *a, int32_t G, int32_t qi, int32_t ai, int dir,
                   int32_t qlen, int32_t n, int32_t h0, const afo_params *P,
                   int32_t *max_out, int32_t *off_out, int32_t *g_out) {
    int32_t cur = h0, mx = h0, off = 0, g = -1;
    (void)G;
    for (int32_t j = 0; j < n; j++) {
        uint8_t qb = q[qi + dir * j], ab = a[ai + dir * j];
        cur += (qb < 4 && qb == ab) ? P->A : -P->B;
        if (cur <= 0) break;
        if (cur > mx) { mx = cur; off = j + 1; }
        if (j + 1 == qlen) g = cur;
        if (mx - cur > P->X) break;
    }
    *max_out = mx; *off_out = off; *g_out = g;
}

/* Evaluate diagonal d of oriented read q (length L).  Returns 0 if unseeded. */
static int diag_eval(const uint8_t *q, int32_t L, const uint8_t *a, int32_t G, int32_t d,
                     const afo_params *P, int32_t *score, int32_t *qb_out, int32_t *qe_out) {
    int32_t lo = d < 0 ? -d : 0, hi = (G - d < L) ? G - d : L, run = 0, qb0 = -1;
    for (int32_t i = lo; i < hi; i++) {
        if (q[i] < 4 && q[i] == a[i + d]) { if (++run >= P->k) { qb0 = i - P->k + 1; break; } }
        else run = 0;
    }
    if (qb0 < 0) return 0;
    int32_t sc = P->k * P->A, qb = 0, qe = L, mx, off, g;
    if (qb0 > 0) {                                    /* left extension */
        int32_t n = qb0 < qb0 + d ? qb0 : qb0 + d;    /* bases left in read / in anchor */
        extend(q, a, G, qb0 - 1, qb0 - 1 + d, -1, qb0, n, sc, P, &mx, &off, &g);
        if (g <= 0 || g <= mx - P->clip5) { qb = qb0 - off; sc = mx; } else { qb = 0; sc = g; }
    }
    int32_t qe0 = qb0 + P->k;
    if (qe0 < L) {                                    /* right extension */
        int32_t nr = L - qe0, na = G - (qe0 + d), n = nr < na ? nr : na;
        extend(q, a, G, qe0, qe0 + d, +1, L - qe0, n, sc, P, &mx, &off, &g);
        if (g <= 0 || g <= mx - P->clip3) { qe = qe0 + off; sc = mx; } else { qe = L; sc = g; }
    }
    *score = sc; *qb_out = qb; *qe_out = qe;
    return 1;
}

static int i32_cmp(const void *x, const void *y) {
    int32_t a = *(const int32_t *)x, b = *(const int32_t *)y;
    return (a > b) - (a < b);
}

/* Anchor one read.  Returns 1 and fills *h iff anchored. */
typedef struct { int32_t *v; size_t cap; } afo_scratch;

static int anchor_read(const uint8_t *r, int32_t L, const uint8_t *a, int32_t G, const afo_ent *idx,
                       int64_t nidx, const uint64_t *present, const afo_params *P, uint8_t *q, afo_scratch *S, afo_hit *h) {
    int k = P->k, found = 0;
    int32_t best_sc = -1, best_qb = 0, best_qe = 0, best_d = 0, best_s = 0;
    if (L < k) return 0;
    uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    for (int s = 0; s < 2; s++) {
        if (s == 0) memcpy(q, r, (size_t)L);
        else for (int32_t i = 0; i < L; i++) { uint8_t c = r[L - 1 - i]; q[i] = c < 4 ? (uint8_t)(3 - c) : 4; }
        int32_t nd = 0, run = 0;
        uint64_t km = 0;
        for (int32_t i = 0; i < L; i++) {
            if (q[i] < 4) { km = ((km << 2) | q[i]) & mask; run++; } else { run = 0; km = 0; }
            if (run < k) continue;
            int32_t qpos = i - k + 1;
            const uint32_t ps = present_slot(km);
            if (!((present[ps >> 6] >> (ps & 63)) & 1ULL)) continue;
            for (int64_t e = lower_bound(idx, nidx, km); e < nidx && idx[e].kmer == km; e++) {
                if ((size_t)nd == S->cap) { S->cap *= 2; S->v = (int32_t *)realloc(S->v, sizeof(int32_t) * S->cap); }
                S->v[nd++] = idx[e].pos - qpos;
            }
        }
        if (!nd) continue;
        int32_t *diags = S->v;
        qsort(diags, (size_t)nd, sizeof(int32_t), i32_cmp);
        for (int32_t t = 0; t < nd; t++) {
            if (t && diags[t] == diags[t - 1]) continue;
            int32_t sc, qb, qe;
            if (!diag_eval(q, L, a, G, diags[t], P, &sc, &qb, &qe)) continue;
            /* best: score desc, strand asc, diagonal asc (loops already run in that order) */
            if (sc > best_sc) { best_sc = sc; best_qb = qb; best_qe = qe; best_d = diags[t]; best_s = s; found = 1; }
        }
    }
    if (!found || best_sc < P->T) return 0;
    h->pos = best_qb + best_d + 1;
    h->clip_l = (uint16_t)best_qb;
    h->m_len = (uint16_t)(best_qe - best_qb);
    h->clip_r = (uint16_t)(L - best_qe);
    h->score_strand = (uint16_t)(best_sc * 2 + best_s);
    return 1;
}

/*
 * reads: n_reads rows of `stride` base codes; lens == NULL means every read has
 * length `stride`.  out must hold cap records; hits come back ordered by read_id.
 * Returns 0, or -1 if cap was too small (n_out then holds the needed count).
 */
int afo_anchor_reads(const uint8_t *anchor, int32_t G, const uint8_t *reads, const uint16_t *lens,
                     int64_t n_reads, int32_t stride, const afo_params *P, afo_hit *out, int64_t cap,
                     int64_t *n_out, int nthreads) {
    int64_t nidx = 0;
    afo_ent *idx = build_index(anchor, G, P->k, &nidx);
    uint64_t *present = build_present(idx, nidx);
    uint8_t *ok = (uint8_t *)calloc((size_t)(n_reads > 0 ? n_reads : 1), 1);
    afo_hit *tmp = (afo_hit *)malloc(sizeof(afo_hit) * (size_t)(n_reads > 0 ? n_reads : 1));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        uint8_t *q = (uint8_t *)malloc((size_t)(stride > 0 ? stride : 1));
        afo_scratch S;
        S.cap = 1024;
        S.v = (int32_t *)malloc(sizeof(int32_t) * S.cap);
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < n_reads; i++) {
            int32_t L = lens ? lens[i] : stride;
            afo_hit h;
            if (anchor_read(reads + i * (int64_t)stride, L, anchor, G, idx, nidx, present, P, q, &S, &h)) {
                h.read_id = (uint32_t)i;
                tmp[i] = h;
                ok[i] = 1;
            }
        }
        free(q); free(S.v);
    }
    int64_t n = 0;
    for (int64_t i = 0; i < n_reads; i++) if (ok[i]) { if (n < cap) out[n] = tmp[i]; n++; }
    *n_out = n;
    free(idx); free(present); free(ok); free(tmp);
    return n <= cap ? 0 : -1;
}

/* exposed for unit tests: evaluate one diagonal of an already-oriented read */
int afo_diag_eval(const uint8_t *q, int32_t L, const uint8_t *a, int32_t G, int32_t d,
                  const afo_params *P, int32_t *score, int32_t *qb, int32_t *qe) {
    return diag_eval(q, L, a, G, d, P, score, qb, qe);
}

void afo_default_params(afo_params *P) {
    P->k = 19; P->A = 1; P->B = 4; P->clip5 = 5; P->clip3 = 5; P->T = 30; P->X = 100;
}
