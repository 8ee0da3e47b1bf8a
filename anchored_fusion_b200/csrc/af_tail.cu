// af_tail.cu -- k_tail: everything behind the seed scan in ONE persistent kernel.
//
// The seed scan (k_seed_scan<..., EMIT>) leaves the flagged reads as a candidate stream: 32-record chunks of
// {read_id, packed words}, grouped by region (a contiguous tile range) through per-region chunk directories
// (af_emit, af_device.cuh).  One CTA of k_tail owns a region at a time:
//
//   phase 1  warps take the region's chunks from a shared-memory counter.  A lane = a candidate: its record
//            is read coalesced (mostly from L2 -- the scan wrote it microseconds ago), the exact SEEDED test
//            of k_verify_smem runs on it (half-size filter in shared memory -> neighbour k'-mer pre-test ->
//            exact table walk -> >= k run against the 2-bit packed anchor).  Seeded reads (4 % of the
//            candidates) are extended on the spot by the whole warp (k_extend's diagonal evaluation and
//            X-drop extension, the anchor's base codes in shared memory); an anchored read's 16-byte record
//            overwrites the head of its own candidate record and sets its bit in the region's hit bitmap
//            (shared memory, one bit per read of the region).
//   phase 2  prefix popcounts over the bitmap give every hit its rank inside the region; the region
//            publishes its hit count and sums the counts of the regions before it (they are being worked on
//            by the other CTAs at the same time; regions are dealt to CTAs round-robin, so the wait is short
//            and cannot deadlock).
//   phase 3  hits are copied to hits[base + rank]: the list comes out ordered by read_id, as the
//            six-kernel path produced it, with no sort and no further kernel.  With a hit exchange attached
//            the records also go to this rank's log on every GPU (coalesced 16-byte peer stores); the CTA of
//            the last region writes the marker, the headers and the new tail.
//
// Replaces k_flag_scatter, k_verify_smem, k_sel_scatter, k_extend and k_hit_scatter (95 us of kernels and
// 20 us of gaps per 10 M pairs; 68 MB of scattered DRAM reads for the flagged reads' quads).
// Same semantics as before: "Anchoring spec v1" (DESIGN.md), bit-exact against oracle/af_oracle.c.
#include <cuda_runtime.h>

#include "af_common.h"
#include "af_device.cuh"

static const int TAIL_THREADS = 1024;
static const int REG_WORDS = AF_REG_TILES * 2;          // hit bitmap: 64 reads per tile
static const int PREF_WORDS = REG_WORDS / 4 + 4;        // prefix popcount per 4 bitmap words

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef AF_TAIL_EXP
#define AF_TAIL_EXP 0
#endif
#ifndef AF_TAIL_PROF
#define AF_TAIL_PROF 0                                  // 1: thread_extend counts cycles per part into g_prof_* (tools/tail_timing.py)
#endif
#if AF_TAIL_PROF
#define PROF_T0 const long long prof_t0_ = clock64();
#define PROF_ADD(x) x += clock64() - prof_t0_;
#else
#define PROF_T0
#define PROF_ADD(x)
#endif
struct ExtProf { long long collect, eval, n_eval, n_load, mask, run, ext, n_iter, t_start, t_consume; };

static const int APK_PAD_WORDS = 16;                    // zero words in front of the padded packed anchor (256 bases)
static const int APN_PAD_WORDS = 8;                     // ... and in front of the anchor's N bitmask (256 bases)

// 32 bases (64 bits) of the padded 2-bit packed anchor starting at base pos >= -256
__device__ __forceinline__ unsigned long long apk_window(const uint32_t *ap, int pos) {
    const int p = pos + 16 * APK_PAD_WORDS, wi = p >> 4, sh = 2 * (p & 15);
    const uint32_t w0 = ap[wi], w1 = ap[wi + 1], w2 = ap[wi + 2];
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}

// bits 0, 2, 4 .. 30 of x -> bits 0 .. 15
__device__ __forceinline__ uint32_t even_bits(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// ---- extension, one THREAD per read ------------------------------------------------------------
// k_extend gives a read a whole warp; most of its instructions do the same thing in 32 lanes, and with ~120
// reads to extend per SM and region the warp-level instruction stream -- not latency -- was the limit
// (measured: 11 k cycles per read, 42 k cycles for the phase).  Here a lane owns a read:
//   * the 256-bit match mask of a diagonal is 8 XORs of 64-bit windows (read words from shared memory against
//     the 2-bit packed anchor, forward or reverse complement) instead of 256 base tests;
//   * the leftmost run of k matches comes from AND-ing shifted copies of the mask words;
//   * the X-drop extension walks the MISMATCHES of the mask, not the bases: between two mismatches the score
//     only rises, so the maximum moves only at the end of a match run and every stop condition (score <= 0,
//     max - score > X) can only fire on a mismatch.  A 150-base read with two substitutions takes three steps.
// Same results as oracle/af_oracle.c::extend / diag_eval, step for step (spec v1).

// first position j in [pos, limit) whose mask bit is 0, or limit
__device__ __forceinline__ int mask_next_zero(const uint32_t *m, int pos, int limit) {
    while (pos < limit) {
        const int wi = pos >> 5;
        const uint32_t z = ~m[wi] & (FULL << (pos & 31));
        if (z) return min(wi * 32 + __ffs(z) - 1, limit);
        pos = (wi + 1) * 32;
    }
    return limit;
}
// last position j in (limit, pos] whose mask bit is 0, or limit (limit >= -1)
__device__ __forceinline__ int mask_prev_zero(const uint32_t *m, int pos, int limit) {
    while (pos > limit) {
        const int wi = pos >> 5;
        const uint32_t z = ~m[wi] & (FULL >> (31 - (pos & 31)));
        if (z) return max(wi * 32 + 31 - __clz(z), limit);
        pos = wi * 32 - 1;
    }
    return limit;
}

// One direction of the ungapped X-drop extension over mask positions start, start + dir, ... (n steps), from
// score h0 (> 0); qlen = read bases left on this side.  oracle/af_oracle.c::extend, mismatch by mismatch.
__device__ __forceinline__ void thread_extend_dir(const uint32_t *m, int start, int dir, int n, int qlen, int h0,
                                                  const ExtParams &P, int &mx_out, int &off_out, int &g_out) {
    int cur = h0, mx = h0, off = 0, g = -1, j = 0;
    for (;;) {
        const int jz = dir > 0 ? mask_next_zero(m, start + j, start + n) - start
                               : start - mask_prev_zero(m, start - j, start - n);   // step of the next mismatch, n if none
        if (jz > j) {                                       // a run of matches: steps j .. jz-1
            cur += P.A * (jz - j);
            if (cur > mx) { mx = cur; off = jz; }
            if (jz == qlen) g = cur;                         // the run ends on the read's last base
        }
        if (jz >= n) break;
        cur -= P.B;                                          // step jz: mismatch
        if (cur <= 0) break;
        if (jz + 1 == qlen) g = cur;
        if (mx - cur > P.X) break;
        j = jz + 1;
    }
    mx_out = mx; off_out = off; g_out = g;
}

// Evaluate diagonal (s, d) of the read whose words sit in shared memory at sw[t * VT].  Returns the score or -1 if
// the diagonal holds no run of k matches.  nm: the read's N-mask words (forward coordinates) or nullptr.
__device__ __forceinline__ int thread_eval_diag(int s, int d, int L, const uint32_t *sw, int VT, const uint32_t *nm,
                                                const uint32_t *apk0p, const uint32_t *apk1p, const uint32_t *apn0p,
                                                const uint32_t *apn1p, int G, const ExtParams &P, int &qb_out, int &qe_out, ExtProf &pf) {
    const int nw = (L + 31) >> 5;
    long long tp0 = AF_TAIL_PROF ? clock64() : 0;
    const int dd = s ? G - L - d : d;                       // forward read base j lies on strand-s anchor base j + dd
    const uint32_t *ap = s ? apk1p : apk0p, *an = s ? apn1p : apn0p;   // an: 1 bit per anchor base that is N, or nullptr
    const int lo = max(0, -dd), hi = min(L, G - dd);        // read bases that face an anchor base: [lo, hi)
    uint32_t mf[9], mo[9];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint32_t w = 0;
        const int b0 = max(lo - 32 * c, 0), b1 = min(hi - 32 * c, 32);
        if (c < nw && b1 > b0) {                            // (then 32c + dd lies in (-32, G): inside the padded array)
            const unsigned long long x = ((unsigned long long)sw[2 * c * VT] | ((unsigned long long)sw[(2 * c + 1) * VT] << 32)) ^
                                         apk_window(ap, 32 * c + dd);
            const unsigned long long ne = x | (x >> 1);     // even bits: 1 = bases differ
            const uint32_t eq = ~(even_bits((uint32_t)ne) | (even_bits((uint32_t)(ne >> 32)) << 16));
            w = eq & ((b1 - b0 >= 32 ? FULL : ((1u << (b1 - b0)) - 1u)) << b0);
            if (nm) w &= ~nm[c];
            if (an) {                                       // the packed anchor holds A where the anchor has N
                const int pn = 32 * c + dd + 32 * APN_PAD_WORDS;
                w &= ~__funnelshift_r(an[pn >> 5], an[(pn >> 5) + 1], pn & 31);
            }
        }
        mf[c] = w;
    }
    mf[8] = 0;
    const uint32_t *m = mf;
    if (s) {
        // oriented position i = L - 1 - j: mask(i) = Rev(i + 256 - L), Rev = the 256-bit mask bit-reversed
        // (Rev word t = brev(forward word 7 - t))
        const int sh = 256 - L, ws = sh >> 5, bs = sh & 31;
        for (int c = 0; c < 8; c++) {
            const int w0 = c + ws;
            const uint32_t a0 = w0 < 8 ? __brev(mf[7 - w0]) : 0u, a1 = w0 + 1 < 8 ? __brev(mf[6 - w0]) : 0u;
            mo[c] = __funnelshift_r(a0, a1, bs);
        }
        mo[8] = 0;
        m = mo;
    }
    if (AF_TAIL_PROF) { const long long t = clock64(); pf.mask += t - tp0; tp0 = t; }
    // leftmost run of k <= 32 matches: starts b < 32 of word c with bits b .. b+k-1 set in (word c, word c+1)
    int qb0 = -1;
    for (int c = 0; c < nw && qb0 < 0; c++) {
        const unsigned long long v = (unsigned long long)m[c] | ((unsigned long long)m[c + 1] << 32);
        unsigned long long acc = ~0ull, p = v;
        for (int kk = P.k, off = 0, len = 1; kk; kk >>= 1, len <<= 1) {
            if (kk & 1) { acc &= p >> off; off += len; }
            p &= p >> len;
        }
        const uint32_t starts = (uint32_t)acc;
        if (starts) qb0 = c * 32 + __ffs(starts) - 1;
    }
    if (AF_TAIL_PROF) { const long long t = clock64(); pf.run += t - tp0; tp0 = t; }
    if (qb0 < 0 || qb0 + P.k > L) return -1;
    int sc = P.k * P.A, qb = 0, qe = L, mx, off, g;
    if (qb0 > 0) {
        const int n = min(qb0, qb0 + d);
        thread_extend_dir(m, qb0 - 1, -1, n, qb0, sc, P, mx, off, g);
        if (g <= 0 || g <= mx - P.clip5) { qb = qb0 - off; sc = mx; } else { qb = 0; sc = g; }
    }
    const int qe0 = qb0 + P.k;
    if (qe0 < L) {
        const int n = min(L - qe0, G - (qe0 + d));
        thread_extend_dir(m, qe0, +1, n, L - qe0, sc, P, mx, off, g);
        if (g <= 0 || g <= mx - P.clip3) { qe = qe0 + off; sc = mx; } else { qe = L; sc = g; }
    }
    if (AF_TAIL_PROF) pf.ext += clock64() - tp0;
    qb_out = qb; qe_out = qe;
    return sc;
}

// All of k_extend for one read, by one thread: exact table lookups for every sample, each distinct diagonal
// evaluated, best = score desc, strand 0 first, smaller d.  Returns true and the record if it scores >= T.
// The lanes of a warp work on different reads, so the code is arranged for convergence: diagonals are first
// COLLECTED (table walks, cheap, divergent) into a short list and then evaluated in one loop with a single call
// site, so that all lanes run thread_eval_diag together (inlining it into the table walk made every lane run it
// alone: 100 k cycles per read instead of 10 k, measured).
__device__ __forceinline__ bool thread_extend(bool active, uint32_t rid, int L, const uint32_t *sw, int VT, const uint32_t *nm,
                                              const uint2 *__restrict__ table, uint32_t tmask, const uint32_t *apk0p,
                                              const uint32_t *apk1p, const uint32_t *apn0p, const uint32_t *apn1p, int G,
                                              int KP, int S, const ExtParams &P, uint4 &out, ExtProf &pf) {
    // Called by ALL 32 lanes of a warp (`active` = this lane has a read).  Every loop below runs a warp-uniform
    // number of times (__any_sync), lanes that are done are predicated off: left to itself the compiler let the
    // lanes drift apart in the table-walk state machine and each lane ended up running it alone (75 k cycles per
    // read, measured with AF_TAIL_PROF).
    const uint32_t kpmask = (1u << (2 * KP)) - 1u;
    constexpr int CAP = 8;
    int best_sc = -1, best_qb = 0, best_qe = 0;
    uint32_t best_key = 0xFFFFFFFFu, seen0 = 0xFFFFFFFFu, seen1 = 0xFFFFFFFFu;   // the last two diagonals evaluated
    uint32_t dl[CAP];
    int nd = 0, j = 0, p = 0;
    const int nprobe = active && L >= KP ? (L - KP) / S + 1 : 0;
    uint32_t key = 0, sl = 0;
    uint2 e0 = make_uint2(AF_T_EMPTY, 0), e1 = e0;
    bool walking = false;
    for (;;) {
        long long tp0 = AF_TAIL_PROF ? clock64() : 0;
        for (;;) {                                           // collect: walk the table for sample after sample
            const bool want = nd < CAP && (walking || j < nprobe);
            if (!__any_sync(FULL, want)) break;
            if (AF_TAIL_PROF) pf.n_iter++;
            const long long ti0 = AF_TAIL_PROF ? clock64() : 0;
            const bool was_walking = walking;
            if (want) {
                if (!walking) {
                    p = j * S;
                    bool skip = false;
                    if (nm) {                                // a k'-mer that overlaps an N is no seed material
                        const uint32_t nn = __funnelshift_r(nm[p >> 5], (p >> 5) + 1 < AF_NMASK_WORDS ? nm[(p >> 5) + 1] : 0u, p & 31);
                        skip = (nn & ((1u << KP) - 1u)) != 0;
                    }
                    if (skip) j++;
                    else {
                        const int o = 2 * p;
                        key = __funnelshift_r(sw[(o >> 5) * VT], sw[((o >> 5) + 1) * VT], o & 31) & kpmask;
                        sl = af_table_hash(key, tmask);
                        e0 = table[sl]; e1 = table[(sl + 1) & tmask];   // the usual walk is {match, empty}: both loads in flight
                        walking = true;
                        if (AF_TAIL_PROF) pf.n_load += 2;
                    }
                } else {
                    const uint2 e = e0;
                    if (e.x == AF_T_EMPTY) { walking = false; j++; }
                    else {
                        sl = (sl + 1) & tmask;
                        e0 = e1;
                        if (e0.x != AF_T_EMPTY) e1 = table[(sl + 1) & tmask];
                        if (e.x == key) {
                            const int s = e.y >> 31, jpos = (int)(e.y & 0x7FFFFFFFu);
                            const int d = jpos - (s ? L - p - KP : p);
                            const uint32_t dk = ((uint32_t)s << 31) | (uint32_t)(d + 1024);
                            bool known = dk == seen0 || dk == seen1;
                            for (int t = 0; t < nd; t++) known |= dl[t] == dk;
                            if (!known) dl[nd++] = dk;
                        }
                    }
                }
            }
            if (AF_TAIL_PROF) { if (was_walking) pf.t_consume += clock64() - ti0; else pf.t_start += clock64() - ti0; }
        }
        if (AF_TAIL_PROF) { const long long t = clock64(); pf.collect += t - tp0; tp0 = t; }
        if (!__any_sync(FULL, nd > 0)) break;
        for (int t = 0; __any_sync(FULL, t < nd); t++) {     // evaluate: one call site, all lanes together
            if (t < nd) {
                const uint32_t dk = dl[t];
                const int s = dk >> 31, d = (int)(dk & 0x7FFFFFFFu) - 1024;
                int qb, qe;
                const int sc = thread_eval_diag(s, d, L, sw, VT, nm, apk0p, apk1p, apn0p, apn1p, G, P, qb, qe, pf);
                if (AF_TAIL_PROF) pf.n_eval++;
                if (sc > best_sc || (sc == best_sc && sc >= 0 && dk < best_key)) { best_sc = sc; best_qb = qb; best_qe = qe; best_key = dk; }
                seen1 = seen0; seen0 = dk;
            }
            __syncwarp();
        }
        if (AF_TAIL_PROF) pf.eval += clock64() - tp0;
        nd = 0;
    }
    if (best_sc < P.T) return false;
    const int s = best_key >> 31, d = (int)(best_key & 0x7FFFFFFFu) - 1024;
    out.x = rid;
    out.y = (uint32_t)(best_qb + d + 1);
    out.z = (uint32_t)best_qb | ((uint32_t)(best_qe - best_qb) << 16);
    out.w = (uint32_t)(L - best_qe) | ((uint32_t)(best_sc * 2 + s) << 16);
    return true;
}

// block-wide exclusive scan over TAIL_THREADS values; *total receives the sum (valid for every thread)
__device__ __forceinline__ uint32_t tail_block_scan(uint32_t v, uint32_t *tmp /* 33 words */, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = tmp[lane], si = s;                      // TAIL_THREADS / 32 == 32 warps
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, si, o); if (lane >= o) si += t; }
        tmp[lane] = si - s;
        if (lane == 31) tmp[32] = si;
    }
    __syncthreads();
    const uint32_t r = tmp[warp] + inc - v;
    *total = tmp[32];
    __syncthreads();
    return r;
}

static const int TAIL_QCAP = 4096;                      // candidates a CTA queues for the extension per round (fewer if shared memory is short)

// The neighbour-tested membership of tail_verify without the table walk: does some sample of the read (a) sit in
// the anchor's k'-mer filter and (b) have the k'-mer H bases to its left or right in it too?  Every read with a
// >= K-base exact match passes (af_neighbour_ok); chance k'-mer hits pass with probability ~0.3 %.  Shared memory
// only -- what passes (4.5 % of the flagged reads, 85 % of them truly seeded) is queued for the extension, whose
// exact table lookups decide.
template <int KP>
__device__ __forceinline__ bool tail_prefilter(const uint32_t *sw, int VT, int L, const uint32_t *nm, const uint32_t *filt,
                                               uint32_t fmul, uint32_t nb, int K) {
    constexpr int S = 20 - KP;
    constexpr uint32_t kpmask = (1u << (2 * KP)) - 1u;
    const int np = L >= KP ? (L - KP) / S + 1 : 0;
    unsigned long long hit = 0;
#pragma unroll 6
    for (int j = 0; j < np; j++) {                       // which samples pass the shared-memory filter
        const int o = 2 * j * S, wi = o >> 5;
        const uint32_t key = __funnelshift_r(sw[wi * VT], sw[(wi + 1) * VT], o & 31) & kpmask;
        uint32_t b, fp3;
        af_filter_hash(key, fmul, nb, b, fp3);
        if (af_filter_test(filt[b], fp3)) hit |= 1ull << j;
    }
    const int H = (K - KP + 1) >> 1;
    while (hit) {
        const int j = __ffsll((long long)hit) - 1;
        hit &= hit - 1;
        const int p = j * S;
        if (nm) {                                        // a k'-mer that overlaps an N is no seed material
            bool n = false;
            for (int t = 0; t < KP; t++) n |= (nm[(p + t) >> 5] >> ((p + t) & 31)) & 1u;
            if (n) continue;
        }
        if (p >= H) {
            const int o2 = 2 * (p - H);
            const uint32_t k2 = __funnelshift_r(sw[(o2 >> 5) * VT], sw[((o2 >> 5) + 1) * VT], o2 & 31) & kpmask;
            uint32_t b2, f2;
            af_filter_hash(k2, fmul, nb, b2, f2);
            if (af_filter_test(filt[b2], f2)) return true;
        }
        if (p + H + KP <= L) {
            const int o2 = 2 * (p + H);
            const uint32_t k2 = __funnelshift_r(sw[(o2 >> 5) * VT], sw[((o2 >> 5) + 1) * VT], o2 & 31) & kpmask;
            uint32_t b2, f2;
            af_filter_hash(k2, fmul, nb, b2, f2);
            if (af_filter_test(filt[b2], f2)) return true;
        }
    }
    return false;
}

template <int KP>
__global__ void __launch_bounds__(TAIL_THREADS, 1)
k_tail(const af_tail_args a, const af_sink sink, const int has_sink) {
    extern __shared__ __align__(128) uint32_t tsm[];
    __shared__ uint32_t s_next, s_base, s_seeded, s_qn, s_tmp[33];
    constexpr int VT = TAIL_THREADS;
    uint32_t *filt = tsm;                                  // nb words: half-size anchor filter
    uint32_t *swb = filt + a.nb;                           // (W + 3) x VT: word t of thread i's candidate at swb[t * VT + i]
    uint32_t *bitmap = swb + (a.W + 3) * VT;               // REG_WORDS: bit (read_id - first read_id of the region)
    uint32_t *pref = bitmap + REG_WORDS;                   // PREF_WORDS
    uint32_t *queue = pref + PREF_WORDS;                   // a.qcap record indices
    uint32_t *s_apk = queue + a.qcap;                   // 2 x apk_words: padded packed anchor, both strands (if it fits)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = a.n_regions;
    if (blockIdx.x >= R) return;
    stage_filter(filt, a.g_filter, a.nb);
    if (a.anchor_in_smem)
        for (int i = tid; i < 2 * a.apk_words; i += VT) s_apk[i] = i < a.apk_words ? a.apk0p[i] : a.apk1p[i - a.apk_words];
    const uint32_t *apk0p = a.anchor_in_smem ? s_apk : a.apk0p, *apk1p = a.anchor_in_smem ? s_apk + a.apk_words : a.apk1p;
    if (tid == 0) s_seeded = 0;
    uint32_t *sw = swb + tid;
    uint32_t n_seeded = 0;
    bool over = false;
    const unsigned long long tail0 = has_sink ? sink.state[sink.seq & 1u] : 0ull;
    const uint4 *recs4 = reinterpret_cast<const uint4 *>(a.recs);
    const long long c_start = clock64();                    // a.dbg: phase boundaries of the CTA's first region, thread 0
    long long c_staged = 0, c_p1a = 0, c_p1b = 0, c_look = 0;

    // the candidate's W words (mate read_id & 1 of the pair the record holds) -> this thread's column, three zero words behind them
    auto load_words = [&](const uint4 *rp, uint32_t rid) {
        const int wofs = (int)(rid & 1u) * a.W, q0 = wofs >> 2, q1 = (wofs + a.W - 1) >> 2;
        for (int q = q0; q <= q1; q++) {
            const uint4 v = rp[1 + q];
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int t = 4 * q + e - wofs;
                if (t >= 0 && t < a.W) sw[t * VT] = vv[e];
            }
        }
        sw[a.W * VT] = 0; sw[(a.W + 1) * VT] = 0; sw[(a.W + 2) * VT] = 0;
    };
    auto nmask_of = [&](uint32_t rid) -> const uint32_t * {
        if (a.n_nreads <= 0) return nullptr;
        int lo = 0, hi = a.n_nreads;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (a.nread_ids[mid] < rid) lo = mid + 1; else hi = mid; }
        return lo < a.n_nreads && a.nread_ids[lo] == rid ? a.nmask + (size_t)lo * AF_NMASK_WORDS : nullptr;
    };

    __syncthreads();
    c_staged = clock64();
    for (uint32_t r = blockIdx.x; r < R; r += gridDim.x) {
        const long long t0 = a.n_tiles * (long long)r / R, t1 = a.n_tiles * (long long)(r + 1) / R;
        const uint32_t rid0 = (uint32_t)(t0 * 64);
        const int nwords = (int)(t1 - t0) * 2;
        for (int i = tid; i < min(nwords + 4, REG_WORDS); i += VT) bitmap[i] = 0;
        if (tid == 0) { s_next = 0; s_qn = 0; }
        const uint32_t ndir = min(a.dir_count[r], AF_DIR_CAP);
        const uint32_t *dir = a.dir + (size_t)r * AF_DIR_CAP;
        __syncthreads();
        // ---- phase 1, in rounds of at most qcap / 32 chunks so that the queue always has room ---------
        for (uint32_t k0 = 0; k0 < ndir; k0 += a.qcap / AF_CHUNK) {
            const uint32_t kend = min(k0 + a.qcap / AF_CHUNK, ndir);
            // 1a: a lane = a candidate; shared-memory pre-filter; survivors are queued
            for (;;) {
                uint32_t k = 0;
                if (lane == 0) k = k0 + atomicAdd(&s_next, 1u);
                k = __shfl_sync(FULL, k, 0);
                if (k >= kend) break;
                const uint32_t chunk = dir[k], ri = chunk * AF_CHUNK + lane;
                if (lane == 0) a.chunk_hits[chunk] = 0;
                const uint4 *rp = recs4 + (size_t)ri * a.rq;
                const uint32_t rid = rp[0].x;
                bool seeded = false;
                if (rid != AF_REC_INVALID) {
                    load_words(rp, rid);
                    seeded = tail_prefilter<KP>(sw, VT, a.uniform_len > 0 ? a.uniform_len : (int)a.lens[rid], nmask_of(rid), filt, a.fmul, a.nb, a.P.k);
                }
                __syncwarp();
                const uint32_t todo = __ballot_sync(FULL, seeded);
                if (todo) {
                    uint32_t pos = 0;
                    if (lane == 0) pos = atomicAdd(&s_qn, (uint32_t)__popc(todo));
                    pos = __shfl_sync(FULL, pos, 0);
                    if (seeded) queue[pos + __popc(todo & ((1u << lane) - 1u))] = ri;
                }
                __syncwarp();
            }
            __syncthreads();
            if (r == blockIdx.x && k0 == 0) c_p1a = clock64();
            // 1b: the queued candidates, one THREAD each, dense lanes
            const uint32_t qn = s_qn;
            n_seeded += tid == 0 ? qn : 0u;
            for (uint32_t i0 = (uint32_t)warp * 32; i0 < qn; i0 += VT) {   // whole warps: thread_extend's loops are warp-uniform
                const uint32_t i = i0 + lane;
                const bool have = i < qn && (!AF_TAIL_EXP || lane == 0);
                uint32_t ri = 0, rid = 0;
                int L = 0;
                const uint32_t *nm = nullptr;
                if (have) {
                    ri = queue[i];
                    const uint4 *rp = recs4 + (size_t)ri * a.rq;
                    rid = rp[0].x;
                    load_words(rp, rid);
                    L = a.uniform_len > 0 ? a.uniform_len : (int)a.lens[rid];
                    nm = nmask_of(rid);
                }
                __syncwarp();
                uint4 out;
                ExtProf pf = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                const long long tq0 = AF_TAIL_PROF ? clock64() : 0;
                const bool anchored = thread_extend(have, rid, L, sw, VT, nm, a.table, a.tmask, apk0p, apk1p, a.apn0p, a.apn1p, a.G, KP, 20 - KP, a.P, out, pf);
                if (AF_TAIL_PROF && a.dbg && have && i < 64 && r == blockIdx.x) {
                    long long *o = a.dbg + 8 * 148 + ((size_t)blockIdx.x * 64 + i) * 8;
                    o[0] = clock64() - tq0; o[1] = pf.collect; o[2] = pf.eval; o[3] = pf.n_eval; o[4] = pf.n_load; o[5] = pf.n_iter; o[6] = pf.t_start; o[7] = pf.t_consume;
                }
                if (have && anchored) {
                    // the anchored read's record replaces the header quad of its own candidate record; its bit goes
                    // into the region's hit bitmap and into its chunk's hit mask
                    reinterpret_cast<uint4 *>(a.recs)[(size_t)ri * a.rq] = out;
                    const uint32_t bit = rid - rid0;
                    atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
                    atomicOr(&a.chunk_hits[ri / AF_CHUNK], 1u << (ri % AF_CHUNK));
                }
            }
            __syncthreads();
            if (tid == 0) { s_next = 0; s_qn = 0; }
            __syncthreads();
        }
        if (r == blockIdx.x) c_p1b = clock64();
        // ---- phase 2: rank of every hit, base of the region ----------------------------------
        uint32_t c[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int g = 2 * tid + u;                       // group of 4 bitmap words; 2 * VT groups == REG_WORDS / 4
            const uint4 b = reinterpret_cast<const uint4 *>(bitmap)[g];
            c[u] = 4 * g < nwords ? __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w) : 0u;
        }
        uint32_t H;
        const uint32_t ex = tail_block_scan(c[0] + c[1], s_tmp, &H);
        pref[2 * tid] = ex; pref[2 * tid + 1] = ex + c[0];
        if (warp == 0) {
            if (lane == 0) st_relaxed_gpu(&a.region_state[r], H | 0x80000000u);
            uint32_t sum = 0;
            for (uint32_t i = lane; i < r; i += 32) {
                uint32_t v;
                do { v = ld_relaxed_gpu(&a.region_state[i]); } while (!(v & 0x80000000u));
                sum += v & 0x7FFFFFFFu;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
            if (lane == 0) s_base = sum;
        }
        __syncthreads();
        if (r == blockIdx.x) c_look = clock64();
        const uint32_t base = s_base;
        // ---- phase 3: placement -----------------------------------------------------------------
        for (uint32_t k = warp; k < ndir; k += VT / 32) {
            const uint32_t chunk = dir[k], hm = a.chunk_hits[chunk];
            if ((hm >> lane) & 1u) {
                const uint4 rec = recs4[(size_t)(chunk * AF_CHUNK + lane) * a.rq];
                const uint32_t bit = rec.x - rid0, wi = bit >> 5;
                uint32_t rank = pref[wi >> 2] + __popc(bitmap[wi] & ((1u << (bit & 31)) - 1u));
                for (uint32_t j = wi & ~3u; j < wi; j++) rank += __popc(bitmap[j]);
                const uint32_t at = base + rank;
                if (at < a.hits_cap) a.hits[at] = rec; else over = true;
            }
        }
        if (has_sink) {
            // push the region's records (now contiguous in hits[base, base + H)) to every rank's log:
            // consecutive threads store consecutive records -- full-size NVLink write packets
            __syncthreads();
            const uint32_t end = min(base + H, a.hits_cap);
            for (uint32_t j = base + tid; j < end; j += VT) {
                const unsigned long long at = tail0 + 1 + j;
                if (at >= sink.log_cap) break;
                const uint4 rec = a.hits[j];
                for (int q = 0; q < sink.world; q++) ((uint4 *)(sink.region[q] + AF_LOG_HEADER_BYTES))[at] = rec;
            }
        }
        if (r == R - 1 && tid == 0) {                        // the batch's totals are known here
            const uint32_t total = base + H;
            a.counts[AF_CNT_HITS] = total;
            if (total > a.hits_cap) atomicOr(&a.counts[AF_CNT_STATUS], AF_STATUS_HIT_OVERFLOW);
            if (a.counts[AF_CNT_FLAGGED] > a.cand_cap) atomicOr(&a.counts[AF_CNT_STATUS], AF_STATUS_CAND_OVERFLOW);
            if (has_sink) {
                const unsigned long long meta = sink.state[2 + (sink.seq & 1u)];
                uint32_t status = (uint32_t)meta;
                const uint32_t batches = (uint32_t)(meta >> 32) + 1;
                unsigned long long tail = tail0;
                if (tail0 < sink.log_cap) {
                    const uint4 marker = make_uint4(AF_LOG_MARKER, (uint32_t)sink.pair_base, (uint32_t)(sink.pair_base >> 32), total);
                    for (int q = 0; q < sink.world; q++) ((uint4 *)(sink.region[q] + AF_LOG_HEADER_BYTES))[tail0] = marker;
                    tail = tail0 + 1 + total;
                }
                if (tail0 >= sink.log_cap || tail > sink.log_cap || total > a.hits_cap) {
                    if (tail > sink.log_cap) tail = sink.log_cap;
                    status |= AF_STATUS_LOG_OVERFLOW;
                    atomicOr(&a.counts[AF_CNT_STATUS], AF_STATUS_LOG_OVERFLOW);
                }
                for (int q = 0; q < sink.world; q++) {
                    af_log_header *h = (af_log_header *)sink.region[q];
                    h->status = status; h->n_batches = batches; h->tail = tail;
                }
                sink.state[(sink.seq & 1u) ^ 1u] = tail;
                sink.state[2 + ((sink.seq & 1u) ^ 1u)] = ((unsigned long long)batches << 32) | status;
            }
        }
        __syncthreads();                                     // the bitmap is zeroed for the next region
    }
    if (a.dbg && tid == 0) {
        long long *o = a.dbg + (size_t)blockIdx.x * 8;
        o[0] = c_staged - c_start; o[1] = c_p1a - c_start; o[2] = c_p1b - c_start; o[3] = c_look - c_start; o[4] = clock64() - c_start;
        o[5] = n_seeded;
    }
    if (tid == 0 && n_seeded) atomicAdd(&s_seeded, n_seeded);
    if (over) atomicOr(&a.counts[AF_CNT_STATUS], AF_STATUS_HIT_OVERFLOW);
    __syncthreads();
    if (tid == 0 && s_seeded) atomicAdd(&a.counts[AF_CNT_SEEDED], s_seeded);
}

static size_t tail_smem_bytes(const af_dev_index *d, int W, int qcap, bool anchor_in_smem) {
    return ((size_t)d->nb2 + (size_t)(W + 3) * TAIL_THREADS + REG_WORDS + PREF_WORDS + qcap + (anchor_in_smem ? 2 * (size_t)d->apk_words : 0)) * 4;
}

int af_tail_launch(const af_dev_index *d, af_tail_args &a, const af_sink *sink, cudaStream_t st) {
    static size_t max_dyn[64][2] = {{0}};                   // per device, per k'
    const int ki = d->kp == 12 ? 0 : 1;
    if (!max_dyn[d->device & 63][ki]) {
        const int rc = d->kp == 12 ? allow_full_smem(k_tail<12>, &max_dyn[d->device & 63][ki]) : allow_full_smem(k_tail<13>, &max_dyn[d->device & 63][ki]);
        if (rc) return rc;
    }
    const size_t avail = max_dyn[d->device & 63][ki];
    a.qcap = TAIL_QCAP;
    while (a.qcap > 256 && tail_smem_bytes(d, a.W, a.qcap, false) > avail) a.qcap /= 2;
    a.anchor_in_smem = tail_smem_bytes(d, a.W, a.qcap, true) <= avail;
    const size_t smem = tail_smem_bytes(d, a.W, a.qcap, a.anchor_in_smem);
    if (smem > avail) { af_set_error("k_tail: %zu bytes of shared memory needed, %zu available", smem, avail); return AF_ERR_ARG; }
    const unsigned grid = a.n_regions < (uint32_t)d->num_sms ? a.n_regions : (unsigned)d->num_sms;
    const af_sink none = af_sink();
    if (d->kp == 12) k_tail<12><<<grid, TAIL_THREADS, smem, st>>>(a, sink ? *sink : none, sink != nullptr);
    else k_tail<13><<<grid, TAIL_THREADS, smem, st>>>(a, sink ? *sink : none, sink != nullptr);
    af_note_launches(1);
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}
