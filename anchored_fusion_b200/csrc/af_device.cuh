// af_device.cuh -- device helpers shared by the kernels of the anchoring path (af_kernels.cu, af_tail.cu):
// streaming / gather loads, filter staging through the TMA engine, the read view used by the slow (N)
// paths, and the warp-cooperative diagonal evaluation + X-drop extension ("Anchoring spec v1").
#pragma once
#include <cuda_runtime.h>

#include "af_common.h"

#define FULL 0xFFFFFFFFu

#define AF_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            af_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return AF_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

struct af_dev_index {
    int device;
    af_params_t P;
    int32_t kp, stride, G;
    uint32_t fmul, nb, tmask;
    uint32_t *d_filter;  // nb words
    uint32_t fmul2, nb2;
    uint32_t *d_filter2; // nb2 words: the half-size filter k_verify stages
    uint2 *d_table;      // tmask+1 entries {key, value}
    uint32_t *d_member;  // 4^kp-bit exact membership bitmap (L2 resident)
    uint8_t *d_anchor;   // G base codes
    uint32_t *d_apk[2];  // 2-bit packed anchor: forward / reverse complement
    uint32_t *d_apkp[2]; // the same behind 16 zero words (256 bases), so that a window may start left of the anchor (k_tail)
    int apk_words;       // words of each d_apkp array
    uint32_t *d_apn[2];  // anchor with N only: 1 bit per base that is N (forward / reverse strand) behind 8 zero words; else nullptr
    int anchor_has_n;
    int pad_byte;
    int num_sms;
    bool saturated;      // many filter buckets overflowed (long anchor): flagged reads take the exact k_verify route
    bool bloom;          // d_filter holds Bloom bits (af_bloom_probe) instead of fingerprint buckets
};

// ---- candidate stream: seed scan -> tail kernel (af_tail.cu) ---------------------------------
// The scan stores every flagged read -- read_id + its packed words, straight from the registers that hold
// them -- into 32-record chunks taken from a pool, so the stage that follows reads its candidates
// coalesced (and mostly from L2) instead of gathering 16-byte quads of 362 k scattered reads from HBM.
// The tile range of the batch is cut into R regions (<= AF_REG_TILES tiles each, scan CTA b produces
// regions [b*m, (b+1)*m)); a warp owns the chunk it is filling, every chunk belongs to one region and is
// listed in that region's directory.  Order is restored at the end by rank in a per-region hit bitmap.
static const int AF_REG_TILES = 4096;                     // tiles per region: 32 KB of hit bitmap in the tail CTA
static const int AF_CHUNK = 32;                           // records per chunk
static const uint32_t AF_DIR_CAP = AF_REG_TILES * 2 + 64; // chunks a region can own: all 64 reads of every tile + one open chunk per scan warp
static const uint32_t AF_REC_INVALID = 0xFFFFFFFFu;       // read_id of an unused record slot
struct af_emit {
    uint32_t *recs;        // pool_chunks x 32 records of rq = Q + 1 quads: {read_id, 0, 0, 0} + the pair's Q quads
    uint32_t *pool;        // [0] chunks handed out so far
    uint32_t pool_chunks;
    uint32_t *dir_count;   // [R] chunks listed per region
    uint32_t *dir;         // [R][AF_DIR_CAP] chunk ids
    uint32_t *counts;      // the batch's counters (AF_CNT_*)
    int32_t m;             // regions per scan CTA
};
AF_HD int af_rec_quads(int W) { return (2 * W + 3) / 4 + 1; }   // header quad + the pair's Q quads

// kernels launched by this library / per-stage CUDA-event spans (af_kernels.cu)
enum { ST_SCAN = 0, ST_COMPACT1 = 1, ST_VERIFY = 2, ST_EXTEND = 3, ST_COMPACT2 = 4, ST_N = 5 };
void af_note_launches(int n);
void af_prof_mark(cudaEvent_t *ev, cudaStream_t st);
void af_prof_span(cudaEvent_t a, cudaStream_t st, int stage);

// opt a kernel in to all the shared memory an SM offers a CTA (227 KB) minus what it declares statically;
// *max_dynamic receives the dynamic part it may then be launched with
template <class K>
static int allow_full_smem(K kernel, size_t *max_dynamic) {
    cudaFuncAttributes a;
    AF_CUDA(cudaFuncGetAttributes(&a, kernel));
    const size_t dyn = 227 * 1024 - a.sharedSizeBytes;
    AF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    if (max_dynamic) *max_dynamic = dyn;
    return AF_OK;
}

__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// One 16-byte quad of a scattered read (verify).  Without the hint L2 fills a whole 128-byte line per
// quad (ncu: 129 MB of DRAM reads for 35 MB of sectors asked for); with it 68 MB.
__device__ __forceinline__ uint4 ld_gather_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Stage the anchor filter into shared memory: 128-bit loads, several in flight per thread (a
// one-word-per-iteration loop spends ~20 us of pure L2 latency here; ncu, round 1).
// The same with the TMA engine: one thread posts bulk copies global -> shared (cp.async.bulk, SASS
// UBLKCP) against an mbarrier, every thread waits for the barrier's phase.  No registers, no LSU
// instructions, and the copy runs while the warps' first tile loads are in flight.  Ends with the
// filter visible to all threads of the CTA.  Measured: the scan takes the same 0.187 ms per 10 M pairs
// either way (the ~200 KB per CTA come from L2 in a few microseconds in both forms); kept because it
// leaves the load/store pipe and 16 registers per thread to the tile loads already in flight.
// AF_STAGE_TMA=0 builds the load/store loop instead.
#ifndef AF_STAGE_TMA
#define AF_STAGE_TMA 1
#endif
__device__ __forceinline__ void stage_filter_tma(uint32_t *filt, const uint32_t *__restrict__ g_filter, uint32_t nb) {
    __shared__ __align__(8) unsigned long long mbar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
    const uint32_t bytes = nb * 4u;                                // nb is a multiple of 32: 128-byte granules
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(filt);
        const uint32_t CH = 32u << 10;
        for (uint32_t off = 0; off < bytes; off += CH) {
            const uint32_t n = bytes - off < CH ? bytes - off : CH;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + off), "l"((const char *)g_filter + off), "r"(n), "r"(bar) : "memory");
        }
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
}

__device__ __forceinline__ void stage_filter_ldst(uint32_t *filt, const uint32_t *__restrict__ g_filter, uint32_t nb) {
    const uint4 *src = reinterpret_cast<const uint4 *>(g_filter);
    uint4 *dst = reinterpret_cast<uint4 *>(filt);
    const uint32_t n4 = nb >> 2, step = blockDim.x;              // nb is a multiple of 32
    uint32_t i = threadIdx.x;
    for (; i + 3 * step < n4; i += 4 * step) {
        const uint4 a = src[i], b = src[i + step], c = src[i + 2 * step], d = src[i + 3 * step];
        dst[i] = a; dst[i + step] = b; dst[i + 2 * step] = c; dst[i + 3 * step] = d;
    }
    for (; i < n4; i += step) dst[i] = src[i];
}

__device__ __forceinline__ void stage_filter(uint32_t *filt, const uint32_t *__restrict__ g_filter, uint32_t nb) {
#if AF_STAGE_TMA
    stage_filter_tma(filt, g_filter, nb);
#else
    stage_filter_ldst(filt, g_filter, nb);
#endif
}


struct ReadRef {
    const uint32_t *packed;   // tile-interleaved words
    size_t base;              // word index of word 0 of this read's pair in its quad 0
    int wofs;                 // mate * W
    const uint32_t *nm;       // N-mask words of this read or nullptr
    int L;
    __device__ __forceinline__ uint32_t word(int t) const {
        const int wi = wofs + t;
        return packed[base + (size_t)(wi >> 2) * 128 + (wi & 3)];
    }
    __device__ __forceinline__ uint32_t base_at(int i) const { return (word(i >> 4) >> (2 * (i & 15))) & 3u; }
    __device__ __forceinline__ bool is_n(int i) const { return nm && ((nm[i >> 5] >> (i & 31)) & 1u); }
};

__device__ __forceinline__ bool diag_match(const ReadRef &r, int s, int i, int d, const uint8_t *__restrict__ anchor, int G) {
    const int ap = i + d;
    if (i < 0 || i >= r.L || ap < 0 || ap >= G) return false;
    const int fi = s ? r.L - 1 - i : i;
    if (r.is_n(fi)) return false;
    uint32_t b = r.base_at(fi);
    if (s) b = 3u - b;
    return anchor[ap] == b;
}

// 32 bases (64 bits) of a 2-bit packed sequence starting at base `pos` (pos >= 0)
__device__ __forceinline__ unsigned long long packed_window(const uint32_t *__restrict__ a, int pos) {
    const int wi = pos >> 4, sh = 2 * (pos & 15);
    const uint32_t w0 = a[wi], w1 = a[wi + 1], w2 = a[wi + 2];
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}


struct ExtParams {
    int k, A, B, clip5, clip3, T, X;
};

#define NEG_INF (-(1 << 29))

// number of set mask bits in positions [0, x), x in [0, 256]; lane t holds word t (mw) and the
// count of set bits in words < t (cp); lanes >= 8 hold mw = 0, cp = total.
__device__ __forceinline__ int mask_cum(uint32_t mw, uint32_t cp, int x) {
    int wi = x >> 5;
    uint32_t wv = __shfl_sync(FULL, mw, wi), cv = __shfl_sync(FULL, cp, wi);
    return (int)cv + __popc(wv & ((1u << (x & 31)) - 1u));
}

// One direction of the ungapped X-drop extension over mask positions start, start+dir, ...
// (n steps).  Lane = step within a 32-step chunk; scores come from popcounts of the match
// mask, the running maximum from a warp prefix-max scan.  Mirrors `extend` in the oracle.
__device__ __forceinline__ void extend_dir(uint32_t mw, uint32_t cp, int start, int dir, int n, int qlen, int h0,
                                           const ExtParams &P, int lane, int &mx_out, int &off_out, int &g_out) {
    int mx = h0, off = 0, g = -1;
    const int base_cum = dir > 0 ? mask_cum(mw, cp, start) : mask_cum(mw, cp, start + 1);
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < n;
        const int jj = valid ? j : 0, pos = start + dir * jj;
        const int c = mask_cum(mw, cp, dir > 0 ? pos + 1 : pos);
        const int ones = dir > 0 ? c - base_cum : base_cum - c;
        int Pj = h0 + P.A * (jj + 1) - (P.A + P.B) * ((jj + 1) - ones);
        if (!valid) Pj = NEG_INF;
        const bool dead = valid && Pj <= 0;
        int pm = Pj;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, pm, o); if (lane >= o) pm = max(pm, t); }
        const int Mj = max(pm, mx);
        const bool xd = valid && (Mj - Pj > P.X);
        const uint32_t bd = __ballot_sync(FULL, dead), bx = __ballot_sync(FULL, xd);
        int endlane = min(32, n - j0);
        bool stop = false;
        if (bd) { endlane = min(endlane, __ffs(bd) - 1); stop = true; }
        if (bx) { endlane = min(endlane, __ffs(bx)); stop = true; }   // the x-drop step itself is processed
        const bool processed = lane < endlane;
        int cm = processed ? Pj : NEG_INF;
#pragma unroll
        for (int o = 16; o; o >>= 1) cm = max(cm, __shfl_xor_sync(FULL, cm, o));
        if (cm > mx) {
            const uint32_t be = __ballot_sync(FULL, processed && Pj == cm);
            mx = cm;
            off = j0 + __ffs(be);
        }
        if (n == qlen) {
            const int gl = qlen - 1 - j0;
            if (gl >= 0 && gl < endlane) g = __shfl_sync(FULL, Pj, gl);
        }
        if (stop) break;
    }
    mx_out = mx; off_out = off; g_out = g;
}

// Score of diagonal d given its 256-bit match mask (lane c < 8 holds mask word c = oriented read positions
// [32c, 32c+32); lanes >= 8 hold 0): leftmost run of k matches, then left and right X-drop extension.
// Returns the score or -1 if the mask holds no run of k.
__device__ __forceinline__ int eval_from_mask(uint32_t mw, int d, int L, int G, const ExtParams &P, int lane,
                                              int &qb_out, int &qe_out) {
    uint32_t pc = __popc(mw), cp = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, cp, o); if (lane >= o) cp += t; }
    cp -= pc;  // exclusive: set bits in words before this lane's word

    // leftmost run of k <= 32 matches, all eight words at once: lane c looks at mask bits [32c, 32c+64) and marks
    // the starts b < 32 with bits b .. b+k-1 set (AND of shifted copies, run lengths doubled: k = sum of powers of 2)
    const uint32_t nx = __shfl_down_sync(FULL, mw, 1);
    const unsigned long long v = (unsigned long long)mw | ((unsigned long long)(lane < 31 ? nx : 0u) << 32);
    unsigned long long acc = ~0ull, p = v;
    for (int kk = P.k, off = 0, len = 1; kk; kk >>= 1, len <<= 1) {
        if (kk & 1) { acc &= p >> off; off += len; }
        p &= p >> len;
    }
    const uint32_t starts = (uint32_t)acc;
    const uint32_t who = __ballot_sync(FULL, starts != 0);
    if (!who) return -1;
    const int c0 = __ffs(who) - 1;
    const int qb0 = c0 * 32 + __ffs(__shfl_sync(FULL, starts, c0)) - 1;
    if (qb0 + P.k > L) return -1;

    int sc = P.k * P.A, qb = 0, qe = L, mx, off, g;
    if (qb0 > 0) {
        const int n = min(qb0, qb0 + d);
        extend_dir(mw, cp, qb0 - 1, -1, n, qb0, sc, P, lane, mx, off, g);
        if (g <= 0 || g <= mx - P.clip5) { qb = qb0 - off; sc = mx; } else { qb = 0; sc = g; }
    }
    const int qe0 = qb0 + P.k;
    if (qe0 < L) {
        const int n = min(L - qe0, G - (qe0 + d));
        extend_dir(mw, cp, qe0, +1, n, L - qe0, sc, P, lane, mx, off, g);
        if (g <= 0 || g <= mx - P.clip3) { qe = qe0 + off; sc = mx; } else { qe = L; sc = g; }
    }
    qb_out = qb; qe_out = qe;
    return sc;
}

// Evaluate diagonal (s, d) of the read held by the warp.  Returns the score or -1 if the
// diagonal holds no run of k matches.  rw: lane t < W holds packed word t of the read;
// nw: lane t < 8 holds N-mask word t.
__device__ __forceinline__ int eval_diag(int s, int d, int L, uint32_t rw, uint32_t nw, bool has_n,
                                         const uint8_t *__restrict__ anchor, int G, const ExtParams &P, int lane,
                                         int &qb_out, int &qe_out) {
    // 256-bit match mask, 32 positions per ballot
    uint32_t mw = 0;
    const int nchunks = (L + 31) >> 5;
    for (int c = 0; c < nchunks; c++) {
        const int i = c * 32 + lane;
        const int fi = min(max(s ? L - 1 - i : i, 0), AF_MAX_READ_LEN - 1);   // position in the stored read
        uint32_t word = __shfl_sync(FULL, rw, fi >> 4);
        uint32_t base = (word >> (2 * (fi & 15))) & 3u;
        if (s) base = 3u - base;
        bool isn = false;
        if (has_n) { uint32_t nword = __shfl_sync(FULL, nw, fi >> 5); isn = (nword >> (fi & 31)) & 1u; }
        const int ap = i + d;
        bool m = false;
        if (i < L && ap >= 0 && ap < G && !isn) m = anchor[ap] == base;
        const uint32_t bal = __ballot_sync(FULL, m);
        if (lane == c) mw = bal;
    }
    return eval_from_mask(mw, d, L, G, P, lane, qb_out, qe_out);
}

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

// ---- extension, one warp per read, walking MISMATCHES instead of bases -------------------------------
// k_extend steps through a diagonal 32 bases at a time with warp prefix scans (~120 warp instructions per 32
// bases); with ~120 reads to extend per SM and region that instruction stream, not latency, was the limit
// (measured: 11 k cycles per read at 8 warps per scheduler).  Here
//   * the 256-bit match mask is 8 XORs of 64-bit windows (lane c < 8: read words against the 2-bit packed anchor,
//     forward or reverse complement) instead of 256 base tests;
//   * the leftmost run of k matches comes from AND-ing shifted copies of the mask words (eval_from_mask's search);
//   * the X-drop extension walks the mismatches of the mask: between two mismatches the score only rises, so the
//     maximum moves only at the end of a match run and every stop condition (score <= 0, max - score > X) can only
//     fire on a mismatch.  A 150-base read with two substitutions takes three steps of ~12 instructions.
// A one-thread-per-read version of the same was tried and rejected: ~3 000 dependent instructions per read at
// ~8 cycles each, 90 k cycles for the phase (tools/tail_timing.py, AF_TAIL_PROF).
// Same results as oracle/af_oracle.c::extend / diag_eval, step for step (spec v1).

// The 256-bit match mask of diagonal (s, d), word-parallel; lane c < 8 returns mask word c (oriented read
// positions [32c, 32c+32)), lanes >= 8 return 0.  apn*: the anchor's N bitmasks or nullptr.
__device__ __forceinline__ uint32_t diag_mask_wp(int s, int d, int L, uint32_t rw, uint32_t nwv, bool has_n,
                                                const uint32_t *apk0p, const uint32_t *apk1p, const uint32_t *apn0p,
                                                const uint32_t *apn1p, int G, int lane) {
    const int c = lane & 7;
    const int dd = s ? G - L - d : d;                       // forward read base j lies on strand-s anchor base j + dd
    const uint32_t *ap = s ? apk1p : apk0p, *an = s ? apn1p : apn0p;
    const uint32_t r0 = __shfl_sync(FULL, rw, 2 * c), r1 = __shfl_sync(FULL, rw, 2 * c + 1);
    const int lo = max(0, -dd), hi = min(L, G - dd);        // read bases that face an anchor base: [lo, hi)
    const int b0 = max(lo - 32 * c, 0), b1 = min(hi - 32 * c, 32);
    uint32_t mf = 0;
    if (b1 > b0) {                                          // (then 32c + dd lies in (-32, G): inside the padded arrays)
        const unsigned long long x = ((unsigned long long)r0 | ((unsigned long long)r1 << 32)) ^ apk_window(ap, 32 * c + dd);
        const unsigned long long ne = x | (x >> 1);         // even bits: 1 = bases differ
        const uint32_t eq = ~(even_bits((uint32_t)ne) | (even_bits((uint32_t)(ne >> 32)) << 16));
        mf = eq & ((b1 - b0 >= 32 ? FULL : ((1u << (b1 - b0)) - 1u)) << b0);
        if (an) {                                           // the packed anchor holds A where the anchor has N
            const int pn = 32 * c + dd + 32 * APN_PAD_WORDS;
            mf &= ~__funnelshift_r(an[pn >> 5], an[(pn >> 5) + 1], pn & 31);
        }
    }
    if (has_n) mf &= ~__shfl_sync(FULL, nwv, c);
    if (!s) return lane < 8 ? mf : 0u;
    // oriented position i = L - 1 - j: mask(i) = Rev(i + 256 - L) with Rev = the 256-bit mask bit-reversed;
    // Rev word t = brev(forward word 7 - t), which lane 7 - t holds
    const uint32_t rev = __brev(mf);
    const int sh = 256 - L, w0 = c + (sh >> 5);
    const uint32_t lo_w = __shfl_sync(FULL, rev, (7 - w0) & 7), hi_w = __shfl_sync(FULL, rev, (6 - w0) & 7);
    const uint32_t m = __funnelshift_r(w0 < 8 ? lo_w : 0u, w0 + 1 < 8 ? hi_w : 0u, sh & 31);
    return lane < 8 ? m : 0u;
}

// first position in [pos, limit) whose bit is 0 in the mask spread over lanes 0..7 (lanes >= 8 hold 0), or limit
__device__ __forceinline__ int dm_next_zero(uint32_t mw, int pos, int limit, int lane) {
    const int wp = pos >> 5;
    uint32_t z = ~mw;
    if (lane < wp) z = 0u; else if (lane == wp) z &= FULL << (pos & 31);
    const uint32_t who = __ballot_sync(FULL, z != 0);       // never empty: lanes >= 8 hold all-zero words
    const int c = __ffs(who) - 1;
    return min(c * 32 + __ffs(__shfl_sync(FULL, z, c)) - 1, limit);
}
// last position in (limit, pos] whose bit is 0, or limit (limit >= -1, pos <= 255)
__device__ __forceinline__ int dm_prev_zero(uint32_t mw, int pos, int limit, int lane) {
    if (pos <= limit) return limit;
    const int wp = pos >> 5;
    uint32_t z = ~mw;
    if (lane > wp) z = 0u; else if (lane == wp) z &= FULL >> (31 - (pos & 31));
    const uint32_t who = __ballot_sync(FULL, z != 0);
    if (!who) return limit;
    const int c = 31 - __clz(who);
    return max(c * 32 + 31 - __clz(__shfl_sync(FULL, z, c)), limit);
}

// One direction of the ungapped X-drop extension over mask positions start, start + dir, ... (n steps), from
// score h0 (> 0); qlen = read bases left on this side.  oracle/af_oracle.c::extend, mismatch by mismatch;
// warp-uniform.
__device__ __forceinline__ void walk_dir(uint32_t mw, int start, int dir, int n, int qlen, int h0, const ExtParams &P,
                                         int lane, int &mx_out, int &off_out, int &g_out) {
    int cur = h0, mx = h0, off = 0, g = -1, j = 0;
    for (;;) {
        const int jz = dir > 0 ? dm_next_zero(mw, start + j, start + n, lane) - start
                               : start - dm_prev_zero(mw, start - j, start - n, lane);   // step of the next mismatch, n if none
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

// Score of diagonal d given its match mask (lane c < 8 holds word c, lanes >= 8 hold 0): eval_from_mask with
// walk_dir in place of extend_dir.  Returns the score or -1 if the mask holds no run of k.
__device__ __forceinline__ int eval_mask_walk(uint32_t mw, int d, int L, int G, const ExtParams &P, int lane, int &qb_out, int &qe_out) {
    const uint32_t nx = __shfl_down_sync(FULL, mw, 1);
    const unsigned long long v = (unsigned long long)mw | ((unsigned long long)(lane < 31 ? nx : 0u) << 32);
    unsigned long long acc = ~0ull, p = v;
    for (int kk = P.k, off = 0, len = 1; kk; kk >>= 1, len <<= 1) {
        if (kk & 1) { acc &= p >> off; off += len; }
        p &= p >> len;
    }
    const uint32_t starts = (uint32_t)acc;
    const uint32_t who = __ballot_sync(FULL, starts != 0);
    if (!who) return -1;
    const int c0 = __ffs(who) - 1;
    const int qb0 = c0 * 32 + __ffs(__shfl_sync(FULL, starts, c0)) - 1;
    if (qb0 + P.k > L) return -1;
    int sc = P.k * P.A, qb = 0, qe = L, mx, off, g;
    if (qb0 > 0) {
        const int n = min(qb0, qb0 + d);
        walk_dir(mw, qb0 - 1, -1, n, qb0, sc, P, lane, mx, off, g);
        if (g <= 0 || g <= mx - P.clip5) { qb = qb0 - off; sc = mx; } else { qb = 0; sc = g; }
    }
    const int qe0 = qb0 + P.k;
    if (qe0 < L) {
        const int n = min(L - qe0, G - (qe0 + d));
        walk_dir(mw, qe0, +1, n, L - qe0, sc, P, lane, mx, off, g);
        if (g <= 0 || g <= mx - P.clip3) { qe = qe0 + off; sc = mx; } else { qe = L; sc = g; }
    }
    qb_out = qb; qe_out = qe;
    return sc;
}

// ---- k_tail (af_tail.cu): verify + extend + ordered placement of the candidate stream, one kernel ----
struct af_tail_args {
    // candidate stream
    uint32_t *recs; int rq;                 // records of rq quads
    const uint32_t *dir_count, *dir;        // per-region chunk directories
    uint32_t n_regions; long long n_tiles;
    uint32_t *chunk_hits, *region_state;    // per chunk: which of its records are anchored; per region: hit count | ready bit
    // batch (slow paths: reads with N, ragged lengths)
    const uint32_t *packed; int W, Q, uniform_len; const uint16_t *lens;
    const uint32_t *nread_ids, *nmask; int n_nreads;
    // anchor index
    const uint32_t *g_filter; uint32_t fmul, nb; const uint2 *table; uint32_t tmask;
    const uint8_t *anchor; const uint32_t *apk0p, *apk1p, *apn0p, *apn1p; int apk_words, anchor_has_n, G, anchor_in_smem, qcap;   // apk*p: padded packed anchor
    ExtParams P;
    // results
    uint4 *hits; uint32_t hits_cap, cand_cap; uint32_t *counts;
    long long *dbg;                         // debug: per CTA, 8 int64 of phase boundaries (cycles), or nullptr
};
int af_tail_launch(const af_dev_index *d, af_tail_args &a, const af_sink *sink, cudaStream_t st);
