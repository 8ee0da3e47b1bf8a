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
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The whole warp extends one candidate (k_extend's lookup + de-duplication, the evaluation above).  rw: lane
// t < W holds packed word t; nwv: lane t < 8 holds N-mask word t.  Returns true (warp-uniform) and the record if
// the best diagonal scores >= T.
__device__ __forceinline__ bool tail_extend(uint32_t rid, int L, uint32_t rw, uint32_t nwv, bool has_n,
                                            const uint2 *__restrict__ table, uint32_t tmask, const uint32_t *apk0p,
                                            const uint32_t *apk1p, const uint32_t *apn0p, const uint32_t *apn1p, int G,
                                            int KP, int S, const ExtParams &P, int lane, uint4 &out) {
    const uint32_t kpmask = (1u << (2 * KP)) - 1u;
    int best_sc = -1, best_qb = 0, best_qe = 0;
    uint32_t best_key = 0xFFFFFFFFu;
    const int nprobe = af_nsamples(L, KP);
    for (int p0 = 0; p0 < nprobe; p0 += 32) {
        const int pi = p0 + lane, p = af_sample0(KP) + pi * S;
        bool active = pi < nprobe;
        const int o = 2 * (active ? p : 0), wi = o >> 5;
        uint32_t w0 = __shfl_sync(FULL, rw, wi), w1 = __shfl_sync(FULL, rw, min(wi + 1, 31));
        const uint32_t key = __funnelshift_r(w0, w1, o & 31) & kpmask;
        if (has_n) {
            const int q = active ? p : 0;
            uint32_t n0 = __shfl_sync(FULL, nwv, q >> 5), n1 = __shfl_sync(FULL, nwv, min((q >> 5) + 1, 31));
            if (__funnelshift_r(n0, n1, q & 31) & ((1u << KP) - 1u)) active = false;  // k'-mer overlaps an N
        }
        uint32_t slot = af_table_hash(key, tmask), val = 0;
        bool found = false;
        auto next_match = [&]() {                            // advance to this lane's next table entry with the same key
            found = false;
            while (active) {
                uint2 e = table[slot];
                slot = (slot + 1) & tmask;
                if (e.x == AF_T_EMPTY) { active = false; break; }
                if (e.x == key) { found = true; val = e.y; break; }
            }
        };
        next_match();
        uint32_t fm;
        while ((fm = __ballot_sync(FULL, found)) != 0) {
            uint32_t dkey = 0;
            bool leader = false;
            if (found) {
                const int s = val >> 31, j = (int)(val & 0x7FFFFFFFu);
                const int d = j - (s ? L - p - KP : p);
                dkey = ((uint32_t)s << 31) | (uint32_t)(d + 1024);
                uint32_t grp = __match_any_sync(fm, dkey);
                leader = (__ffs(grp) - 1) == lane;
            }
            uint32_t leaders = __ballot_sync(FULL, leader);
            while (leaders) {
                const int src = __ffs(leaders) - 1;
                leaders &= leaders - 1;
                const uint32_t dk = __shfl_sync(FULL, dkey, src);
                const int s = dk >> 31, d = (int)(dk & 0x7FFFFFFFu) - 1024;
                int qb, qe;
                const int sc = eval_mask_walk(diag_mask_wp(s, d, L, rw, nwv, has_n, apk0p, apk1p, apn0p, apn1p, G, lane), d, L, G, P, lane, qb, qe);
                if (sc > best_sc || (sc == best_sc && sc >= 0 && dk < best_key)) { best_sc = sc; best_qb = qb; best_qe = qe; best_key = dk; }
            }
            if (found) next_match();
        }
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
    const int np = af_nsamples(L, KP);
    unsigned long long hit = 0;
#pragma unroll 6
    for (int j = 0; j < np; j++) {                       // which samples pass the shared-memory filter
        const int o = 2 * (af_sample0(KP) + j * S), wi = o >> 5;
        const uint32_t key = __funnelshift_r(sw[wi * VT], sw[(wi + 1) * VT], o & 31) & kpmask;
        uint32_t b, fp3;
        af_filter_hash(key, fmul, nb, b, fp3);
        if (af_filter_test(filt[b], fp3)) hit |= 1ull << j;
    }
    const int H = (K - KP + 1) >> 1;
    while (hit) {
        const int j = __ffsll((long long)hit) - 1;
        hit &= hit - 1;
        const int p = af_sample0(KP) + j * S;
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
    __shared__ uint32_t s_next, s_done, s_base, s_seeded, s_qn, s_qhead, s_tmp[33];
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
        const uint32_t ndir = min(a.dir_count[r], AF_DIR_CAP);
        const uint32_t *dir = a.dir + (size_t)r * AF_DIR_CAP;
        __syncthreads();
        // ---- phase 1, in rounds of at most qcap / 32 chunks so that the queue always has room ---------
        // One work loop per round: a warp extends a queued candidate if there is one, else pre-filters the next
        // chunk (a lane = a candidate; survivors are queued), else waits for the last chunks to be finished by
        // their warps.  No barrier between filtering and extending, and the extension -- the part whose cost
        // varies from read to read -- is balanced over all warps from the first queued candidate on.
        for (uint32_t k0 = 0; k0 < ndir; k0 += a.qcap / AF_CHUNK) {
            const uint32_t nch = min(a.qcap / AF_CHUNK, ndir - k0);
            for (uint32_t i = tid; i < nch * AF_CHUNK; i += VT) queue[i] = AF_REC_INVALID;
            if (tid == 0) { s_next = 0; s_done = 0; s_qn = 0; s_qhead = 0; }
            __syncthreads();
            for (;;) {
                uint32_t what = 0, arg = 0;                  // 1 = extend queue entry `arg`, 2 = filter chunk `arg`, 3 = all done
                if (lane == 0) {
                    for (;;) {                               // take a queue entry if one is reserved and not yet taken
                        const uint32_t h = ld_vol(&s_qhead);
                        if (h >= ld_vol(&s_qn)) break;
                        if (atomicCAS(&s_qhead, h, h + 1) == h) { what = 1; arg = h; break; }
                    }
                    if (!what && ld_vol(&s_next) < nch) {
                        const uint32_t k = atomicAdd(&s_next, 1u);
                        if (k < nch) { what = 2; arg = k; }
                    }
                    if (!what && ld_vol(&s_done) >= nch && ld_vol(&s_qhead) >= ld_vol(&s_qn)) what = 3;
                }
                what = __shfl_sync(FULL, what, 0);
                arg = __shfl_sync(FULL, arg, 0);
                if (what == 3) break;
                if (what == 0) { __nanosleep(200); continue; }
                if (what == 2) {
                    const uint32_t chunk = dir[k0 + arg], ri = chunk * AF_CHUNK + lane;
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
                        if (seeded) st_vol(&queue[pos + __popc(todo & ((1u << lane) - 1u))], ri);
                        n_seeded += lane == 0 ? __popc(todo) : 0;
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) atomicAdd(&s_done, 1u);   // after its entries are in the queue
                    continue;
                }
                // what == 1: the whole warp extends queued candidate `arg`
                uint32_t ri;
                while ((ri = ld_vol(&queue[arg])) == AF_REC_INVALID) { }   // reserved a moment ago, about to be written
                const uint32_t *rec = a.recs + (size_t)ri * (a.rq * 4);
                const uint32_t rid = rec[0];
                const uint32_t rw = lane < a.W ? rec[4 + (rid & 1u) * a.W + lane] : 0u;
                const int L = a.uniform_len > 0 ? a.uniform_len : (int)a.lens[rid];
                const uint32_t *nm = nmask_of(rid);
                const uint32_t nwv = nm && lane < AF_NMASK_WORDS ? nm[lane] : 0u;
                uint4 out;
                if (tail_extend(rid, L, rw, nwv, nm != nullptr, a.table, a.tmask, apk0p, apk1p, a.apn0p, a.apn1p, a.G, KP, 20 - KP, a.P, lane, out)) {
                    if (lane == 0) {
                        // the anchored read's record replaces the header quad of its own candidate record; its bit goes
                        // into the region's hit bitmap and into its chunk's hit mask
                        reinterpret_cast<uint4 *>(a.recs)[(size_t)ri * a.rq] = out;
                        const uint32_t bit = rid - rid0;
                        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
                        atomicOr(&a.chunk_hits[ri / AF_CHUNK], 1u << (ri % AF_CHUNK));
                    }
                }
            }
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
    if (lane == 0 && n_seeded) atomicAdd(&s_seeded, n_seeded);
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
