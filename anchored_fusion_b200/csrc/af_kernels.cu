// af_kernels.cu -- sm_100a kernels of the read-anchoring path and their stream-ordered drivers.
//
//   k_seed_scan      HBM-bound filter: streams the 2-bit packed tiles with coalesced 128-bit
//                    loads, cuts every read into k'-mers at stride s (k'+s-1 <= k, so every exact
//                    match of >= k bases contains one), probes a shared-memory fingerprint table
//                    of the anchor's k'-mers (both strands), ballots one flag word per 32 reads.
//                    No false negatives; false positives are removed by k_extend.
//   k_flag_count /   ballot + prefix-sum stream compaction of the flag words into a candidate
//   k_flag_scatter   read list ordered by read_id.
//   k_extend         one warp per candidate: exact k'-mer table lookups (lanes = sample
//                    positions), match_any de-duplication of diagonals, 256-bit match masks by
//                    ballot, leftmost >=k run, X-drop extension with warp prefix-max scans.
//   k_hit_count /    compaction of the per-candidate result slots into the hit list.
//   k_hit_scatter
//   k_synth_pairs    seeded synthetic read pairs written straight into packed tiles.
//
// Stands in for `bwa mem -M ... | samtools view -F 772` (Anchored_Fusion.py:182,194); the
// semantics are "Anchoring spec v1" in DESIGN.md, restated on the CPU in oracle/af_oracle.c.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "af_common.h"
#include "af_device.cuh"

#ifndef AF_SCAN_BOUND
#define AF_SCAN_BOUND 768   // __launch_bounds__ of the seed scan (register cap 65536 / bound); launched with <= 768 threads
#endif

static std::atomic<long long> g_launches{0};
extern "C" int64_t af_kernel_launches(void) { return (int64_t)g_launches.load(); }

// ---- optional per-stage timing with CUDA events on the launching stream (bench.py) ---------
#include <mutex>
#include <vector>
struct ProfSpan { cudaEvent_t a, b; int stage; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_prof;
static std::mutex g_prof_mu;
static void prof_mark(cudaEvent_t *ev, cudaStream_t st) {
    *ev = nullptr;
    if (!g_prof_on) return;
    if (cudaEventCreate(ev) == cudaSuccess) cudaEventRecord(*ev, st);
}
static void prof_span(cudaEvent_t a, cudaStream_t st, int stage) {
    if (!g_prof_on || !a) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({a, b, stage});
}
void af_note_launches(int n) { g_launches += n; }
void af_prof_mark(cudaEvent_t *ev, cudaStream_t st) { prof_mark(ev, st); }
void af_prof_span(cudaEvent_t a, cudaStream_t st, int stage) { prof_span(a, st, stage); }
extern "C" void af_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &p : g_prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    g_prof.clear();
    g_prof_on = true;
}
// Call after the stream has been synchronised.  ms_out[5] = summed device time of the seed scan,
// flag compaction, verify (+ its compaction), extension, hit compaction; calls_out[5] = spans.
extern "C" int af_profile_end(double *ms_out, int64_t *calls_out) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    for (int i = 0; i < ST_N; i++) { ms_out[i] = 0; calls_out[i] = 0; }
    for (auto &p : g_prof) {
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { ms_out[p.stage] += ms; calls_out[p.stage]++; }
        cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    g_prof.clear();
    return AF_OK;
}

// ------------------------------------------------------------------------------------------
// seed scan
// ------------------------------------------------------------------------------------------
static const int CB_THREADS = 256, CB_ITEMS = 8, CB_PER_BLOCK = CB_THREADS * CB_ITEMS;   // compaction chunk

template <int Q>
__device__ __forceinline__ void load_tile(uint32_t (&w)[4 * Q], const uint4 *__restrict__ packed, long long tile, int lane) {
    const uint4 *src = packed + tile * (Q * 32) + lane;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        uint4 v = ld_stream_v4(src + q * 32);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
}

__device__ __forceinline__ uint32_t tile_valid_mask(long long tile, long long n_pairs) {
    long long left = n_pairs - tile * 32;
    return left >= 32 ? FULL : (left <= 0 ? 0u : ((1u << left) - 1u));
}

static const int SCAN_LOCAL_CHUNKS = 64;
static const int RQ_CAP = 96;   // per-warp refine queue: up to 31 waiting + the 64 reads of one tile

// flags[tile] = (ballot of mate-1 flags, ballot of mate-2 flags), lanes past n_pairs cleared.
// The number of flagged reads per compaction chunk is accumulated in shared memory (cc_local,
// chunks relative to the CTA's first chunk) and flushed once per CTA: every CTA owns a
// contiguous tile range, so concurrent CTAs never hammer the same global counter.
template <int W, int KP, int Q, bool BLOOM = false>
__device__ __forceinline__ void scan_tile(const uint32_t (&w)[4 * Q], long long tile, long long n_pairs, int lane,
                                          int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                                          uint2 *__restrict__ flags, uint32_t *cc_local, long long chunk0,
                                          uint32_t *__restrict__ chunk_counts) {
    uint32_t a1, a2;
    af_scan_pair<W, KP, 4 * Q, BLOOM>(w, nprobe, filt, fmul, nb, a1, a2);
    const uint32_t vm = tile_valid_mask(tile, n_pairs);
    const uint32_t b1 = __ballot_sync(FULL, a1 != 0) & vm, b2 = __ballot_sync(FULL, a2 != 0) & vm;
    if (lane == 0) {
        flags[tile] = make_uint2(b1, b2);
        const uint32_t c = __popc(b1) + __popc(b2);
        if (chunk_counts && c) {
            const long long rel = tile / CB_PER_BLOCK - chunk0;
            if (rel < SCAN_LOCAL_CHUNKS) atomicAdd(&cc_local[rel], c);
            else atomicAdd(&chunk_counts[tile / CB_PER_BLOCK], c);
        }
    }
}

// ---- in-scan refinement (RQ = true) --------------------------------------------------------
// 1.8 % of the reads pass the filter, 86 % of them because one sampled k'-mer really is an anchor
// k'-mer by chance.  A second look removes nearly all of them (af_neighbour_ok: a >= k match
// around a sample also holds the k'-mer 4 bases to its left or right; 362 k -> 16 k reads per
// 10 M pairs) but costs a branch per sample, which halves the speed of the probe loop when it is
// taken lane by lane.  So the scan only QUEUES the flagged reads' ids (per warp, shared memory)
// and, whenever 32 are waiting, the warp runs the second look for 32 flagged reads at once:
// dense lanes, the reads' quads fetched again from L2 (they were streamed a few microseconds
// ago), ~2 % of the scan's instructions.  What survives is OR-ed into the flag words, so the
// stage that re-fetches 362 k scattered reads from HBM (k_verify_smem + k_sel_scatter, 45 us)
// disappears: the survivors go straight to k_extend.
// MEASURED (B200, 10 M pairs): the scan goes from 186 us to 233 us -- 8.8 % more instructions, 22 %
// more DRAM reads (the re-fetched quads mostly miss L2), and the warps that sit in a refinement pass
// leave the other warps of their scheduler short of latency cover (LSU pipe 92 % -> 79 % busy).  The
// step does not get shorter (0.306 vs 0.299 ms) and the scan's roofline fraction drops to 0.50, so
// this stays an option (af_seed_scan_config(0, 11)) and the separate verify stage is the default.
template <int W, int KP, int Q>
__device__ __forceinline__ void refine_pass(const uint4 *__restrict__ packed, uint32_t rid, bool have, int nprobe,
                                            const uint32_t *filt, uint32_t fmul, uint32_t nb,
                                            uint32_t *__restrict__ flags, uint32_t *__restrict__ chunk_counts) {
    const uint32_t pair = rid >> 1, mate = rid & 1u;
    const uint4 *src = packed + (size_t)(pair >> 5) * (Q * 32) + (pair & 31);
    uint32_t w[W + 1];
#pragma unroll
    for (int t = 0; t <= W; t++) w[t] = 0u;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (have) v = __ldg(src + q * 32);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int wi = 4 * q + e;
            if (wi < W) w[wi] = mate ? w[wi] : vv[e];
            else if (wi < 2 * W) w[wi - W] = mate ? vv[e] : w[wi - W];
        }
    }
    const uint32_t r = af_scan_read<W, KP, 0, W + 1, true>(w, nprobe, filt, fmul, nb);
    if (have && r) {
        atomicOr(&flags[(size_t)(pair >> 5) * 2 + mate], 1u << (pair & 31));
        atomicAdd(&chunk_counts[(pair >> 5) / CB_PER_BLOCK], 1u);
    }
}

template <int W, int KP, int Q>
__device__ __forceinline__ void rq_drain(uint32_t *q, int &qn, int lane, int threshold, const uint4 *__restrict__ packed,
                                         int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                                         uint32_t *__restrict__ flags, uint32_t *__restrict__ chunk_counts) {
    while (qn >= threshold && qn > 0) {                      // warp-uniform
        __syncwarp();                                        // queue entries and the zeroed flag words of this warp are visible
        const bool have = lane < qn;
        const uint32_t rid = have ? q[lane] : 0u;
        const uint32_t m1 = 32 + lane < qn ? q[32 + lane] : 0u, m2 = 64 + lane < qn ? q[64 + lane] : 0u;
        __syncwarp();
        q[lane] = m1; q[32 + lane] = m2;
        __syncwarp();
        qn = qn > 32 ? qn - 32 : 0;
        refine_pass<W, KP, Q>(packed, rid, have, nprobe, filt, fmul, nb, flags, chunk_counts);
    }
}

template <int W, int KP, int Q>
__device__ __forceinline__ void scan_tile_rq(const uint32_t (&w)[4 * Q], long long tile, long long n_pairs, int lane,
                                             int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                                             const uint4 *__restrict__ packed, uint32_t *__restrict__ flags,
                                             uint32_t *__restrict__ chunk_counts, uint32_t *q, int &qn, uint32_t &nflag) {
    uint32_t a1 = af_scan_read<W, KP, 0, 4 * Q>(w, nprobe, filt, fmul, nb);
    uint32_t a2 = af_scan_read<W, KP, W, 4 * Q>(w, nprobe, filt, fmul, nb);
    const uint32_t vm = tile_valid_mask(tile, n_pairs);
    const uint32_t b1 = __ballot_sync(FULL, a1 != 0) & vm, b2 = __ballot_sync(FULL, a2 != 0) & vm;
    if (lane == 0) reinterpret_cast<uint2 *>(flags)[tile] = make_uint2(0u, 0u);   // refined bits are OR-ed in later
    if (b1 | b2) {
        const uint32_t lt = (1u << lane) - 1u, rid = (uint32_t)(tile * 32 + lane) * 2u;
        if ((b1 >> lane) & 1u) q[qn + __popc(b1 & lt)] = rid;
        qn += __popc(b1);
        if ((b2 >> lane) & 1u) q[qn + __popc(b2 & lt)] = rid + 1u;
        qn += __popc(b2);
        nflag += __popc(b1) + __popc(b2);
        rq_drain<W, KP, Q>(q, qn, lane, 32, packed, nprobe, filt, fmul, nb, flags, chunk_counts);
    }
}

// ---- candidate emission (EMIT = true) ---------------------------------------------------------
// A flagged lane still holds its read in registers: it is stored as one record of the candidate stream
// (af_emit, af_device.cuh) instead of being re-gathered from HBM later.  The warp owns the chunk it fills
// (no atomics per record); a chunk is taken from the pool with one atomic per 32 records.
struct EmitState { uint32_t cur, end, reg, reg_end, nflag; };

__device__ __forceinline__ void emit_pad(EmitState &S, const af_emit &E, int rq, int lane) {
    const uint32_t n = S.end - S.cur;                       // < 32 unused slots of the open chunk
    if ((uint32_t)lane < n) E.recs[(size_t)(S.cur + lane) * (rq * 4)] = AF_REC_INVALID;
    S.cur = S.end;
}

// One chunk from the pool, listed in region `reg`'s directory (lane 0 only).  AF_REC_INVALID when the pool is empty.
__device__ __forceinline__ uint32_t emit_grab(uint32_t reg, const af_emit &E) {
    uint32_t c = atomicAdd(E.pool, 1u);
    if (c < E.pool_chunks) {
        const uint32_t k = atomicAdd(&E.dir_count[reg], 1u);
        if (k < AF_DIR_CAP) E.dir[(size_t)reg * AF_DIR_CAP + k] = c;
        else c = AF_REC_INVALID;                             // cannot happen (AF_DIR_CAP covers a fully flagged region)
    }
    if (c >= E.pool_chunks) { atomicOr(&E.counts[AF_CNT_STATUS], AF_STATUS_CAND_OVERFLOW); c = AF_REC_INVALID; }
    return c;
}

// A record = one header quad {read_id, 0, 0, 0} + the Q quads of the read's PAIR exactly as the tile load left
// them in registers (aligned register quads: no repacking; the tail picks the mate's words by read_id & 1).
template <int Q>
__device__ __forceinline__ void emit_store(uint32_t at, uint32_t rid, const uint32_t (&w)[4 * Q], const af_emit &E) {
    if (at == AF_REC_INVALID) return;
    uint4 *dst = reinterpret_cast<uint4 *>(E.recs) + (size_t)at * (Q + 1);
    dst[0] = make_uint4(rid, 0u, 0u, 0u);
#pragma unroll
    for (int q = 0; q < Q; q++) dst[1 + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

template <int W, int KP, int Q>
__device__ __forceinline__ void scan_tile_emit(const uint32_t (&w)[4 * Q], long long tile, long long n_tiles, long long n_pairs,
                                               int lane, int nprobe, const uint32_t *filt, uint32_t fmul, uint32_t nb,
                                               const af_emit &E, EmitState &S) {
    uint32_t a1 = af_scan_read<W, KP, 0, 4 * Q>(w, nprobe, filt, fmul, nb);
    uint32_t a2 = af_scan_read<W, KP, W, 4 * Q>(w, nprobe, filt, fmul, nb);
    const uint32_t vm = tile_valid_mask(tile, n_pairs);
    const uint32_t b1 = __ballot_sync(FULL, a1 != 0) & vm, b2 = __ballot_sync(FULL, a2 != 0) & vm;
    if (b1 | b2) {                                           // ~70 % of the tiles hold a flagged read (1.2 on average)
        if (E.m > 1) {                                       // the warp's tiles ascend: step to the region of this tile
            while ((uint32_t)tile >= S.reg_end) {
                emit_pad(S, E, Q + 1, lane);
                S.reg++;
                S.reg_end = (uint32_t)(n_tiles * (long long)(S.reg + 1) / ((long long)gridDim.x * E.m));
            }
        }
        const int n1 = __popc(b1), n = n1 + __popc(b2), room = (int)(S.end - S.cur);
        uint32_t nb0 = AF_REC_INVALID, nb1 = AF_REC_INVALID;
        if (n > room) {                                      // once per 32 records: one or (n > 32 + room) two new chunks
            uint32_t c0 = AF_REC_INVALID, c1 = AF_REC_INVALID;
            if (lane == 0) {
                c0 = emit_grab(S.reg, E);
                if (n - room > AF_CHUNK) c1 = emit_grab(S.reg, E);
            }
            c0 = __shfl_sync(FULL, c0, 0); c1 = __shfl_sync(FULL, c1, 0);
            if (c0 != AF_REC_INVALID) nb0 = c0 * AF_CHUNK;
            if (c1 != AF_REC_INVALID) nb1 = c1 * AF_CHUNK;
        }
        auto slot = [&](int k) -> uint32_t {
            if (k < room) return S.cur + (uint32_t)k;
            k -= room;
            if (k < AF_CHUNK) return nb0 == AF_REC_INVALID ? AF_REC_INVALID : nb0 + (uint32_t)k;
            return nb1 == AF_REC_INVALID ? AF_REC_INVALID : nb1 + (uint32_t)(k - AF_CHUNK);
        };
        const uint32_t lt = (1u << lane) - 1u, rid = (uint32_t)(tile * 32 + lane) * 2u;
        if ((b1 >> lane) & 1u) emit_store<Q>(slot(__popc(b1 & lt)), rid, w, E);
        if ((b2 >> lane) & 1u) emit_store<Q>(slot(n1 + __popc(b2 & lt)), rid + 1u, w, E);
        if (n > room) {
            const bool two = n - room > AF_CHUNK;
            const uint32_t last = two ? nb1 : nb0;
            if (last != AF_REC_INVALID) { S.cur = last + (uint32_t)(n - room - (two ? AF_CHUNK : 0)); S.end = last + AF_CHUNK; }
            else S.cur = S.end = 0;                          // pool exhausted (status set): the rest is dropped
        } else S.cur += (uint32_t)n;
        S.nflag += (uint32_t)n;
    }
}

// Persistent kernel, one CTA per SM: the anchor filter is staged into shared memory once, then
// the CTA's warps walk its contiguous range of tiles (32 pairs per tile, one pair per lane, all
// of it in registers).
// PF = true (<= 768 threads): the loads of the warp's NEXT tile are issued before the current
// tile is scanned (register double buffer), so HBM latency overlaps the probes of the same warp.
// PF = false (<= 1024 threads, 64 registers): latency is hidden by occupancy alone.
// RQ = true: flags[] receives the REFINED flag words (see refine_pass), chunk_counts their per-chunk
// counts and counts[AF_CNT_FLAGGED] the number of reads that passed the plain filter.
// BLOOM = true (plain mode only): the filter words are Bloom bits (long anchors, af_bloom_probe).
// EMIT = true: no flag words at all -- every flagged read goes into the candidate stream (emit_reads) that
// k_tail consumes; counts[AF_CNT_FLAGGED] receives their number.
template <int W, int KP, int MAXT, bool PF, bool RQ, bool EMIT = false, bool BLOOM = false>
__global__ void __launch_bounds__(MAXT, 1)
k_seed_scan(const uint4 *__restrict__ packed, long long n_tiles, long long n_pairs, int nprobe,
            const uint32_t *__restrict__ g_filter, uint32_t fmul, uint32_t nb, uint2 *__restrict__ flags,
            uint32_t *__restrict__ chunk_counts, uint32_t *__restrict__ counts, const af_emit E) {
    extern __shared__ __align__(128) uint32_t filt[];
    __shared__ uint32_t cc_local[SCAN_LOCAL_CHUNKS];
    __shared__ uint32_t rq[RQ ? (MAXT / 32) * RQ_CAP : 1];
    constexpr int Q = (2 * W + 3) / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long t_begin = n_tiles * blockIdx.x / gridDim.x, t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const long long chunk0 = t_begin / CB_PER_BLOCK;
    const long long stride = nwarps;
    long long tile = t_begin + warp;
    if (threadIdx.x < SCAN_LOCAL_CHUNKS) cc_local[threadIdx.x] = 0;
    uint32_t *q = rq + (RQ ? warp * RQ_CAP : 0);
    int qn = 0;
    uint32_t nflag = 0;
    EmitState S;
    S.cur = S.end = 0; S.nflag = 0;
    S.reg = EMIT ? blockIdx.x * (uint32_t)E.m : 0u;
    S.reg_end = EMIT ? (uint32_t)(n_tiles * (long long)(S.reg + 1) / ((long long)gridDim.x * E.m)) : 0u;
    auto do_tile = [&](const uint32_t (&w)[4 * Q], long long t) {
        if constexpr (EMIT) scan_tile_emit<W, KP, Q>(w, t, n_tiles, n_pairs, lane, nprobe, filt, fmul, nb, E, S);
        else if constexpr (RQ) scan_tile_rq<W, KP, Q>(w, t, n_pairs, lane, nprobe, filt, fmul, nb, packed, (uint32_t *)flags, chunk_counts, q, qn, nflag);
        else scan_tile<W, KP, Q, BLOOM>(w, t, n_pairs, lane, nprobe, filt, fmul, nb, flags, cc_local, chunk0, chunk_counts);
    };
    if constexpr (PF) {
        uint32_t wa[4 * Q], wb[4 * Q];
        if (tile < t_end) load_tile<Q>(wa, packed, tile, lane);   // in flight while the filter is staged
        stage_filter(filt, g_filter, nb);
        __syncthreads();
        while (tile < t_end) {
            const long long t2 = tile + stride;
            if (t2 < t_end) load_tile<Q>(wb, packed, t2, lane);
            do_tile(wa, tile);
            if (t2 >= t_end) break;
            const long long t3 = t2 + stride;
            if (t3 < t_end) load_tile<Q>(wa, packed, t3, lane);
            do_tile(wb, t2);
            tile = t3;
        }
    } else {
        stage_filter(filt, g_filter, nb);
        __syncthreads();
        for (; tile < t_end; tile += stride) {
            uint32_t w[4 * Q];
            load_tile<Q>(w, packed, tile, lane);
            do_tile(w, tile);
        }
    }
    if constexpr (EMIT) {
        emit_pad(S, E, Q + 1, lane);
        if (lane == 0 && S.nflag) atomicAdd(&cc_local[0], S.nflag);
        __syncthreads();
        if (threadIdx.x == 0 && cc_local[0]) atomicAdd(&counts[AF_CNT_FLAGGED], cc_local[0]);
    } else if constexpr (RQ) {
        rq_drain<W, KP, Q>(q, qn, lane, 1, packed, nprobe, filt, fmul, nb, (uint32_t *)flags, chunk_counts);
        if (lane == 0 && nflag) atomicAdd(&cc_local[0], nflag);
        __syncthreads();
        if (threadIdx.x == 0 && cc_local[0]) atomicAdd(&counts[AF_CNT_FLAGGED], cc_local[0]);
    } else if (chunk_counts) {
        __syncthreads();
        if (threadIdx.x < SCAN_LOCAL_CHUNKS && cc_local[threadIdx.x]) atomicAdd(&chunk_counts[chunk0 + threadIdx.x], cc_local[threadIdx.x]);
    }
}

static long long *g_tail_dbg = nullptr;   // af_debug_tail_timing
static int env_flag(const char *name, int dflt) { const char *v = getenv(name); return v && *v ? atoi(v) : dflt; }
static int g_scan_threads = 768, g_fused = 0, g_middle = 7, g_verify_smem = 1;
static int g_walk = 1;                             // k_extend evaluates diagonals with eval_mask_walk (1) or eval_diag (0); af_seed_scan_config(0, 14 / 15)
static int g_stream = env_flag("AF_STREAM", 0);   // 1: candidate-stream path (k_seed_scan<EMIT> + k_tail); also af_seed_scan_config(0, 12 / 13)
// tuning knobs.  Scan variant (mode 0/3): register double buffer under an 85-register cap, up to 768
// threads = 24 warps per SM (a 512-thread / 128-register variant and a 1024-thread variant without
// prefetch measured the same and were dropped to keep the build short).
extern "C" int af_seed_scan_config(int32_t threads_per_block, int32_t mode) {
    // modes 4 / 5 switch af_anchor_batch between the fused, warp-specialised scan+verify kernel (4)
    // and the separate seed-scan / verify kernels (5, default); the stand-alone scan variant is untouched
    if (mode == 4 || mode == 5) { g_fused = mode == 4; return AF_OK; }
    // modes 7 / 8: what stands between the scan and k_extend -- 7 = k_verify (default: exact seeded test
    // against the table and the anchor), 8 = nothing (every flagged read gets a warp of k_extend; for
    // experiments with a scan built with the neighbour test, af_scan_read<..., REFINE = true>)
    if (mode == 7 || mode == 8) { g_middle = mode == 8 ? 0 : 7; return AF_OK; }
    // mode 11 (experiment, measured slower: profiles/r02_scan_refine_queue_ncu.md): the second look at flagged
    // reads happens inside the scan (refine queue, see refine_pass) and the survivors go straight to k_extend;
    // an index whose filter is saturated (long anchor) still takes the k_verify route
    if (mode == 11) { g_middle = 11; return AF_OK; }
    // modes 9 / 10: k_verify_smem (9, default: membership from a half-size filter in shared memory) or
    // k_verify (10: membership from the L2-resident bitmap)
    if (mode == 9 || mode == 10) { g_verify_smem = mode == 9; return AF_OK; }
    // modes 12 / 13: the candidate-stream path (12, default: k_seed_scan<EMIT> + k_tail, 2 kernels) or the
    // six-kernel path (13: scan -> flag compaction -> verify -> selection -> extend -> hit compaction)
    if (mode == 12 || mode == 13) { g_stream = mode == 12; return AF_OK; }
    // modes 14 / 15: k_extend's diagonal evaluation -- 14 (default) word-parallel mask + mismatch walk, 15 the round-1 code
    if (mode == 14 || mode == 15) { g_walk = mode == 14; return AF_OK; }
    if (mode != 0 && mode != 3) { af_set_error("af_seed_scan_config: mode must be 0/3 (scan variant), 4/5 (fused on/off), 7/8/11, 9/10, 12/13 or 14/15"); return AF_ERR_ARG; }
    const int maxt = 768;
    if (threads_per_block == 0) threads_per_block = maxt;
    if (threads_per_block < 64 || threads_per_block > maxt || threads_per_block % 32) { af_set_error("af_seed_scan_config: threads must be 64..%d, multiple of 32", maxt); return AF_ERR_ARG; }
    g_scan_threads = threads_per_block;
    return AF_OK;
}

// SM partition (af_sm_partition): the scan's persistent grid leaves `g_small_sms` SMs to the kernels of the previous
// batch (the verify stage needs whole SMs for its shared-memory filter and can never sit next to a scan CTA); 0 = none.
static int g_small_sms = env_flag("AF_SMALL_SMS", 0);
static int scan_grid(const af_dev_index *d, long long n_tiles, int threads = 0) {
    const int nwarps = (threads ? threads : g_scan_threads) / 32;
    const long long want = (n_tiles + nwarps - 1) / nwarps;
    const int sms = g_small_sms > 0 && g_small_sms < d->num_sms ? d->num_sms - g_small_sms : d->num_sms;
    return (int)(want < sms ? (want > 0 ? want : 1) : sms);
}

template <int W, int KP, int MAXT, bool PF, bool RQ, bool EMIT, bool BLOOM = false>
static int launch_scan(const af_dev_index *d, const af_batch_t *b, long long n_tiles, int nprobe, uint32_t *flags,
                       uint32_t *chunk_counts, uint32_t *counts, const af_emit &E, cudaStream_t st, int threads = 0) {
    if (!threads) threads = g_scan_threads;
    size_t smem = (size_t)d->nb * 4;
    static bool attr_set[64] = {false};  // per device
    if (!attr_set[d->device & 63]) {
        int rc = allow_full_smem(k_seed_scan<W, KP, MAXT, PF, RQ, EMIT, BLOOM>, nullptr);  // the filter + a few static words (chunk counters, mbarrier, refine queues)
        if (rc) return rc;
        attr_set[d->device & 63] = true;
    }
    const int grid = scan_grid(d, n_tiles, threads);
    k_seed_scan<W, KP, MAXT, PF, RQ, EMIT, BLOOM><<<grid, threads, smem, st>>>((const uint4 *)b->packed, n_tiles, b->n_pairs, nprobe,
                                                                               d->d_filter, d->fmul, d->nb, (uint2 *)flags, chunk_counts, counts, E);
    g_launches++;
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}

template <int W, int KP>
static int launch_scan_mode(const af_dev_index *d, const af_batch_t *b, long long n_tiles, int nprobe, uint32_t *flags,
                            uint32_t *cc, uint32_t *counts, bool rq, const af_emit *emit, cudaStream_t st) {
    if (d->bloom) {
        if (emit || rq) { af_set_error("seed scan: a Bloom-filter index (long anchor) runs the plain scan only"); return AF_ERR_ARG; }
        return launch_scan<W, KP, AF_SCAN_BOUND, true, false, false, true>(d, b, n_tiles, nprobe, flags, cc, counts, af_emit(), st);
    }
    if (emit) return launch_scan<W, KP, AF_SCAN_BOUND, true, false, true>(d, b, n_tiles, nprobe, flags, cc, counts, *emit, st);
    if (rq) return launch_scan<W, KP, AF_SCAN_BOUND, true, true, false>(d, b, n_tiles, nprobe, flags, cc, counts, af_emit(), st);
    return launch_scan<W, KP, AF_SCAN_BOUND, true, false, false>(d, b, n_tiles, nprobe, flags, cc, counts, af_emit(), st);
}

// Reads of 257..512 bases (W = 20, 24, 28, 32; af_layout rounds W up to a multiple of 4 beyond 16): the plain scan
// on 384 threads, with the register double buffer up to W = 24 (a pair alone is up to 64 registers).  2x300 MiSeq runs and
// merged pairs are small next to the 2x150 bulk, so this instance is built for coverage, not tuned.
static const int AF_LONG_SCAN_THREADS = 384;
template <int W, int KP>
static int launch_scan_long(const af_dev_index *d, const af_batch_t *b, long long n_tiles, int nprobe, uint32_t *flags,
                            uint32_t *cc, uint32_t *counts, bool rq, const af_emit *emit, cudaStream_t st) {
    if (rq || emit) { af_set_error("seed scan: reads beyond 256 bases run the plain scan only"); return AF_ERR_ARG; }
    const int threads = g_scan_threads < AF_LONG_SCAN_THREADS ? g_scan_threads : AF_LONG_SCAN_THREADS;
    constexpr bool PF = W <= 24;                            // two tiles of 2 x 24 words still fit the 170 registers of a 384-thread CTA
    return d->bloom ? launch_scan<W, KP, AF_LONG_SCAN_THREADS, PF, false, false, true>(d, b, n_tiles, nprobe, flags, cc, counts, af_emit(), st, threads)
                    : launch_scan<W, KP, AF_LONG_SCAN_THREADS, PF, false, false, false>(d, b, n_tiles, nprobe, flags, cc, counts, af_emit(), st, threads);
}
#define AF_SCAN_CASE_LONG(WW)                                                                                  \
    case WW:                                                                                                   \
        return kp == 12 ? launch_scan_long<WW, 12>(d, b, n_tiles, nprobe, flags, cc, counts, rq, emit, st)     \
                        : launch_scan_long<WW, 13>(d, b, n_tiles, nprobe, flags, cc, counts, rq, emit, st);

#define AF_SCAN_CASE(WW)                                                                                       \
    case WW:                                                                                                   \
        return kp == 12 ? launch_scan_mode<WW, 12>(d, b, n_tiles, nprobe, flags, cc, counts, rq, emit, st)     \
                        : launch_scan_mode<WW, 13>(d, b, n_tiles, nprobe, flags, cc, counts, rq, emit, st);

static int batch_check(const af_dev_index *d, const af_batch_t *b, af_layout_t *lay) {
    if (!d || !b) { af_set_error("null index or batch"); return AF_ERR_ARG; }
    int rc = af_layout(b->max_read_len, b->n_pairs, lay);
    if (rc) return rc;
    if (b->n_pairs >= (1ll << 31)) { af_set_error("a batch holds at most 2^31-1 pairs"); return AF_ERR_ARG; }
    if (b->n_pairs && !b->packed) { af_set_error("batch.packed is null"); return AF_ERR_ARG; }
    if (b->uniform_len <= 0 && !b->lens && b->n_pairs) { af_set_error("batch needs uniform_len or lens"); return AF_ERR_ARG; }
    if (b->uniform_len > b->max_read_len) { af_set_error("uniform_len exceeds max_read_len"); return AF_ERR_ARG; }
    return AF_OK;
}

static int seed_scan_impl(const af_dev_index *d, const af_batch_t *b, uint32_t *flags, uint32_t *cc, uint32_t *counts, bool rq,
                          const af_emit *emit, cudaStream_t st) {
    af_layout_t lay;
    int rc = batch_check(d, b, &lay);
    if (rc) return rc;
    if (lay.n_tiles == 0) return AF_OK;
    const int kp = d->kp;
    if (d->P.k != 19 || (kp != 12 && kp != 13)) { af_set_error("seed scan is built for k=19, k' in {12,13}"); return AF_ERR_ARG; }
    // sample positions p_j = (k - k') + j*s (af_common.h); L = longest read of the batch
    int L = b->uniform_len > 0 ? b->uniform_len : b->max_read_len;
    int nprobe = af_nsamples(L, kp);
    long long n_tiles = lay.n_tiles;
    switch (lay.words_per_read) {
        AF_SCAN_CASE(1) AF_SCAN_CASE(2) AF_SCAN_CASE(3) AF_SCAN_CASE(4) AF_SCAN_CASE(5) AF_SCAN_CASE(6)
        AF_SCAN_CASE(7) AF_SCAN_CASE(8) AF_SCAN_CASE(9) AF_SCAN_CASE(10) AF_SCAN_CASE(11) AF_SCAN_CASE(12)
        AF_SCAN_CASE(13) AF_SCAN_CASE(14) AF_SCAN_CASE(15) AF_SCAN_CASE(16)
        AF_SCAN_CASE_LONG(20) AF_SCAN_CASE_LONG(24) AF_SCAN_CASE_LONG(28) AF_SCAN_CASE_LONG(32)
    }
    af_set_error("unsupported words_per_read %d", lay.words_per_read);
    return AF_ERR_ARG;
}

extern "C" int af_seed_scan(const af_dev_index_t *d, const af_batch_t *batch, uint32_t *d_flags, void *stream) {
    AF_CUDA(cudaSetDevice(d ? d->device : 0));
    return seed_scan_impl(d, batch, d_flags, nullptr, nullptr, false, nullptr, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// stream compaction (ballot / popc / prefix sums); deterministic, ordered by read_id.
// Every producer kernel leaves per-chunk counts (chunk = CB_PER_BLOCK items) behind with a
// handful of atomics, so each compaction is ONE scatter kernel: chunk base = sum of the counts
// of the chunks before it, offsets inside the chunk by a block-wide exclusive scan.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *smem /*>=9 words*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < (CB_THREADS / 32) ? smem[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, si, o); if (lane >= o) si += t; }
        if (lane < (CB_THREADS / 32)) smem[lane] = si - s;
    }
    __syncthreads();
    uint32_t r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ uint32_t sum_before(const uint32_t *chunk_counts, uint32_t upto, uint32_t *sm) {
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i < upto; i += blockDim.x) s += chunk_counts[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    uint32_t r = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += sm[i];
    __syncthreads();
    return r;
}

// flag words -> candidate read_ids; total -> counts[AF_CNT_FLAGGED]
__global__ void __launch_bounds__(CB_THREADS)
k_flag_scatter(const uint2 *__restrict__ flags, long long n_tiles, const uint32_t *__restrict__ chunk_counts,
               uint32_t n_chunks, uint32_t *__restrict__ cand, uint32_t cand_cap, uint32_t *counts, int count_idx,
               int count_idx2) {
    __shared__ uint32_t sm[9];
    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const uint32_t mine = chunk_counts[chunk];
        if (mine == 0) continue;                       // uniform across the block
        const uint32_t base = sum_before(chunk_counts, chunk, sm);
        if (threadIdx.x == 0) { atomicAdd(&counts[count_idx], mine); if (count_idx2 >= 0) atomicAdd(&counts[count_idx2], mine); }
        uint2 f[CB_ITEMS];
        uint32_t c = 0;
        const long long t0 = (long long)chunk * CB_PER_BLOCK + threadIdx.x * CB_ITEMS;
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++) {
            f[i] = (t0 + i < n_tiles) ? flags[t0 + i] : make_uint2(0, 0);
            c += __popc(f[i].x) + __popc(f[i].y);
        }
        uint32_t off = base + block_excl_scan(c, sm);
        bool over = false;
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++) {
            uint32_t any = f[i].x | f[i].y;
            while (any) {
                const int l = __ffs(any) - 1;
                any &= any - 1;
                const uint32_t pair = (uint32_t)((t0 + i) * 32 + l);
                if ((f[i].x >> l) & 1) { if (off < cand_cap) cand[off] = pair * 2; else over = true; off++; }
                if ((f[i].y >> l) & 1) { if (off < cand_cap) cand[off] = pair * 2 + 1; else over = true; off++; }
            }
        }
        if (over) atomicOr(&counts[AF_CNT_STATUS], AF_STATUS_CAND_OVERFLOW);
    }
}

// candidates kept by k_verify -> seeded list; total -> counts[AF_CNT_SEEDED]
__global__ void __launch_bounds__(CB_THREADS)
k_sel_scatter(const uint32_t *__restrict__ items, const uint8_t *__restrict__ keep, uint32_t cap,
              const uint32_t *__restrict__ chunk_counts, uint32_t *__restrict__ out, uint32_t *counts) {
    __shared__ uint32_t sm[9];
    const uint32_t n = min(counts[AF_CNT_FLAGGED], cap), n_chunks = (n + CB_PER_BLOCK - 1) / CB_PER_BLOCK;
    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const uint32_t mine = chunk_counts[chunk];
        if (mine == 0) continue;
        const uint32_t base = sum_before(chunk_counts, chunk, sm);
        if (threadIdx.x == 0) atomicAdd(&counts[AF_CNT_SEEDED], mine);
        const uint32_t i0 = chunk * CB_PER_BLOCK + threadIdx.x * CB_ITEMS;
        uint32_t v[CB_ITEMS], c = 0;
        bool k[CB_ITEMS];
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++) {
            k[i] = (i0 + i < n) && keep[i0 + i] != 0;
            v[i] = k[i] ? items[i0 + i] : 0u;
            c += k[i];
        }
        uint32_t off = base + block_excl_scan(c, sm);
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++)
            if (k[i]) out[off++] = v[i];   // out has the capacity of items: cannot overflow
    }
}

// result slots of k_extend with m_len != 0 -> hit list; total -> counts[AF_CNT_HITS]
template <bool SINK>
__global__ void __launch_bounds__(CB_THREADS)
k_hit_scatter(const uint4 *__restrict__ slots, uint32_t cap, const uint32_t *__restrict__ chunk_counts,
              uint4 *__restrict__ hits, uint32_t hits_cap, uint32_t *counts, const af_sink sink) {
    __shared__ uint32_t sm[9];
    const uint32_t n = min(counts[AF_CNT_SEEDED], cap), n_chunks = (n + CB_PER_BLOCK - 1) / CB_PER_BLOCK;
    // SINK: this batch's records also go to log (rank, slot) on every rank, after one marker record at
    // the log's tail.  The writer-side tail is double-buffered by batch parity: launch `seq` reads
    // state[seq & 1] and (block 0, below) writes state[~seq & 1] for the next launch on this slot's
    // stream, so every block of a launch sees the same tail and nothing waits on anything.
    const unsigned long long tail0 = SINK ? sink.state[sink.seq & 1u] : 0ull;
    bool over = false;
    if (SINK && blockIdx.x == 0) {
        // marker, new tail and the region headers, up front: the batch's record count is the sum of
        // the chunk counts k_extend left
        const uint32_t total = sum_before(chunk_counts, n_chunks, sm);
        if (threadIdx.x == 0) {
            const unsigned long long meta = sink.state[2 + (sink.seq & 1u)];
            uint32_t status = (uint32_t)meta;
            const uint32_t batches = (uint32_t)(meta >> 32) + 1;
            unsigned long long tail = tail0;
            if (tail0 < sink.log_cap) {
                const uint4 marker = make_uint4(AF_LOG_MARKER, (uint32_t)sink.pair_base, (uint32_t)(sink.pair_base >> 32), total);
                for (int r = 0; r < sink.world; r++) ((uint4 *)(sink.region[r] + AF_LOG_HEADER_BYTES))[tail0] = marker;
                tail = tail0 + 1 + total;
            }
            if (tail0 >= sink.log_cap || tail > sink.log_cap || total > hits_cap) {
                if (tail > sink.log_cap) tail = sink.log_cap;
                status |= AF_STATUS_LOG_OVERFLOW;
                atomicOr(&counts[AF_CNT_STATUS], AF_STATUS_LOG_OVERFLOW);
            }
            for (int r = 0; r < sink.world; r++) {
                af_log_header *h = (af_log_header *)sink.region[r];
                h->status = status; h->n_batches = batches; h->tail = tail;
            }
            sink.state[(sink.seq & 1u) ^ 1u] = tail;
            sink.state[2 + ((sink.seq & 1u) ^ 1u)] = ((unsigned long long)batches << 32) | status;
        }
        __syncthreads();
    }
    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const uint32_t mine = chunk_counts[chunk];
        if (mine == 0) continue;
        const uint32_t base = sum_before(chunk_counts, chunk, sm);
        if (threadIdx.x == 0) atomicAdd(&counts[AF_CNT_HITS], mine);
        const uint32_t i0 = chunk * CB_PER_BLOCK + threadIdx.x * CB_ITEMS;
        uint4 v[CB_ITEMS];
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++) {
            v[i] = (i0 + i < n) ? slots[i0 + i] : make_uint4(0, 0, 0, 0);
            c += (v[i].z >> 16) != 0;   // m_len
        }
        uint32_t off = base + block_excl_scan(c, sm);
#pragma unroll
        for (int i = 0; i < CB_ITEMS; i++)
            if ((v[i].z >> 16) != 0) { if (off < hits_cap) hits[off] = v[i]; else over = true; off++; }
        if (SINK) {
            // push this chunk's records (now contiguous in hits[base, base + mine)) to every rank's log:
            // consecutive threads store consecutive records, so a warp's store is 512 contiguous bytes
            // per destination -- full-size NVLink write packets instead of one packet per record
            __syncthreads();
            const uint32_t end = min(base + mine, hits_cap);
            for (uint32_t j = base + threadIdx.x; j < end; j += CB_THREADS) {
                const unsigned long long at = tail0 + 1 + j;
                if (at >= sink.log_cap) break;                 // flagged by block 0
                const uint4 rec = hits[j];
                for (int r = 0; r < sink.world; r++) ((uint4 *)(sink.region[r] + AF_LOG_HEADER_BYTES))[at] = rec;
            }
        }
    }
    if (over) atomicOr(&counts[AF_CNT_STATUS], AF_STATUS_HIT_OVERFLOW);
}

// ------------------------------------------------------------------------------------------
// verify: one THREAD per flagged read.  Removes the filter's false positives cheaply: exact
// k'-mer membership from an L2-resident bitmap -- every sample's bit is loaded before any is
// used, so one thread keeps ~18 loads in flight -- and for the rare members a table walk plus
// a check that the exact match around the sample reaches k bases.  A read is kept iff some
// diagonal holds >= k consecutive matches: exactly the SEEDED predicate of the spec, so
// k_extend only ever sees reads it has to extend.
// ------------------------------------------------------------------------------------------
template <int KP, int WMAX = 16>                        // WMAX = 32: the long-read instance (reads of 257..512 bases)
__global__ void __launch_bounds__(256)
k_verify(const uint32_t *__restrict__ packed, int W, int Q, int uniform_len, const uint16_t *__restrict__ lens,
         const uint32_t *__restrict__ nread_ids, const uint32_t *__restrict__ nmask, int n_nreads,
         const uint32_t *__restrict__ cand, const uint32_t *__restrict__ counts, uint32_t cand_cap,
         const uint32_t *__restrict__ member, const uint2 *__restrict__ table, uint32_t tmask,
         const uint8_t *__restrict__ anchor, int G, int K, uint8_t *__restrict__ keep,
         uint32_t *__restrict__ chunk_counts) {
    constexpr int S = 20 - KP;
    constexpr int NPMAX = af_nsamples(16 * WMAX, KP);        // <= 63, the hit bitmap below is 64 bits
    constexpr uint32_t kpmask = (1u << (2 * KP)) - 1u;
    const uint32_t ncand = min(counts[AF_CNT_FLAGGED], cand_cap);
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count so the ballots below are well defined
    for (uint32_t c0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; c0 < ncand; c0 += gridDim.x * blockDim.x) {
        const uint32_t c = c0 + lane;
        bool seeded = false;
        if (c < ncand) {
            const uint32_t rid = cand[c], pair = rid >> 1;
            ReadRef r;
            r.packed = packed;
            r.base = ((size_t)(pair >> 5) * Q * 32 + (pair & 31)) * 4;
            r.wofs = (int)(rid & 1u) * W;
            r.L = uniform_len > 0 ? uniform_len : (int)lens[rid];
            r.nm = nullptr;
            if (n_nreads > 0) {
                int lo = 0, hi = n_nreads;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (nread_ids[mid] < rid) lo = mid + 1; else hi = mid; }
                if (lo < n_nreads && nread_ids[lo] == rid) r.nm = nmask + (size_t)lo * AF_NMASK_WORDS;
            }
            const int nprobe = af_nsamples(r.L, KP);
            uint32_t w[WMAX + 1];
#pragma unroll
            for (int t = 0; t < WMAX; t++) w[t] = t < W ? r.word(t) : 0u;
            w[WMAX] = 0;
            // phase 1: membership bit of every sample; all loads are independent
            uint32_t mbits[NPMAX];
#pragma unroll
            for (int j = 0; j < NPMAX; j++) {
                const int o = 2 * (af_sample0(KP) + j * S), wi = o >> 5;
                const uint32_t key = __funnelshift_r(w[wi], w[wi + 1], o & 31) & kpmask;
                mbits[j] = j < nprobe ? ((member[key >> 5] >> (key & 31)) & 1u) : 0u;
            }
            unsigned long long hit = 0;
#pragma unroll
            for (int j = 0; j < NPMAX; j++) hit |= (unsigned long long)mbits[j] << j;
            // phase 2: the rare members -- walk the table, check that the exact run reaches K
            while (hit && !seeded) {
                const int j = __ffsll((long long)hit) - 1;
                hit &= hit - 1;
                const int p = af_sample0(KP) + j * S, o = 2 * p;
                const uint32_t key = __funnelshift_r(r.word(o >> 5), (o >> 5) + 1 < W ? r.word((o >> 5) + 1) : 0u, o & 31) & kpmask;
                if (r.nm) {   // a k'-mer that overlaps an N is no seed material
                    bool n = false;
                    for (int t = 0; t < KP; t++) n |= r.is_n(p + t);
                    if (n) continue;
                }
                for (uint32_t slot = af_table_hash(key, tmask);; slot = (slot + 1) & tmask) {
                    const uint2 e = table[slot];
                    if (e.x == AF_T_EMPTY) break;
                    if (e.x != key) continue;
                    const int s = e.y >> 31, jpos = (int)(e.y & 0x7FFFFFFFu);
                    const int qp = s ? r.L - p - KP : p, d = jpos - qp;
                    int run = KP;
                    for (int i = qp - 1; run < K && diag_match(r, s, i, d, anchor, G); i--) run++;
                    for (int i = qp + KP; run < K && diag_match(r, s, i, d, anchor, G); i++) run++;
                    if (run >= K) { seeded = true; break; }
                }
            }
            keep[c] = seeded ? 1 : 0;
        }
        const uint32_t bal = __ballot_sync(FULL, seeded);
        if (lane == 0 && bal) atomicAdd(&chunk_counts[c0 / CB_PER_BLOCK], __popc(bal));
    }
}


// ------------------------------------------------------------------------------------------
// verify, shared-memory variant (default).  k_verify above spends its time in L1TEX: ~36 sector
// requests per flagged read (ncu: 13 M sectors, 0.8 per cycle per SM), half of them the
// membership gathers.  Here one persistent 1024-thread CTA per SM stages a HALF-SIZE copy of the
// anchor filter next to the candidates' words, so "which samples are anchor k'-mers" costs no
// global traffic at all; only those samples (1.3 per flagged read) go to the exact table, and the
// >= k run is checked word-parallel against the 2-bit packed anchor: ~11 sector requests per read.
// Same exact SEEDED predicate as k_verify.
// ------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(1024, 1)
k_verify_smem(const uint32_t *__restrict__ packed, int W, int Q, int uniform_len, const uint16_t *__restrict__ lens,
              const uint32_t *__restrict__ nread_ids, const uint32_t *__restrict__ nmask, int n_nreads,
              const uint32_t *__restrict__ cand, const uint32_t *__restrict__ counts, uint32_t cand_cap,
              const uint32_t *__restrict__ g_filter, uint32_t fmul, uint32_t nb, const uint2 *__restrict__ table,
              uint32_t tmask, const uint8_t *__restrict__ anchor, const uint32_t *__restrict__ apk0,
              const uint32_t *__restrict__ apk1, int anchor_has_n, int G, int K, uint8_t *__restrict__ keep,
              uint32_t *__restrict__ chunk_counts) {
    constexpr int S = 20 - KP, FL = 7;                      // flank bases needed on a side: k - k' <= 7
    constexpr uint32_t kpmask = (1u << (2 * KP)) - 1u;
    extern __shared__ __align__(128) uint32_t vsm[];
    uint32_t *filt = vsm, *swb = vsm + nb;                  // word k of this thread's read at swb[k*VT + tid]
    const int VT = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    const uint32_t ncand = min(counts[AF_CNT_FLAGGED], cand_cap);
    if ((uint32_t)blockIdx.x * VT >= ncand) return;         // whole CTA idle: skip the staging too
    stage_filter(filt, g_filter, nb);
    __syncthreads();
    uint32_t *sw = swb + tid;
    // warp-uniform trip count so the ballot below is well defined
    for (uint32_t c0 = (blockIdx.x * VT + tid) & ~31u; c0 < ncand; c0 += gridDim.x * VT) {
        const uint32_t c = c0 + lane;
        bool seeded = false;
        if (c < ncand) {
            const uint32_t rid = cand[c], pair = rid >> 1;
            ReadRef r;
            r.packed = packed;
            r.base = ((size_t)(pair >> 5) * Q * 32 + (pair & 31)) * 4;
            r.wofs = (int)(rid & 1u) * W;
            r.L = uniform_len > 0 ? uniform_len : (int)lens[rid];
            r.nm = nullptr;
            if (n_nreads > 0) {
                int lo = 0, hi = n_nreads;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (nread_ids[mid] < rid) lo = mid + 1; else hi = mid; }
                if (lo < n_nreads && nread_ids[lo] == rid) r.nm = nmask + (size_t)lo * AF_NMASK_WORDS;
            }
            const int np = af_nsamples(r.L, KP);
            // the read's words: whole quads (128-bit loads), then the W words it owns into shared memory
            // Three gathers are issued before the first is used: a 150-base read is three quads, and a loop that
            // stores each quad before it fetches the next one pays three DRAM latencies in a row per candidate
            // (SASS of the one-quad-per-iteration loop: LDG, STS, BRA, LDG, ...).  Measured: 39.2 -> 36.5 us for the stage --
            // most of its time is the NUMBER of scattered loads (29.9 us with one quad per read), not their order.
            const int q0 = r.wofs >> 2, q1 = (r.wofs + W - 1) >> 2;
            for (int q = q0; q <= q1; q += 3) {
                const uint4 *qp = reinterpret_cast<const uint4 *>(packed + r.base + (size_t)q * 128);
                uint4 v[3];
                v[0] = ld_gather_v4(qp);
                v[1] = q + 1 <= q1 ? ld_gather_v4(qp + 32) : make_uint4(0, 0, 0, 0);
                v[2] = q + 2 <= q1 ? ld_gather_v4(qp + 64) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const uint32_t vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int t = 4 * (q + i) + e - r.wofs;
                        if (t >= 0 && t < W) sw[t * VT] = vv[e];
                    }
                }
            }
            sw[W * VT] = 0; sw[(W + 1) * VT] = 0; sw[(W + 2) * VT] = 0;
            unsigned long long hit = 0;
            for (int j = 0; j < np; j++) {                   // which samples pass the shared-memory filter
                const int o = 2 * (af_sample0(KP) + j * S), wi = o >> 5;
                const uint32_t key = __funnelshift_r(sw[wi * VT], sw[(wi + 1) * VT], o & 31) & kpmask;
                uint32_t b, fp3;
                af_filter_hash(key, fmul, nb, b, fp3);
                if (af_filter_test(filt[b], fp3)) hit |= 1ull << j;
            }
            const bool fast = !r.nm && !anchor_has_n;
            while (hit && !seeded) {
                const int j = __ffsll((long long)hit) - 1;
                hit &= hit - 1;
                const int p = af_sample0(KP) + j * S, o = 2 * p;
                const uint32_t key = __funnelshift_r(sw[(o >> 5) * VT], sw[((o >> 5) + 1) * VT], o & 31) & kpmask;
                if (r.nm) {   // a k'-mer that overlaps an N is no seed material
                    bool n = false;
                    for (int t = 0; t < KP; t++) n |= r.is_n(p + t);
                    if (n) continue;
                }
                // Cheap necessary condition before any global memory is touched: a >= K-base exact match around
                // the sample has left and right slack adding up to K - k', so it also contains the k'-mer H bases
                // to the left or the one H bases to the right of the sample.  A chance k'-mer hit (86 % of what
                // the scan flags) passes with probability ~0.3 %; everything else skips the table walk.
                {
                    const int H = (K - KP + 1) >> 1;
                    bool near = false;
                    if (p >= H) {
                        const int o2 = 2 * (p - H);
                        const uint32_t k2 = __funnelshift_r(sw[(o2 >> 5) * VT], sw[((o2 >> 5) + 1) * VT], o2 & 31) & kpmask;
                        uint32_t b2, f2;
                        af_filter_hash(k2, fmul, nb, b2, f2);
                        near = af_filter_test(filt[b2], f2);
                    }
                    if (!near && p + H + KP <= r.L) {
                        const int o2 = 2 * (p + H);
                        const uint32_t k2 = __funnelshift_r(sw[(o2 >> 5) * VT], sw[((o2 >> 5) + 1) * VT], o2 & 31) & kpmask;
                        uint32_t b2, f2;
                        af_filter_hash(k2, fmul, nb, b2, f2);
                        near = af_filter_test(filt[b2], f2);
                    }
                    if (!near) continue;
                }
                for (uint32_t slot = af_table_hash(key, tmask);; slot = (slot + 1) & tmask) {
                    const uint2 e = table[slot];
                    if (e.x == AF_T_EMPTY) break;
                    if (e.x != key) continue;
                    const int s = e.y >> 31, jpos = (int)(e.y & 0x7FFFFFFFu);
                    int run = KP;
                    if (fast) {
                        // read-forward frame: the sample at p sits at js on the forward (s=0) or
                        // reverse-complemented (s=1) anchor; compare 2-bit windows word-parallel
                        const int js = s ? G - jpos - KP : jpos, dd = js - p;
                        const int i0 = max(max(p - FL, 0), -dd), i1 = min(min(p + KP + FL, r.L), G - dd), n = i1 - i0;
                        const int wi = i0 >> 4, sh = 2 * (i0 & 15);
                        const uint32_t r0 = sw[wi * VT], r1 = sw[(wi + 1) * VT], r2 = sw[(wi + 2) * VT];
                        const unsigned long long rb = (unsigned long long)__funnelshift_r(r0, r1, sh) |
                                                      ((unsigned long long)__funnelshift_r(r1, r2, sh) << 32);
                        const unsigned long long x = rb ^ packed_window(s ? apk1 : apk0, i0 + dd);
                        unsigned long long ne = (x | (x >> 1)) & 0x5555555555555555ull;   // 1 = bases differ
                        ne |= 0x5555555555555555ull << (2 * n);                             // past the overlap
                        const int a = p - i0;
                        const unsigned long long lm = ne & ((1ull << (2 * a)) - 1ull), rm = ne >> (2 * (a + KP));
                        run += lm ? a - 1 - ((63 - __clzll((long long)lm)) >> 1) : a;
                        run += (__ffsll((long long)rm) - 1) >> 1;                           // rm != 0: n < 32
                    } else {
                        const int qp = s ? r.L - p - KP : p, d = jpos - qp;
                        for (int i = qp - 1; run < K && diag_match(r, s, i, d, anchor, G); i--) run++;
                        for (int i = qp + KP; run < K && diag_match(r, s, i, d, anchor, G); i++) run++;
                    }
                    if (run >= K) { seeded = true; break; }
                }
            }
            keep[c] = seeded ? 1 : 0;
        }
        const uint32_t bal = __ballot_sync(FULL, seeded);
        if (lane == 0 && bal) atomicAdd(&chunk_counts[c0 / CB_PER_BLOCK], __popc(bal));
    }
}

// ------------------------------------------------------------------------------------------
// fused seed scan + verify: warp-specialised persistent kernel
//
// 24 warps per CTA: 20 SCAN warps run exactly the probe sequence of k_seed_scan; a flagged lane
// pushes (read_id, its packed words -- already in registers) into its warp's single-producer
// queue in shared memory.  4 VERIFY warps, each serving five of those queues, pop up to 32
// candidates at a time, re-probe their samples in the shared-memory filter to learn which
// samples hit (all 32 lanes busy), walk the exact table for those and check the >= k run
// word-parallel against the 2-bit packed anchor.  Scan warps keep the LSU busy while verify
// warps wait on L2: the latency-bound stage hides under the bandwidth-bound one, the flagged
// reads never go back to HBM, and two compaction passes disappear.  Output: seeded flag words in
// the same (tile, mate) ballot layout the scan uses, plus per-chunk counts -> k_flag_scatter.
// ------------------------------------------------------------------------------------------
#ifndef AF_FZ_SCAN_WARPS
#define AF_FZ_SCAN_WARPS 20
#endif
#ifndef AF_FZ_VERIFY_WARPS
#define AF_FZ_VERIFY_WARPS 4
#endif
#ifndef AF_FZ_QCAP
#define AF_FZ_QCAP 16
#endif
static const int FZ_SCAN_WARPS = AF_FZ_SCAN_WARPS, FZ_VERIFY_WARPS = AF_FZ_VERIFY_WARPS, FZ_QPC = FZ_SCAN_WARPS / FZ_VERIFY_WARPS,
                 FZ_QCAP = AF_FZ_QCAP, FZ_THREADS = 32 * (FZ_SCAN_WARPS + FZ_VERIFY_WARPS);
static_assert(FZ_SCAN_WARPS % FZ_VERIFY_WARPS == 0 && FZ_QPC <= 8, "each verify warp serves FZ_QPC <= 8 scan warps");

__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// push the flagged lanes' reads (registers w[OFF..OFF+W)) into this scan warp's queue
template <int W, int OFF, int NW>
__device__ __forceinline__ void fz_push(uint32_t mrem, const uint32_t (&w)[NW], uint32_t rid, int lane, uint32_t *qdata,
                                        uint32_t *qhead, uint32_t *qtail, uint32_t &tail) {
    constexpr int ES = W + 1;
    while (mrem) {                                           // warp-uniform
        const int room = FZ_QCAP - (int)(tail - ld_vol(qhead));
        if (room <= 0) { __nanosleep(40); continue; }
        const bool mine = (mrem >> lane) & 1u;
        const int rank = __popc(mrem & ((1u << lane) - 1u));
        const bool go = mine && rank < room;
        if (go) {
            uint32_t *e = qdata + ((tail + rank) % FZ_QCAP) * ES;
            e[0] = rid;
#pragma unroll
            for (int t = 0; t < W; t++) e[1 + t] = w[OFF + t];
        }
        const uint32_t taken = __ballot_sync(FULL, go);
        mrem &= ~taken;
        __threadfence_block();
        __syncwarp();
        tail += __popc(taken);
        if (lane == 0) st_vol(qtail, tail);
    }
}

template <int W, int KP>
__global__ void __launch_bounds__(FZ_THREADS, 1)
k_scan_verify(const uint4 *__restrict__ packed, long long n_tiles, long long n_pairs, int nprobe, int uniform_len,
              const uint16_t *__restrict__ lens, const uint32_t *__restrict__ nread_ids,
              const uint32_t *__restrict__ nmask, int n_nreads, const uint32_t *__restrict__ g_filter, uint32_t fmul,
              uint32_t nb, const uint2 *__restrict__ table, uint32_t tmask, const uint8_t *__restrict__ anchor,
              const uint32_t *__restrict__ apk0, const uint32_t *__restrict__ apk1, int anchor_has_n, int G, int K,
              uint32_t *__restrict__ seeded_flags, uint32_t *__restrict__ chunk_counts, uint32_t *__restrict__ counts) {
    extern __shared__ __align__(128) uint32_t fsm[];
    constexpr int Q = (2 * W + 3) / 4, ES = W + 1, SWW = W + 3, S = 20 - KP;
    constexpr uint32_t kpmask = (1u << (2 * KP)) - 1u;
    uint32_t *filt = fsm;
    uint32_t *qdata = filt + nb;                                         // [FZ_SCAN_WARPS][FZ_QCAP][ES]
    uint32_t *qhead = qdata + FZ_SCAN_WARPS * FZ_QCAP * ES;              // [FZ_SCAN_WARPS]
    uint32_t *qtail = qhead + FZ_SCAN_WARPS, *qdone = qtail + FZ_SCAN_WARPS;
    uint32_t *scratch = qdone + FZ_SCAN_WARPS;                           // [FZ_VERIFY_WARPS][SWW][32]
    uint32_t *nflag = scratch + FZ_VERIFY_WARPS * SWW * 32;              // [1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t_begin = n_tiles * blockIdx.x / gridDim.x, t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x < 3 * FZ_SCAN_WARPS) qhead[threadIdx.x] = 0;         // heads, tails, done flags
    if (threadIdx.x == 0) *nflag = 0;
    uint32_t wa[4 * Q], wb[4 * Q];
    long long tile = t_begin + warp;
    const bool scanner = warp < FZ_SCAN_WARPS;
    if (scanner && tile < t_end) load_tile<Q>(wa, packed, tile, lane);   // in flight while the filter is staged
    stage_filter_ldst(filt, g_filter, nb);    // this kernel's queues fill shared memory to the last byte: no room for an mbarrier
    __syncthreads();
    if (scanner) {
        uint32_t tail = 0, flagged = 0;
        uint32_t *qd = qdata + warp * FZ_QCAP * ES;
        auto scan_and_push = [&](const uint32_t (&w)[4 * Q], long long tl) {
            const uint32_t a1 = af_scan_read<W, KP, 0, 4 * Q>(w, nprobe, filt, fmul, nb);
            const uint32_t a2 = af_scan_read<W, KP, W, 4 * Q>(w, nprobe, filt, fmul, nb);
            const uint32_t vm = tile_valid_mask(tl, n_pairs);
            const uint32_t b1 = __ballot_sync(FULL, a1 != 0) & vm, b2 = __ballot_sync(FULL, a2 != 0) & vm;
            const uint32_t rid = (uint32_t)(tl * 32 + lane) * 2u;
            if (b1) fz_push<W, 0, 4 * Q>(b1, w, rid, lane, qd, qhead + warp, qtail + warp, tail);
            if (b2) fz_push<W, W, 4 * Q>(b2, w, rid + 1u, lane, qd, qhead + warp, qtail + warp, tail);
            flagged += __popc(b1) + __popc(b2);
        };
        while (tile < t_end) {
            const long long t2 = tile + FZ_SCAN_WARPS;
            if (t2 < t_end) load_tile<Q>(wb, packed, t2, lane);
            scan_and_push(wa, tile);
            if (t2 >= t_end) break;
            const long long t3 = t2 + FZ_SCAN_WARPS;
            if (t3 < t_end) load_tile<Q>(wa, packed, t3, lane);
            scan_and_push(wb, t2);
            tile = t3;
        }
        __threadfence_block();
        if (lane == 0) { st_vol(qdone + warp, 1u); if (flagged) atomicAdd(nflag, flagged); }
    } else {
        const int c = warp - FZ_SCAN_WARPS, q0 = c * FZ_QPC;
        uint32_t *sw = scratch + c * SWW * 32 + lane;                    // word k of this lane's candidate at sw[k*32]
        uint32_t hq = 0;                                                 // lanes < FZ_QPC: head of queue q0+lane
        for (;;) {
            uint32_t avail = 0;
            if (lane < FZ_QPC) avail = ld_vol(qtail + q0 + lane) - hq;
            uint32_t pre = avail;                                        // inclusive prefix over the five queues
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
            const uint32_t total = __shfl_sync(FULL, pre, FZ_QPC - 1);
            // A round is latency bound (dependent L2 loads), so it pays to run it with all 32 lanes:
            // wait for a full warp's worth of candidates unless a queue is about to fill up (its scan
            // warp would stall) or the scan warps are done.
            if (total < 32u) {
                const uint32_t dn = lane < FZ_QPC ? ld_vol(qdone + q0 + lane) : 1u;
                const bool all_done = __all_sync(FULL, dn != 0);
                const bool pressure = __any_sync(FULL, avail >= (uint32_t)(FZ_QCAP - 4));
                if (!all_done && !pressure) { __nanosleep(200); continue; }
                if (total == 0) {
                    const uint32_t again = lane < FZ_QPC ? ld_vol(qtail + q0 + lane) - hq : 0u;   // a push may precede the done flag
                    if (!__any_sync(FULL, again != 0)) break;
                    continue;
                }
            }
            const uint32_t ntake = min(total, 32u);
            __threadfence_block();                                       // entries were written before the tails we just read
            // lane i < ntake takes the i-th available entry, queues in order
            uint32_t rid = 0;
            bool have = false;
#pragma unroll
            for (int q = 0; q < FZ_QPC; q++) {
                const uint32_t incl = __shfl_sync(FULL, pre, q), av = __shfl_sync(FULL, avail, q), h = __shfl_sync(FULL, hq, q);
                const uint32_t start = incl - av;
                if ((uint32_t)lane < ntake && (uint32_t)lane >= start && (uint32_t)lane < incl) {
                    const uint32_t *e = qdata + ((q0 + q) * FZ_QCAP + (h + (lane - start)) % FZ_QCAP) * ES;
                    rid = e[0];
#pragma unroll
                    for (int t = 0; t < W; t++) sw[t * 32] = e[1 + t];
                    have = true;
                }
            }
            if (have) { sw[W * 32] = 0; sw[(W + 1) * 32] = 0; sw[(W + 2) * 32] = 0; }
            __threadfence_block();
            __syncwarp();
            if (lane < FZ_QPC) {                                         // release what was copied out
                const uint32_t start = pre - avail;
                const uint32_t took = ntake > start ? min(avail, ntake - start) : 0u;
                hq += took;
                if (took) st_vol(qhead + q0 + lane, hq);
            }
            // ---- verify this lane's candidate --------------------------------------------
            bool seeded = false;
            if (have) {
                const uint32_t pair = rid >> 1;
                ReadRef r;                                              // global view, used only by the slow (N) path
                r.packed = reinterpret_cast<const uint32_t *>(packed);
                r.base = ((size_t)(pair >> 5) * Q * 32 + (pair & 31)) * 4;
                r.wofs = (int)(rid & 1u) * W;
                r.L = uniform_len > 0 ? uniform_len : (int)lens[rid];
                r.nm = nullptr;
                if (n_nreads > 0) {
                    int lo = 0, hi = n_nreads;
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (nread_ids[mid] < rid) lo = mid + 1; else hi = mid; }
                    if (lo < n_nreads && nread_ids[lo] == rid) r.nm = nmask + (size_t)lo * AF_NMASK_WORDS;
                }
                const int np = af_nsamples(r.L, KP);
                unsigned long long hit = 0;
                for (int j = 0; j < np; j++) {                           // which samples pass the filter
                    const int o = 2 * (af_sample0(KP) + j * S), wi = o >> 5;
                    const uint32_t key = __funnelshift_r(sw[wi * 32], sw[(wi + 1) * 32], o & 31) & kpmask;
                    uint32_t b, fp3;
                    af_filter_hash(key, fmul, nb, b, fp3);
                    if (af_filter_test(filt[b], fp3)) hit |= 1ull << j;
                }
                const bool fast = !r.nm && !anchor_has_n;
                while (hit && !seeded) {
                    const int j = __ffsll((long long)hit) - 1;
                    hit &= hit - 1;
                    const int p = af_sample0(KP) + j * S, o = 2 * p;
                    const uint32_t key = __funnelshift_r(sw[(o >> 5) * 32], sw[((o >> 5) + 1) * 32], o & 31) & kpmask;
                    if (r.nm) {   // a k'-mer that overlaps an N is no seed material
                        bool n = false;
                        for (int t = 0; t < KP; t++) n |= r.is_n(p + t);
                        if (n) continue;
                    }
                    for (uint32_t slot = af_table_hash(key, tmask);; slot = (slot + 1) & tmask) {
                        const uint2 e = table[slot];
                        if (e.x == AF_T_EMPTY) break;
                        if (e.x != key) continue;
                        const int s = e.y >> 31, jpos = (int)(e.y & 0x7FFFFFFFu);
                        int run = KP;
                        if (fast) {
                            // read-forward frame: the sample at p sits at js on the forward (s=0) or
                            // reverse-complemented (s=1) anchor; compare 2-bit windows word-parallel
                            constexpr int FL = 7;
                            const int js = s ? G - jpos - KP : jpos, dd = js - p;
                            const int i0 = max(max(p - FL, 0), -dd), i1 = min(min(p + KP + FL, r.L), G - dd), n = i1 - i0;
                            const int wi = i0 >> 4, sh = 2 * (i0 & 15);
                            const uint32_t r0 = sw[wi * 32], r1 = sw[(wi + 1) * 32], r2 = sw[(wi + 2) * 32];
                            const unsigned long long rb = (unsigned long long)__funnelshift_r(r0, r1, sh) |
                                                          ((unsigned long long)__funnelshift_r(r1, r2, sh) << 32);
                            const unsigned long long x = rb ^ packed_window(s ? apk1 : apk0, i0 + dd);
                            unsigned long long ne = (x | (x >> 1)) & 0x5555555555555555ull;   // 1 = bases differ
                            ne |= 0x5555555555555555ull << (2 * n);                             // past the overlap
                            const int a = p - i0;
                            const unsigned long long lm = ne & ((1ull << (2 * a)) - 1ull), rm = ne >> (2 * (a + KP));
                            run += lm ? a - 1 - ((63 - __clzll((long long)lm)) >> 1) : a;
                            run += (__ffsll((long long)rm) - 1) >> 1;                           // rm != 0: n < 32
                        } else {
                            const int qp = s ? r.L - p - KP : p, d = jpos - qp;
                            for (int i = qp - 1; run < K && diag_match(r, s, i, d, anchor, G); i--) run++;
                            for (int i = qp + KP; run < K && diag_match(r, s, i, d, anchor, G); i++) run++;
                        }
                        if (run >= K) { seeded = true; break; }
                    }
                }
                if (seeded) {
                    const uint32_t tl = pair >> 5;
                    atomicOr(&seeded_flags[tl * 2 + (rid & 1u)], 1u << (pair & 31));
                    atomicAdd(&chunk_counts[tl / CB_PER_BLOCK], 1u);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && *nflag) atomicAdd(&counts[AF_CNT_FLAGGED], *nflag);
}

// ------------------------------------------------------------------------------------------
// extend: one warp per seeded read
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_extend(const uint32_t *__restrict__ packed, int W, int Q, int uniform_len, const uint16_t *__restrict__ lens,
         const uint32_t *__restrict__ nread_ids, const uint32_t *__restrict__ nmask, int n_nreads,
         const uint32_t *__restrict__ cand, const uint32_t *__restrict__ counts, uint32_t cand_cap,
         const uint2 *__restrict__ table, uint32_t tmask, const uint8_t *__restrict__ anchor, int G, int KP, int S,
         ExtParams P, uint4 *__restrict__ slots, uint32_t *__restrict__ chunk_counts, uint32_t *__restrict__ next_read,
         const uint32_t *__restrict__ apk0p, const uint32_t *__restrict__ apk1p, const uint32_t *__restrict__ apn0p,
         const uint32_t *__restrict__ apn1p, int walk) {
    const int lane = threadIdx.x & 31;
    const uint32_t ncand = min(counts[AF_CNT_SEEDED], cand_cap);
    const uint32_t kpmask = (1u << (2 * KP)) - 1u;
    // reads differ a lot in cost (number of diagonals, extension length): warps take the next read from a
    // counter instead of a fixed stride, so no SM is left with the long tail (ncu: active cycles 44k avg /
    // 62k max per SM with the static assignment)
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(next_read, 1u);
        c = __shfl_sync(FULL, c, 0);
        if (c >= ncand) break;
        const uint32_t rid = cand[c];
        const uint32_t pair = rid >> 1, mate = rid & 1u;
        const int L = uniform_len > 0 ? uniform_len : (int)lens[rid];
        // the read's packed words: lane t < W holds word t (tile-interleaved layout)
        uint32_t rw = 0;
        if (lane < W) {
            const uint32_t wi = mate * W + lane;
            rw = packed[(((size_t)(pair >> 5) * Q + (wi >> 2)) * 32 + (pair & 31)) * 4 + (wi & 3)];
        }
        // N mask: binary search of the sorted N-read list (uniform across the warp)
        uint32_t nwv = 0;
        bool has_n = false;
        if (n_nreads > 0) {
            int lo = 0, hi = n_nreads;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (nread_ids[mid] < rid) lo = mid + 1; else hi = mid; }
            if (lo < n_nreads && nread_ids[lo] == rid) { has_n = true; if (lane < AF_NMASK_WORDS) nwv = nmask[(size_t)lo * AF_NMASK_WORDS + lane]; }
        }
        int best_sc = -1, best_qb = 0, best_qe = 0;
        uint32_t best_key = 0xFFFFFFFFu;
        const int nprobe = af_nsamples(L, KP);                                                    // S == 20 - KP
        for (int p0 = 0; p0 < nprobe; p0 += 32) {
            const int pi = p0 + lane, p = af_sample0(KP) + pi * S;
            bool active = pi < nprobe;
            // this lane's k'-mer (forward read orientation)
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
            // advance to this lane's next table entry with the same key
            auto next_match = [&]() {
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
                    // walk != 0 (default): word-parallel match mask + an extension that steps from mismatch to mismatch
                    // (eval_mask_walk); 0: base-by-base mask and 32-step warp scans (eval_diag), kept for A/B runs
                    const int sc = walk ? eval_mask_walk(diag_mask_wp(s, d, L, rw, nwv, has_n, apk0p, apk1p, apn0p, apn1p, G, lane), d, L, G, P, lane, qb, qe)
                                        : eval_diag(s, d, L, rw, nwv, has_n, anchor, G, P, lane, qb, qe);
                    if (sc > best_sc || (sc == best_sc && sc >= 0 && dk < best_key)) { best_sc = sc; best_qb = qb; best_qe = qe; best_key = dk; }
                }
                if (found) next_match();
            }
        }
        if (lane == 0) {
            uint4 out = make_uint4(rid, 0, 0, 0);
            if (best_sc >= P.T) {
                const int s = best_key >> 31, d = (int)(best_key & 0x7FFFFFFFu) - 1024;
                out.y = (uint32_t)(best_qb + d + 1);
                out.z = (uint32_t)best_qb | ((uint32_t)(best_qe - best_qb) << 16);
                out.w = (uint32_t)(L - best_qe) | ((uint32_t)(best_sc * 2 + s) << 16);
                atomicAdd(&chunk_counts[c / CB_PER_BLOCK], 1u);
            }
            slots[c] = out;
        }
    }
}

// debug hook (tools/tail_timing.py): a device buffer of num_sms x 8 int64 that k_tail fills with the phase
// boundaries (cycles since CTA start) of each CTA's first region, or NULL to switch it off; not part of the ABI header
extern "C" void af_debug_tail_timing(long long *d_buf) { g_tail_dbg = d_buf; }

template <int W, int KP>
static int launch_fused(const af_dev_index *d, const af_batch_t *b, long long n_tiles, int nprobe, uint32_t *flags,
                        uint32_t *cc, uint32_t *counts, cudaStream_t st) {
    const size_t smem = ((size_t)d->nb + FZ_SCAN_WARPS * FZ_QCAP * (W + 1) + 3 * FZ_SCAN_WARPS + FZ_VERIFY_WARPS * (W + 3) * 32 + 1) * 4;
    static bool attr_set[64] = {false};  // per device
    static size_t max_dyn = 0;
    if (!attr_set[d->device & 63]) {
        int rc = allow_full_smem(k_scan_verify<W, KP>, &max_dyn);
        if (rc) return rc;
        attr_set[d->device & 63] = true;
    }
    if (smem > max_dyn) { af_set_error("fused kernel: %zu bytes of shared memory needed, %zu available", smem, max_dyn); return AF_ERR_ARG; }
    long long want = (n_tiles + FZ_SCAN_WARPS - 1) / FZ_SCAN_WARPS;
    int grid = (int)(want < d->num_sms ? (want > 0 ? want : 1) : d->num_sms);
    k_scan_verify<W, KP><<<grid, FZ_THREADS, smem, st>>>((const uint4 *)b->packed, n_tiles, b->n_pairs, nprobe, b->uniform_len, b->lens,
                                                  b->nread_ids, b->nmask, (int)b->n_nreads, d->d_filter, d->fmul, d->nb,
                                                  d->d_table, d->tmask, d->d_anchor, d->d_apk[0], d->d_apk[1], d->anchor_has_n,
                                                  d->G, d->P.k, flags, cc, counts);
    g_launches++;
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}

#define AF_FUSED_CASE(WW)                                                                               \
    case WW:                                                                                            \
        return d->kp == 12 ? launch_fused<WW, 12>(d, b, lay.n_tiles, nprobe, flags, cc, counts, st)      \
                           : launch_fused<WW, 13>(d, b, lay.n_tiles, nprobe, flags, cc, counts, st);

static int fused_impl(const af_dev_index *d, const af_batch_t *b, const af_layout_t &lay, uint32_t *flags, uint32_t *cc,
                      uint32_t *counts, cudaStream_t st) {
    if (d->P.k != 19 || (d->kp != 12 && d->kp != 13)) { af_set_error("seed scan is built for k=19, k' in {12,13}"); return AF_ERR_ARG; }
    const int L = b->uniform_len > 0 ? b->uniform_len : b->max_read_len;
    const int nprobe = af_nsamples(L, d->kp);
    switch (lay.words_per_read) {
        AF_FUSED_CASE(1) AF_FUSED_CASE(2) AF_FUSED_CASE(3) AF_FUSED_CASE(4) AF_FUSED_CASE(5) AF_FUSED_CASE(6)
        AF_FUSED_CASE(7) AF_FUSED_CASE(8) AF_FUSED_CASE(9) AF_FUSED_CASE(10) AF_FUSED_CASE(11) AF_FUSED_CASE(12)
        AF_FUSED_CASE(13) AF_FUSED_CASE(14) AF_FUSED_CASE(15) AF_FUSED_CASE(16)
    }
    af_set_error("unsupported words_per_read %d", lay.words_per_read);
    return AF_ERR_ARG;
}

// ------------------------------------------------------------------------------------------
// the hot path on one GPU
// ------------------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WsLayout { size_t flags, cc, cand, keep, cand2, slots, total, cc_bytes; uint32_t nch1, nch2;
                  // candidate-stream path (k_seed_scan<EMIT> + k_tail); shares the bytes behind `cc` with the lists of the six-kernel path
                  size_t ctl, ctl_bytes, dir, chunk_hits, recs; uint32_t r_max, pool_chunks; int rq; };

static WsLayout ws_layout(long long n_pairs, long long cand_cap, int max_read_len) {
    WsLayout w;
    long long n_tiles = (n_pairs + 31) / 32;
    w.nch1 = (uint32_t)((n_tiles + CB_PER_BLOCK - 1) / CB_PER_BLOCK);
    w.nch2 = (uint32_t)((cand_cap + CB_PER_BLOCK - 1) / CB_PER_BLOCK);
    size_t o = 0;
    w.flags = o; o += align256((size_t)n_tiles * 8);
    w.cc = o;                                            // chunk counts of the three compactions, zeroed per call
    w.cc_bytes = align256(((size_t)w.nch1 + 2 * (size_t)w.nch2 + 3) * 4);
    o += w.cc_bytes;
    const size_t u0 = o;
    w.cand = o; o += align256((size_t)cand_cap * 4);
    w.keep = o; o += align256((size_t)cand_cap);
    w.cand2 = o; o += align256((size_t)cand_cap * 4);
    w.slots = o; o += align256((size_t)cand_cap * 16);
    const size_t classic_end = o;
    // candidate stream: R <= n_tiles / AF_REG_TILES + 1 + #SMs regions (R = scan grid x regions per CTA); every
    // (scan warp, region) pair may leave one chunk partly filled
    o = u0;
    w.r_max = (uint32_t)(n_tiles / AF_REG_TILES + 1 + 256);
    w.pool_chunks = (uint32_t)((cand_cap + AF_CHUNK - 1) / AF_CHUNK + 24ll * w.r_max + 1);
    w.rq = af_rec_quads(max_read_len > 256 ? 16 : (max_read_len + 15) / 16);   // reads beyond 256 bases never take the candidate-stream path
    w.ctl = o; w.ctl_bytes = align256((16 + 2 * (size_t)w.r_max) * 4); o += w.ctl_bytes;   // [0] pool counter, [16..) dir_count[R], region_state[R]: zeroed per call
    w.dir = o; o += align256((size_t)w.r_max * AF_DIR_CAP * 4);
    w.chunk_hits = o; o += align256((size_t)w.pool_chunks * 4);
    w.recs = o; o += align256((size_t)w.pool_chunks * AF_CHUNK * w.rq * 16);
    w.total = o > classic_end ? o : classic_end;
    return w;
}

extern "C" size_t af_workspace_bytes_len(int64_t n_pairs, int64_t cand_cap, int32_t max_read_len) {
    if (n_pairs < 0 || cand_cap < 0 || max_read_len < 1 || max_read_len > AF_MAX_READ_LEN) return 0;
    return ws_layout(n_pairs, cand_cap, max_read_len).total + 256;
}

extern "C" size_t af_workspace_bytes(int64_t n_pairs, int64_t cand_cap) {
    return af_workspace_bytes_len(n_pairs, cand_cap, AF_MAX_READ_LEN);
}

static int anchor_batch_impl(const af_dev_index_t *d, const af_batch_t *b, void *workspace, size_t workspace_bytes,
                             int64_t cand_cap, af_hit_t *d_hits, int64_t hits_cap, uint32_t *d_counts, const af_sink *sink,
                             void *stream) {
    af_layout_t lay;
    int rc = batch_check(d, b, &lay);
    if (rc) return rc;
    if (!workspace || !d_counts || (hits_cap > 0 && !d_hits) || cand_cap <= 0 || cand_cap >= (1ll << 32) || hits_cap < 0 || hits_cap >= (1ll << 32)) {
        af_set_error("af_anchor_batch: bad buffers or capacities");
        return AF_ERR_ARG;
    }
    if (workspace_bytes < af_workspace_bytes_len(b->n_pairs, cand_cap, b->max_read_len)) { af_set_error("af_anchor_batch: workspace too small"); return AF_ERR_CAPACITY; }
    AF_CUDA(cudaSetDevice(d->device));
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    WsLayout w = ws_layout(b->n_pairs, cand_cap, b->max_read_len);
    uint32_t *flags = (uint32_t *)(ws + w.flags), *cand = (uint32_t *)(ws + w.cand), *cand2 = (uint32_t *)(ws + w.cand2);
    uint32_t *cc1 = (uint32_t *)(ws + w.cc), *cc2 = cc1 + w.nch1 + 1, *cc3 = cc2 + w.nch2 + 1;
    uint8_t *keep = (uint8_t *)(ws + w.keep);
    uint4 *slots = (uint4 *)(ws + w.slots);
    AF_CUDA(cudaMemsetAsync(d_counts, 0, AF_N_COUNTS * sizeof(uint32_t), st));
    if (lay.n_tiles == 0) return AF_OK;
    const bool long_reads = lay.words_per_read > 16;        // 257..512 bases: plain scan, bitmap verify, base-by-base extension masks
    if (g_stream && !g_fused && g_middle == 7 && !d->saturated && !long_reads) {
        // candidate-stream path: the scan emits the flagged reads, k_tail does the rest -- 2 kernels
        const int grid = scan_grid(d, lay.n_tiles);
        const int m = (int)((lay.n_tiles + (long long)grid * AF_REG_TILES - 1) / ((long long)grid * AF_REG_TILES));
        const uint32_t R = (uint32_t)grid * (uint32_t)m;
        if (R > w.r_max) { af_set_error("af_anchor_batch: %u regions, workspace laid out for %u", R, w.r_max); return AF_ERR_ARG; }
        uint32_t *ctl = (uint32_t *)(ws + w.ctl);
        AF_CUDA(cudaMemsetAsync(ctl, 0, w.ctl_bytes, st));
        af_emit E;
        E.recs = (uint32_t *)(ws + w.recs); E.pool = ctl; E.pool_chunks = w.pool_chunks;
        E.dir_count = ctl + 16; E.dir = (uint32_t *)(ws + w.dir); E.counts = d_counts; E.m = m;
        cudaEvent_t ev;
        prof_mark(&ev, st);
        rc = seed_scan_impl(d, b, nullptr, nullptr, d_counts, false, &E, st);
        if (rc) return rc;
        prof_span(ev, st, ST_SCAN);
        prof_mark(&ev, st);
        af_tail_args a;
        a.recs = E.recs; a.rq = w.rq; a.dir_count = E.dir_count; a.dir = E.dir; a.n_regions = R; a.n_tiles = lay.n_tiles;
        a.chunk_hits = (uint32_t *)(ws + w.chunk_hits); a.region_state = ctl + 16 + w.r_max;
        a.packed = (const uint32_t *)b->packed; a.W = lay.words_per_read; a.Q = lay.quads_per_pair; a.uniform_len = b->uniform_len;
        a.lens = b->lens; a.nread_ids = b->nread_ids; a.nmask = b->nmask; a.n_nreads = (int)b->n_nreads;
        a.g_filter = d->d_filter2; a.fmul = d->fmul2; a.nb = d->nb2; a.table = d->d_table; a.tmask = d->tmask;
        a.anchor = d->d_anchor; a.apk0p = d->d_apkp[0]; a.apk1p = d->d_apkp[1]; a.apn0p = d->d_apn[0]; a.apn1p = d->d_apn[1]; a.apk_words = d->apk_words; a.anchor_has_n = d->anchor_has_n; a.G = d->G;
        a.anchor_in_smem = 0;
        a.P = {d->P.k, d->P.A, d->P.B, d->P.clip5, d->P.clip3, d->P.T, d->P.X};
        a.hits = (uint4 *)d_hits; a.hits_cap = (uint32_t)hits_cap; a.cand_cap = (uint32_t)cand_cap; a.counts = d_counts;
        a.dbg = g_tail_dbg;
        rc = af_tail_launch(d, a, sink, st);
        if (rc) return rc;
        prof_span(ev, st, ST_VERIFY);
        return AF_OK;
    }
    AF_CUDA(cudaMemsetAsync(cc1, 0, w.cc_bytes, st));
    const int scatter_grid = d->num_sms * 4;
    const int sg2 = (int)(w.nch2 < (uint32_t)scatter_grid ? w.nch2 : scatter_grid);
    cudaEvent_t ev;
    prof_mark(&ev, st);
    if (g_fused && !d->bloom && !long_reads) {
        // fused: seed scan + verify in one warp-specialised kernel -> seeded flag words
        AF_CUDA(cudaMemsetAsync(flags, 0, (size_t)lay.n_tiles * 8, st));
        rc = fused_impl(d, b, lay, flags, cc1, d_counts, st);
        if (rc) return rc;
        prof_span(ev, st, ST_SCAN);
        prof_mark(&ev, st);
        k_flag_scatter<<<(int)(w.nch1 < (uint32_t)scatter_grid ? w.nch1 : scatter_grid), CB_THREADS, 0, st>>>(
            (const uint2 *)flags, lay.n_tiles, cc1, w.nch1, cand2, (uint32_t)cand_cap, d_counts, AF_CNT_SEEDED, -1);
        prof_span(ev, st, ST_COMPACT1);
        prof_mark(&ev, st);
        g_launches -= 2;                                    // this path has 4 kernels, the code below counts 5 more
    } else {
    const bool rq = g_middle == 11 && !d->saturated && !long_reads;
    rc = seed_scan_impl(d, b, flags, cc1, d_counts, rq, nullptr, st);
    if (rc) return rc;
    prof_span(ev, st, ST_SCAN);
    prof_mark(&ev, st);
    if (rq) {
        // the flag words already hold the refined survivors: one compaction, then k_extend
        k_flag_scatter<<<(int)(w.nch1 < (uint32_t)scatter_grid ? w.nch1 : scatter_grid), CB_THREADS, 0, st>>>(
            (const uint2 *)flags, lay.n_tiles, cc1, w.nch1, cand2, (uint32_t)cand_cap, d_counts, AF_CNT_SEEDED, -1);
        prof_span(ev, st, ST_COMPACT1);
        prof_mark(&ev, st);
        g_launches -= 2;                                    // 4 kernels on this path
    } else
    if (g_middle == 0) {
        // the flagged reads go straight to k_extend (a flagged read without a seeded diagonal yields no record)
        k_flag_scatter<<<(int)(w.nch1 < (uint32_t)scatter_grid ? w.nch1 : scatter_grid), CB_THREADS, 0, st>>>(
            (const uint2 *)flags, lay.n_tiles, cc1, w.nch1, cand2, (uint32_t)cand_cap, d_counts, AF_CNT_SEEDED, AF_CNT_FLAGGED);
        prof_span(ev, st, ST_COMPACT1);
        prof_mark(&ev, st);
        g_launches -= 2;                                    // 4 kernels on this path
    } else {
    k_flag_scatter<<<(int)(w.nch1 < (uint32_t)scatter_grid ? w.nch1 : scatter_grid), CB_THREADS, 0, st>>>(
        (const uint2 *)flags, lay.n_tiles, cc1, w.nch1, cand, (uint32_t)cand_cap, d_counts, AF_CNT_FLAGGED, -1);
    prof_span(ev, st, ST_COMPACT1);
    prof_mark(&ev, st);
    {
    // the shared-memory variant needs its half-size filter plus (W + 3) x 1024 words: fits up to W = 28 (448 bases)
    const bool vsmem_fits = ((size_t)d->nb2 + (size_t)(lay.words_per_read + 3) * 1024) * 4 + 2048 <= (size_t)227 * 1024;
    if (lay.words_per_read > 16 && !(g_verify_smem && !d->bloom && vsmem_fits)) {   // long reads without room (or a Bloom index): the bitmap kernel with 32 words per read
        long long vthreads = cand_cap < (long long)d->num_sms * 2048 ? cand_cap : (long long)d->num_sms * 2048;
        const unsigned vgrid = (unsigned)((vthreads + 255) / 256);
#define AF_VERIFY_LONG_ARGS (const uint32_t *)b->packed, lay.words_per_read, lay.quads_per_pair, b->uniform_len, b->lens, \
        b->nread_ids, b->nmask, (int)b->n_nreads, cand, d_counts, (uint32_t)cand_cap, d->d_member, d->d_table,         \
        d->tmask, d->d_anchor, d->G, d->P.k, keep, cc2
        if (d->kp == 12) k_verify<12, 32><<<vgrid, 256, 0, st>>>(AF_VERIFY_LONG_ARGS);
        else k_verify<13, 32><<<vgrid, 256, 0, st>>>(AF_VERIFY_LONG_ARGS);
    } else
    if (g_verify_smem && !d->bloom && vsmem_fits) {        // (a Bloom index's half-size fingerprint copy is saturated: k_verify's exact bitmap instead)
        const size_t vsmem = ((size_t)d->nb2 + (size_t)(lay.words_per_read + 3) * 1024) * 4;
        static bool vattr[64][2] = {{false}};
        if (!vattr[d->device & 63][d->kp == 12 ? 0 : 1]) {
            const int rc2 = d->kp == 12 ? allow_full_smem(k_verify_smem<12>, nullptr) : allow_full_smem(k_verify_smem<13>, nullptr);
            if (rc2) return rc2;
            vattr[d->device & 63][d->kp == 12 ? 0 : 1] = true;
        }
        long long vb = (cand_cap + 1023) / 1024;
        const int vsms = g_small_sms > 0 && g_small_sms < d->num_sms ? g_small_sms : d->num_sms;
        const unsigned vg = (unsigned)(vb < vsms ? vb : vsms);
#define AF_VERIFY_SMEM_ARGS (const uint32_t *)b->packed, lay.words_per_read, lay.quads_per_pair, b->uniform_len, b->lens, \
        b->nread_ids, b->nmask, (int)b->n_nreads, cand, d_counts, (uint32_t)cand_cap, d->d_filter2, d->fmul2, d->nb2,       \
        d->d_table, d->tmask, d->d_anchor, d->d_apk[0], d->d_apk[1], d->anchor_has_n, d->G, d->P.k, keep, cc2
        if (d->kp == 12) k_verify_smem<12><<<vg, 1024, vsmem, st>>>(AF_VERIFY_SMEM_ARGS);
        else k_verify_smem<13><<<vg, 1024, vsmem, st>>>(AF_VERIFY_SMEM_ARGS);
    } else {
    long long vthreads = cand_cap < (long long)d->num_sms * 2048 ? cand_cap : (long long)d->num_sms * 2048;
    const unsigned vgrid = (unsigned)((vthreads + 255) / 256);
#define AF_VERIFY_ARGS (const uint32_t *)b->packed, lay.words_per_read, lay.quads_per_pair, b->uniform_len, b->lens, \
        b->nread_ids, b->nmask, (int)b->n_nreads, cand, d_counts, (uint32_t)cand_cap, d->d_member, d->d_table,         \
        d->tmask, d->d_anchor, d->G, d->P.k, keep, cc2
    if (d->kp == 12) k_verify<12><<<vgrid, 256, 0, st>>>(AF_VERIFY_ARGS);
    else k_verify<13><<<vgrid, 256, 0, st>>>(AF_VERIFY_ARGS);
    }
    }
    k_sel_scatter<<<sg2, CB_THREADS, 0, st>>>(cand, keep, (uint32_t)cand_cap, cc2, cand2, d_counts);
    }
    }
    prof_span(ev, st, ST_VERIFY);
    prof_mark(&ev, st);
    ExtParams P = {d->P.k, d->P.A, d->P.B, d->P.clip5, d->P.clip3, d->P.T, d->P.X};
    long long warps_wanted = cand_cap < (long long)d->num_sms * 40 ? cand_cap : (long long)d->num_sms * 40;
    int ext_blocks = (int)((warps_wanted + 7) / 8);
    k_extend<<<ext_blocks, 256, 0, st>>>((const uint32_t *)b->packed, lay.words_per_read, lay.quads_per_pair,
                                         b->uniform_len, b->lens, b->nread_ids, b->nmask, (int)b->n_nreads, cand2,
                                         d_counts, (uint32_t)cand_cap, d->d_table, d->tmask, d->d_anchor, d->G, d->kp,
                                         d->stride, P, slots, cc3, d_counts + AF_CNT_SCRATCH, d->d_apkp[0], d->d_apkp[1], d->d_apn[0], d->d_apn[1], long_reads ? 0 : g_walk);   // the word-parallel mask is built for 256 positions
    prof_span(ev, st, ST_EXTEND);
    prof_mark(&ev, st);
    if (sink) k_hit_scatter<true><<<sg2, CB_THREADS, 0, st>>>(slots, (uint32_t)cand_cap, cc3, (uint4 *)d_hits, (uint32_t)hits_cap, d_counts, *sink);
    else k_hit_scatter<false><<<sg2, CB_THREADS, 0, st>>>(slots, (uint32_t)cand_cap, cc3, (uint4 *)d_hits, (uint32_t)hits_cap, d_counts, af_sink());
    prof_span(ev, st, ST_COMPACT2);
    g_launches += 5;
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}

extern "C" int af_anchor_batch(const af_dev_index_t *d, const af_batch_t *b, void *workspace, size_t workspace_bytes,
                               int64_t cand_cap, af_hit_t *d_hits, int64_t hits_cap, uint32_t *d_counts, void *stream) {
    return anchor_batch_impl(d, b, workspace, workspace_bytes, cand_cap, d_hits, hits_cap, d_counts, nullptr, stream);
}

// The same path with the hit exchange fused into its last kernel: k_hit_scatter<true> stores each record
// into this rank's log on every GPU of the job (NVLink peer stores), see af_exchange.cu.
extern "C" int af_anchor_batch_exchange(const af_dev_index_t *d, const af_batch_t *b, void *workspace, size_t workspace_bytes,
                                        int64_t cand_cap, af_hit_t *d_hits, int64_t hits_cap, uint32_t *d_counts,
                                        af_exchange_t *ex, int32_t slot, int64_t pair_base, void *stream) {
    af_sink sink;
    int rc = af_exchange_sink(ex, slot, pair_base, &sink);
    if (rc) return rc;
    if (d && ex->device != d->device) { af_set_error("af_anchor_batch_exchange: exchange lives on device %d, index on %d", ex->device, d->device); return AF_ERR_ARG; }
    if (b && b->n_pairs == 0) return anchor_batch_impl(d, b, workspace, workspace_bytes, cand_cap, d_hits, hits_cap, d_counts, nullptr, stream);
    rc = anchor_batch_impl(d, b, workspace, workspace_bytes, cand_cap, d_hits, hits_cap, d_counts, &sink, stream);
    if (rc == AF_OK) ex->seq[slot]++;      // the batch parity k_hit_scatter double-buffers the log tail by
    return rc;
}

// ------------------------------------------------------------------------------------------
// index upload
// ------------------------------------------------------------------------------------------
extern "C" int af_index_upload(const af_index_t *idx, int device, af_dev_index_t **out) {
    if (!idx || !out) { af_set_error("af_index_upload: null"); return AF_ERR_ARG; }
    int ndev = 0;
    AF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { af_set_error("af_index_upload: device %d of %d", device, ndev); return AF_ERR_CUDA; }
    AF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { af_set_error("af_index_upload: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return AF_ERR_CUDA; }
    af_dev_index *d = new af_dev_index();
    d->device = device; d->P = idx->P; d->kp = idx->kp; d->stride = idx->stride; d->G = idx->G;
    d->fmul = idx->fmul; d->nb = idx->nb; d->tmask = idx->tmask; d->pad_byte = idx->pad_byte;
    d->num_sms = prop.multiProcessorCount;
    d->saturated = (long long)idx->n_overflow * 200 > (long long)idx->nb;
    d->bloom = idx->bloom != 0;
    d->d_filter = nullptr; d->d_table = nullptr; d->d_anchor = nullptr; d->d_member = nullptr;
    d->d_apk[0] = d->d_apk[1] = nullptr;
    d->anchor_has_n = idx->anchor_has_n;
    d->fmul2 = idx->fmul2; d->nb2 = idx->nb2; d->d_filter2 = nullptr;
    cudaError_t e = cudaMalloc(&d->d_filter, idx->filter.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_filter2, idx->filter2.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d->d_filter2, idx->filter2.data(), idx->filter2.size() * 4, cudaMemcpyHostToDevice);
    d->d_apkp[0] = d->d_apkp[1] = nullptr;
    d->d_apn[0] = d->d_apn[1] = nullptr;
    if (idx->anchor_has_n) {                                // N bitmasks for k_tail's word-parallel match masks
        const size_t nw = ((size_t)idx->G + 31) / 32 + 8 + 10;
        std::vector<uint32_t> bits[2];
        bits[0].assign(nw, 0u); bits[1].assign(nw, 0u);
        for (int i = 0; i < idx->G; i++)
            if (idx->codes[(size_t)i] > 3) {
                const int f = i + 256, r = idx->G - 1 - i + 256;
                bits[0][(size_t)f >> 5] |= 1u << (f & 31);
                bits[1][(size_t)r >> 5] |= 1u << (r & 31);
            }
        for (int o = 0; o < 2; o++) {
            if (e == cudaSuccess) e = cudaMalloc(&d->d_apn[o], nw * 4);
            if (e == cudaSuccess) e = cudaMemcpy(d->d_apn[o], bits[o].data(), nw * 4, cudaMemcpyHostToDevice);
        }
    }
    d->apk_words = (int)idx->apk[0].size() + 16;
    for (int o = 0; o < 2; o++) {
        if (e == cudaSuccess) e = cudaMalloc(&d->d_apk[o], idx->apk[o].size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d->d_apk[o], idx->apk[o].data(), idx->apk[o].size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&d->d_apkp[o], (size_t)d->apk_words * 4);
        if (e == cudaSuccess) e = cudaMemset(d->d_apkp[o], 0, 16 * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d->d_apkp[o] + 16, idx->apk[o].data(), idx->apk[o].size() * 4, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&d->d_member, idx->member.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d->d_member, idx->member.data(), idx->member.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_table, idx->table.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_anchor, idx->codes.size() + 256);
    if (e == cudaSuccess) e = cudaMemcpy(d->d_filter, idx->filter.data(), idx->filter.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d->d_table, idx->table.data(), idx->table.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d->d_anchor, idx->codes.data(), idx->codes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        af_set_error("af_index_upload: %s", cudaGetErrorString(e));
        cudaFree(d->d_filter); cudaFree(d->d_table); cudaFree(d->d_anchor); cudaFree(d->d_member); cudaFree(d->d_apk[0]); cudaFree(d->d_apk[1]); cudaFree(d->d_apkp[0]); cudaFree(d->d_apkp[1]); cudaFree(d->d_apn[0]); cudaFree(d->d_apn[1]); cudaFree(d->d_filter2);
        delete d;
        return AF_ERR_CUDA;
    }
    *out = d;
    return AF_OK;
}

extern "C" int af_dev_index_device(const af_dev_index_t *d) { return d ? d->device : -1; }

extern "C" void af_dev_index_free(af_dev_index_t *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaFree(d->d_filter); cudaFree(d->d_table); cudaFree(d->d_anchor); cudaFree(d->d_member); cudaFree(d->d_apk[0]); cudaFree(d->d_apk[1]); cudaFree(d->d_apkp[0]); cudaFree(d->d_apkp[1]); cudaFree(d->d_apn[0]); cudaFree(d->d_apn[1]); cudaFree(d->d_filter2);
    delete d;
}

// ------------------------------------------------------------------------------------------
// synthetic pairs straight into packed tiles
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_synth_pairs(af_synth_t s, long long first_pair, long long n_pairs, long long n_tiles, int W, int Q, uint32_t padw,
              uint4 *__restrict__ packed) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_tiles * 32) return;
    const long long tile = p >> 5;
    const int lane = (int)(p & 31);
    af_frag f;
    const bool real = p < n_pairs;
    if (real) f = af_make_frag(s, first_pair + p);
    for (int q = 0; q < Q; q++) {
        uint32_t w4[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int wi = 4 * q + c;
            uint32_t word = wi < 2 * W ? padw : 0u;
            if (real && wi < 2 * W) {
                const int m = wi >= W, w0 = (wi - m * W) * 16;
                for (int i = 0; i < 16 && w0 + i < s.read_len; i++) {
                    uint32_t b = af_read_base(s, f, m, w0 + i);
                    word = (word & ~(3u << (2 * i))) | (b << (2 * i));
                }
            }
            w4[c] = word;
        }
        packed[(tile * Q + q) * 32 + lane] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

extern "C" int af_synth_pairs_device(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, int32_t pad_byte,
                                     void *d_packed, void *stream) {
    if (!s || !d_packed || n_pairs < 0 || s->n_ppm != 0) { af_set_error("af_synth_pairs_device: bad argument (n_ppm must be 0)"); return AF_ERR_ARG; }
    af_layout_t lay;
    int rc = af_layout(s->read_len, n_pairs, &lay);
    if (rc) return rc;
    if (lay.n_tiles == 0) return AF_OK;
    uint32_t padw = 0;
    for (int i = 0; i < 16; i++) padw |= (uint32_t)((pad_byte >> (2 * (i & 3))) & 3) << (2 * i);
    long long threads = lay.n_tiles * 32;
    k_synth_pairs<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *s, first_pair, n_pairs, lay.n_tiles, lay.words_per_read, lay.quads_per_pair, padw, (uint4 *)d_packed);
    g_launches++;
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}

// ------------------------------------------------------------------------------------------
// pinned host memory
// ------------------------------------------------------------------------------------------
extern "C" void *af_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { af_set_error("af_host_alloc: cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}
extern "C" void af_host_free(void *p) { if (p) cudaFreeHost(p); }
