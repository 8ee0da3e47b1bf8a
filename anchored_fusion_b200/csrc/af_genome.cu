// af_genome.cu -- genome pass of the contiguity filter.
//
// In the reference the 2-op anchored reads of a gene are re-aligned to the whole genome by
//     bwa mem -M -t T <genome> <w>_del_tmp.fa > <w>_del_tmp.sam                     functions.py:716
// and del_too_many_reads (functions.py:717-763) drops the reads the genome explains without a junction.
// Here the same question is answered under the anchoring spec of the main path (spec v1: exact
// seed of k = 19, ungapped X-drop extension, bwa's clip rule, one primary record per read) with the roles
// of the main path reversed: the READS are few, so their k'-mers (k' = 12, both orientations, ~90 reads at
// a time) form the shared-memory filter, and the GENOME -- 2 bit/base, resident in HBM, no index of any
// kind -- is streamed past it, sampled every 8 bases (k' + 8 - 1 = 19: every exact match of >= 19 bases
// holds a sampled 12-mer).  Four kernels per pass:
//   k_genome_scan    persistent, 1 CTA per SM, filter staged by the TMA engine; a lane holds 64 genome bases
//                    (one 128-bit load), 8 probes; filter hits -> candidate sample positions      (HBM / LSU)
//   k_genome_seed    1 thread per candidate: exact table (key -> oriented read, offset), 32-base window
//                    compare; the FIRST sample of every run of >= 19 matches becomes a seed       (latency)
//   k_genome_extend  1 warp per seed: 256-bit match mask word-parallel, mismatch-walk extension
//                    (af_device.cuh, shared with k_extend / k_tail), atomicMax of (score, strand, diagonal)
//   k_genome_finish  1 warp per read: the winning diagonal once more -> record
// The genome is ONE sequence, contigs joined by 256 N (af_genome_from_*): an N never matches, a read cannot
// span two contigs, and no diagonal that holds a seed leaves the sequence, so the kernels need no bounds.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "af_common.h"
#include "af_device.cuh"

static const int GQ_WORDS = 18;        // packed words per oriented read: 16 + 2 zero words (64-bit windows may start at base 255)
static const int GQ_NWORDS = 9;        // N-mask words per oriented read: 8 + 1
static const int64_t G_BLOCK = 8192;   // genome bases per warp iteration of the scan: 32 lanes x 4 loads x 64 bases
static const int G_KP = 12, G_STRIDE = 8;
static const uint32_t G_KMASK = (1u << (2 * G_KP)) - 1u;
static const int G_KEY_BUDGET = 24000; // (k'-mer, read, offset) entries per pass: ~86 reads of 150 bases, both orientations

struct af_genome {
    int device = 0, num_sms = 0;
    int64_t G = 0;        // bases of the concatenation (what positions refer to)
    int64_t G_scan = 0;   // G rounded up to G_BLOCK: what the scan covers; the arrays hold one more block, all N
    uint32_t *d_pk = nullptr, *d_nm = nullptr;
    std::vector<std::string> names;
    std::vector<int64_t> starts, lens;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // buffers of af_genome_align, grown on demand
    uint32_t *d_cand = nullptr; int64_t cand_cap = 0;
    ulonglong2 *d_seeds = nullptr; int64_t seed_cap = 0;
    uint32_t *d_counters = nullptr;
    uint32_t *d_filter = nullptr; uint2 *d_table = nullptr; int64_t table_cap = 0;
    uint32_t *d_qpk = nullptr, *d_qnm = nullptr; uint16_t *d_qlen = nullptr; unsigned long long *d_best = nullptr;
    af_genome_hit_t *d_recs = nullptr; int64_t q_cap = 0;
    std::mutex call_mu;   // af_genome_align uses the buffers above: one call at a time per genome
};

// ---- host side: concatenation and 2-bit packing live in af_genome_host.cpp (no CUDA: fuzzed under ASan) ---------
static int genome_upload(af_genome_host &h, int device, af_genome_t **out);

extern "C" int af_genome_from_contigs(const char *const *names, const char *const *seqs, const int64_t *lens, int32_t n, int device,
                                      af_genome_t **out) {
    if (!out) { af_set_error("af_genome_from_contigs: null"); return AF_ERR_ARG; }
    af_genome_host h;
    const int rc = af_genome_host_from_contigs(names, seqs, lens, n, h);
    return rc ? rc : genome_upload(h, device, out);
}

extern "C" int af_genome_from_fasta(const char *path, int device, af_genome_t **out) {
    if (!out) { af_set_error("af_genome_from_fasta: null"); return AF_ERR_ARG; }
    af_genome_host h;
    const int rc = af_genome_host_from_fasta(path, h);
    return rc ? rc : genome_upload(h, device, out);
}

static int genome_device(int device, int *num_sms) {
    int ndev = 0;
    AF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { af_set_error("af_genome: device %d of %d", device, ndev); return AF_ERR_CUDA; }
    AF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { af_set_error("af_genome: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return AF_ERR_CUDA; }
    *num_sms = prop.multiProcessorCount;
    return AF_OK;
}

static int genome_alloc(af_genome *g) {
    g->G_scan = (g->G + G_BLOCK - 1) / G_BLOCK * G_BLOCK;
    const size_t pkw = (size_t)((g->G_scan + G_BLOCK) >> 4) + 8, nmw = (size_t)((g->G_scan + G_BLOCK) >> 5) + 8;
    AF_CUDA(cudaMalloc(&g->d_pk, pkw * 4));
    AF_CUDA(cudaMalloc(&g->d_nm, nmw * 4));
    AF_CUDA(cudaMemset(g->d_pk, 0, pkw * 4));
    AF_CUDA(cudaMemset(g->d_nm, 0xFF, nmw * 4));                 // everything not written below is N
    AF_CUDA(cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) AF_CUDA(cudaEventCreate(&g->ev[i]));
    AF_CUDA(cudaMalloc(&g->d_counters, 16 * 4));
    AF_CUDA(cudaMalloc(&g->d_filter, (size_t)AF_MAX_BUCKETS * 4));
    return AF_OK;
}

static int genome_upload(af_genome_host &h, int device, af_genome_t **out) {
    if (h.n >= ((int64_t)1 << 35)) { af_set_error("af_genome: %lld bases; the candidate list indexes at most 2^35", (long long)h.n); return AF_ERR_ARG; }
    af_genome *g = new af_genome();
    int rc = genome_device(device, &g->num_sms);
    if (rc == AF_OK) { g->device = device; g->G = h.n; rc = genome_alloc(g); }
    if (rc == AF_OK) {
        cudaError_t e = cudaMemcpy(g->d_pk, h.pk.data(), (size_t)((h.n + 15) >> 4) * 4, cudaMemcpyHostToDevice);
        // whole N words of the stream; the partial last word keeps the 1 bits past G
        const size_t nw = (size_t)((h.n + 31) >> 5);
        if (h.n & 31) h.nm[nw - 1] |= ~0u << (h.n & 31);
        if (e == cudaSuccess) e = cudaMemcpy(g->d_nm, h.nm.data(), nw * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { af_set_error("af_genome: upload: %s", cudaGetErrorString(e)); rc = AF_ERR_CUDA; }
    }
    if (rc != AF_OK) { af_genome_free(g); return rc; }
    g->names.swap(h.names); g->starts.swap(h.starts); g->lens.swap(h.lens);
    *out = g;
    return AF_OK;
}

// one thread = 32 bases of the synthetic genome: 2 packed words + 1 N word
__global__ void k_genome_synth(unsigned long long seed, long long len, long long G, uint32_t *__restrict__ pk, uint32_t *__restrict__ nm) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, x0 = t * 32;
    if (x0 >= G) return;
    uint32_t w[2] = {0u, 0u}, nb = 0u;
    for (int i = 0; i < 32; i++) {
        const long long x = x0 + i, r = x - AF_GENOME_SEP;
        uint32_t c;
        if (x < G && r >= 0 && r < len) c = af_ref_base(seed, r);
        else { nb |= 1u << i; c = af_mix32((uint32_t)x * 0x9E3779B1u) & 3u; }
        w[i >> 4] |= c << (2 * (i & 15));
    }
    pk[2 * t] = w[0]; pk[2 * t + 1] = w[1]; nm[t] = nb;
}

extern "C" int af_genome_synth(uint64_t seed, int64_t len, int device, af_genome_t **out) {
    if (len <= 0 || !out || len >= ((int64_t)1 << 35) - 4096) { af_set_error("af_genome_synth: bad length"); return AF_ERR_ARG; }
    af_genome *g = new af_genome();
    int rc = genome_device(device, &g->num_sms);
    if (rc == AF_OK) { g->device = device; g->G = len + 2 * AF_GENOME_SEP; rc = genome_alloc(g); }
    if (rc == AF_OK) {
        const long long threads = (g->G + 31) / 32;
        k_genome_synth<<<(unsigned)((threads + 255) / 256), 256, 0, g->st>>>(seed, len, g->G, g->d_pk, g->d_nm);
        af_note_launches(1);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->st);
        if (e != cudaSuccess) { af_set_error("af_genome_synth: %s", cudaGetErrorString(e)); rc = AF_ERR_CUDA; }
    }
    if (rc != AF_OK) { af_genome_free(g); return rc; }
    g->names.push_back("synth"); g->starts.push_back(AF_GENOME_SEP); g->lens.push_back(len);
    *out = g;
    return AF_OK;
}

extern "C" void af_genome_free(af_genome_t *g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaFree(g->d_pk); cudaFree(g->d_nm); cudaFree(g->d_cand); cudaFree(g->d_seeds); cudaFree(g->d_counters); cudaFree(g->d_filter);
    cudaFree(g->d_table); cudaFree(g->d_qpk); cudaFree(g->d_qnm); cudaFree(g->d_qlen); cudaFree(g->d_best); cudaFree(g->d_recs);
    for (int i = 0; i < 4; i++) if (g->ev[i]) cudaEventDestroy(g->ev[i]);
    if (g->st) cudaStreamDestroy(g->st);
    delete g;
}
extern "C" int64_t af_genome_length(const af_genome_t *g) { return g ? g->G : 0; }
extern "C" int32_t af_genome_n_contigs(const af_genome_t *g) { return g ? (int32_t)g->names.size() : 0; }
extern "C" int af_genome_contig(const af_genome_t *g, int32_t i, const char **name, int64_t *start, int64_t *len) {
    if (!g || i < 0 || i >= (int32_t)g->names.size()) { af_set_error("af_genome_contig: index"); return AF_ERR_ARG; }
    if (name) *name = g->names[(size_t)i].c_str();
    if (start) *start = g->starts[(size_t)i];
    if (len) *len = g->lens[(size_t)i];
    return AF_OK;
}

// ---- device side ----------------------------------------------------------------------------------
// 32 bases (64 bits) of the packed genome from base pos >= 0
__device__ __forceinline__ unsigned long long gen_window(const uint32_t *__restrict__ a, long long pos) {
    const long long wi = pos >> 4;
    const int sh = 2 * (int)(pos & 15);
    const uint32_t w0 = __ldg(a + wi), w1 = __ldg(a + wi + 1), w2 = __ldg(a + wi + 2);
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}
// 32 bits of a 1-bit-per-base mask from base pos >= 0
__device__ __forceinline__ uint32_t gen_bits(const uint32_t *__restrict__ m, long long pos) {
    const long long wi = pos >> 5;
    return __funnelshift_r(__ldg(m + wi), __ldg(m + wi + 1), (int)(pos & 31));
}
// bit i = bases i of the two 32-base windows are equal
__device__ __forceinline__ uint32_t eq32(unsigned long long a, unsigned long long b) {
    const unsigned long long x = a ^ b, ne = x | (x >> 1);
    return ~(even_bits((uint32_t)ne) | (even_bits((uint32_t)(ne >> 32)) << 16));
}

// Persistent, one CTA per SM.  A warp iteration covers G_BLOCK = 8192 genome bases: four 128-bit loads per lane (each
// 512 contiguous bytes per warp), the next iteration's four already in flight (register double buffer), 32 probes
// per lane.  Filter hits are staged per warp in shared memory and leave with one global atomic per ~64 candidates
// (one atomic per hit-holding warp iteration serialised 300 k atomics per pass on one address: 0.78 ms per pass
// instead of 0.3, measured).
static const int GS_WBUF = 160;                          // staged candidates per warp: flushed at >= 64, one iteration adds <= 64 (else direct)
__global__ void __launch_bounds__(1024, 1)
k_genome_scan(const uint4 *__restrict__ pk4, long long n_iters, const uint32_t *__restrict__ g_filter, uint32_t fmul, uint32_t nb,
              uint32_t *__restrict__ cand, uint32_t cand_cap, uint32_t *__restrict__ counters) {
    extern __shared__ __align__(128) uint32_t filt[];
    uint32_t *wbuf = filt + nb + (threadIdx.x >> 5) * GS_WBUF;
    stage_filter(filt, g_filter, nb);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    long long it = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t *pkw = reinterpret_cast<const uint32_t *>(pk4);
    uint4 v[4];
    uint32_t tailw = 0;
    int staged = 0;                                      // warp-uniform
#pragma unroll
    for (int b = 0; b < 4; b++) v[b] = make_uint4(0, 0, 0, 0);
    if (it < n_iters) {
#pragma unroll
        for (int b = 0; b < 4; b++) v[b] = ld_stream_v4(pk4 + (it * 4 + b) * 32 + lane);
        if (lane == 31) tailw = __ldg(pkw + (it * 4 + 4) * 128);
    }
    auto flush = [&]() {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], (uint32_t)staged);
        base = __shfl_sync(FULL, base, 0);
        __syncwarp();
        for (int i = lane; i < staged; i += 32) if (base + i < cand_cap) cand[base + i] = wbuf[i];
        __syncwarp();
        staged = 0;
    };
    for (; it < n_iters; it += nw) {
        uint4 cur[4];
#pragma unroll
        for (int b = 0; b < 4; b++) cur[b] = v[b];
        const uint32_t tw = tailw;
        const long long itn = it + nw;
        if (itn < n_iters) {                              // next blocks in flight while this one is probed
#pragma unroll
            for (int b = 0; b < 4; b++) v[b] = ld_stream_v4(pk4 + (itn * 4 + b) * 32 + lane);
            if (lane == 31) tailw = __ldg(pkw + (itn * 4 + 4) * 128);
        }
        uint32_t hits = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t nx = __shfl_down_sync(FULL, cur[b].x, 1);
            const uint32_t first_next = b < 3 ? __shfl_sync(FULL, cur[b < 3 ? b + 1 : b].x, 0) : tw;
            if (lane == 31) nx = first_next;
            const uint32_t w[5] = {cur[b].x, cur[b].y, cur[b].z, cur[b].w, nx};
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t key = __funnelshift_r(w[t >> 1], w[(t >> 1) + 1], (t & 1) * 16) & G_KMASK;
                hits |= ((af_filter_probe(key, filt, fmul, nb) & AF_F_HIGH) != 0u ? 1u : 0u) << (8 * b + t);
            }
        }
        if (__ballot_sync(FULL, hits != 0u)) {
            const int cnt = __popc(hits);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
            const int total = __shfl_sync(FULL, inc, 31);
            const bool direct = total > 64;               // a low-complexity stretch: straight to the list
            uint32_t base = 0;
            if (direct) { if (lane == 0) base = atomicAdd(&counters[0], (uint32_t)total); base = __shfl_sync(FULL, base, 0); }
            uint32_t pos = (direct ? base : (uint32_t)staged) + (uint32_t)(inc - cnt);
            while (hits) {
                const int bt = __ffs(hits) - 1;
                hits &= hits - 1;
                const uint32_t sample = (uint32_t)((((it * 4 + (bt >> 3)) * 32 + lane) << 3) + (bt & 7));   // genome position / 8
                if (direct) { if (pos < cand_cap) cand[pos] = sample; } else wbuf[pos] = sample;
                pos++;
            }
            if (!direct) { staged += total; if (staged >= 64) flush(); }
        }
    }
    if (staged) flush();
}

// One thread per candidate sample.  A table entry (read, offset i) with the sample's 12-mer puts the read on diagonal
// d = p - i; the 32 bases around the sample tell whether the sample lies in a run of >= 19 matches and is the FIRST
// sample of that run (fewer than 8 matching bases to its left) -- every such run is reported exactly once, by the one
// sample the sampling guarantees.
__global__ void __launch_bounds__(256)
k_genome_seed(const uint32_t *__restrict__ pk, const uint32_t *__restrict__ nm, const uint32_t *__restrict__ cand, uint32_t cand_cap,
              const uint2 *__restrict__ table, uint32_t tmask, const uint32_t *__restrict__ qpk, const uint32_t *__restrict__ qnm,
              const uint16_t *__restrict__ qlen, ulonglong2 *__restrict__ seeds, uint32_t seed_cap, uint32_t *__restrict__ counters) {
    const uint32_t n = min(counters[0], cand_cap);
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const long long p = (long long)cand[c] * G_STRIDE;
        const uint32_t key = (uint32_t)gen_window(pk, p) & G_KMASK;
        uint32_t s = af_table_hash(key, tmask);
        for (;;) {
            const uint2 e = __ldg(table + s);
            if (e.x == AF_T_EMPTY) break;
            s = (s + 1) & tmask;
            if (e.x != key) continue;
            const uint32_t qid = e.y >> 8;
            const int i = (int)(e.y & 255u), L = qlen[qid];
            const int a = max(i - G_STRIDE, 0), o = i - a;       // window = read bases [a, a + 32), the sample starts at bit o
            const long long gs = p - o;
            if (gs < 0) continue;
            uint32_t eq = eq32(packed_window(qpk + (size_t)qid * GQ_WORDS, a), gen_window(pk, gs));
            const uint32_t *qn = qnm + (size_t)qid * GQ_NWORDS;
            eq &= ~__funnelshift_r(qn[a >> 5], qn[(a >> 5) + 1], a & 31) & ~gen_bits(nm, gs);
            if (L - a < 32) eq &= (1u << (L - a)) - 1u;
            const uint32_t up = ~(eq >> o);                       // zeros shifted in from the top end every run
            const int r = up ? __ffs(up) - 1 : 32;                // matches from the sample's first base on
            const int l = o ? min(__clz(~(eq << (32 - o))), o) : 0;   // matches immediately left of it
            if (l < G_STRIDE && r >= G_KP && l + r >= 19) {
                const uint32_t at = atomicAdd(&counters[1], 1u);
                if (at < seed_cap) seeds[at] = make_ulonglong2((unsigned long long)qid, (unsigned long long)(p - i));
            }
        }
    }
}

// the 256-bit match mask of oriented read qid on genome diagonal d (read base i faces genome base i + d); lane c < 8
// returns word c, the other lanes 0
__device__ __forceinline__ uint32_t genome_mask(const uint32_t *__restrict__ pk, const uint32_t *__restrict__ nm, const uint32_t *__restrict__ qpk,
                                                const uint32_t *__restrict__ qnm, uint32_t qid, int L, long long d, int lane) {
    const int c = lane & 7, left = L - 32 * c;
    uint32_t m = 0;
    if (lane < 8 && left > 0) {
        const uint32_t *q = qpk + (size_t)qid * GQ_WORDS;
        const unsigned long long rw = (unsigned long long)q[2 * c] | ((unsigned long long)q[2 * c + 1] << 32);
        m = eq32(rw, gen_window(pk, d + 32 * c)) & ~qnm[(size_t)qid * GQ_NWORDS + c] & ~gen_bits(nm, d + 32 * c);
        if (left < 32) m &= (1u << left) - 1u;
    }
    return m;
}

// best-record key: score, then strand 0 before 1, then the smaller diagonal (oracle/af_oracle.c::anchor_read)
static const unsigned long long G_DMAX = (1ull << 41) - 1ull;
__device__ __forceinline__ unsigned long long best_key(int sc, int strand, long long d) {
    return ((unsigned long long)sc << 42) | ((unsigned long long)(1 - strand) << 41) | (G_DMAX - (unsigned long long)d);
}

__global__ void __launch_bounds__(256)
k_genome_extend(const uint32_t *__restrict__ pk, const uint32_t *__restrict__ nm, const ulonglong2 *__restrict__ seeds, uint32_t seed_cap,
                const uint32_t *__restrict__ qpk, const uint32_t *__restrict__ qnm, const uint16_t *__restrict__ qlen, ExtParams P,
                unsigned long long *__restrict__ best, const uint32_t *__restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(counters[1], seed_cap);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nwarps) {
        const ulonglong2 sd = seeds[w];
        const uint32_t qid = (uint32_t)sd.x;
        const long long d = (long long)sd.y;
        const int L = qlen[qid];
        const uint32_t mw = genome_mask(pk, nm, qpk, qnm, qid, L, d, lane);
        int qb, qe;
        // no diagonal with a seed leaves the sequence (256 N at both ends), so the anchor-end clamps of the
        // evaluation are switched off by a far-away diagonal / length
        const int sc = eval_mask_walk(mw, 1 << 20, L, 0x7FFFFFFF, P, lane, qb, qe);
        if (lane == 0 && sc >= P.T) atomicMax(&best[qid >> 1], best_key(sc, (int)(qid & 1u), d));
    }
}

__global__ void __launch_bounds__(256)
k_genome_finish(const uint32_t *__restrict__ pk, const uint32_t *__restrict__ nm, const uint32_t *__restrict__ qpk,
                const uint32_t *__restrict__ qnm, const uint16_t *__restrict__ qlen, ExtParams P,
                const unsigned long long *__restrict__ best, long long n_reads, af_genome_hit_t *__restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_reads) return;
    const unsigned long long key = best[j];
    af_genome_hit_t h;
    h.pos = 0; h.read_id = (uint32_t)j; h.clip_l = h.m_len = h.clip_r = h.score_strand = 0; h.reserved = 0;
    if (key) {
        const int s = 1 - (int)((key >> 41) & 1ull);
        const long long d = (long long)(G_DMAX - (key & G_DMAX));
        const uint32_t qid = (uint32_t)(2 * j + s);
        const int L = qlen[qid];
        int qb = 0, qe = 0;
        const int sc = eval_mask_walk(genome_mask(pk, nm, qpk, qnm, qid, L, d, lane), 1 << 20, L, 0x7FFFFFFF, P, lane, qb, qe);
        h.pos = d + qb + 1;
        h.clip_l = (uint16_t)qb; h.m_len = (uint16_t)(qe - qb); h.clip_r = (uint16_t)(L - qe);
        h.score_strand = (uint16_t)(sc * 2 + s);
    }
    if (lane == 0) recs[j] = h;
}

// ---- af_genome_align ------------------------------------------------------------------------------
template <class T>
static int grow(T **p, int64_t *cap, int64_t want) {
    if (*cap >= want) return AF_OK;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    AF_CUDA(cudaMalloc(p, (size_t)want * sizeof(T)));
    *cap = want;
    return AF_OK;
}

struct PassIndex {
    int64_t j0 = 0, j1 = 0;           // reads [j0, j1)
    std::vector<uint32_t> filter, table;
    uint32_t nb = 0, fmul = 0, tmask = 0;
    int64_t n_entries = 0;
    double ms = 0;
    bool ready = false;
};

// filter + exact table over every 12-mer (no N) of the oriented reads 2*j0 .. 2*j1-1; value = oriented read << 8 | offset
static void build_pass(const std::vector<uint32_t> &qpk, const std::vector<uint32_t> &qnm, const std::vector<uint16_t> &qlen, PassIndex &pi) {
    const auto t0 = std::chrono::steady_clock::now();
    struct Ent { uint32_t key, val; };
    std::vector<Ent> ents;
    for (int64_t q = 2 * pi.j0; q < 2 * pi.j1; q++) {
        const int L = qlen[(size_t)q];
        int run = 0;
        uint32_t km = 0;
        for (int i = 0; i < L; i++) {
            const bool isn = (qnm[(size_t)q * GQ_NWORDS + (size_t)(i >> 5)] >> (i & 31)) & 1u;
            const uint32_t c = (qpk[(size_t)q * GQ_WORDS + (size_t)(i >> 4)] >> (2 * (i & 15))) & 3u;
            if (isn) { run = 0; km = 0; continue; }
            km = (km >> 2) | (c << (2 * (G_KP - 1)));
            if (++run >= G_KP) ents.push_back({km, ((uint32_t)q << 8) | (uint32_t)(i - G_KP + 1)});
        }
    }
    pi.n_entries = (int64_t)ents.size();
    uint32_t slots = 1024;
    while (slots < 4 * ents.size()) slots <<= 1;
    pi.tmask = slots - 1;
    pi.table.assign((size_t)slots * 2, AF_T_EMPTY);
    std::vector<uint32_t> keys;
    keys.reserve(ents.size());
    for (const Ent &e : ents) {
        uint32_t s = af_table_hash(e.key, pi.tmask);
        while (pi.table[(size_t)s * 2] != AF_T_EMPTY) s = (s + 1) & pi.tmask;
        pi.table[(size_t)s * 2] = e.key;
        pi.table[(size_t)s * 2 + 1] = e.val;
        keys.push_back(e.key);
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    uint32_t nb = (uint32_t)std::min<uint64_t>(AF_MAX_BUCKETS, std::max<uint64_t>(AF_MIN_BUCKETS, (uint64_t)keys.size() * 4));
    pi.nb = (nb + 31u) & ~31u;
    int32_t ov = 0;
    af_filter_pick(keys, G_KMASK, pi.nb, 4, 14, pi.fmul, pi.filter, &ov);
    pi.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// one pass on the device: upload, scan -> seeds -> extension; repeated with larger buffers if a list overflowed
static int run_pass(af_genome *g, const PassIndex &pi, const ExtParams &EP, long long n_iters, size_t max_dyn, af_genome_stats_t &S) {
    cudaStream_t st = g->st;
    const size_t smem = (size_t)pi.nb * 4 + 32 * GS_WBUF * 4;
    if (smem > max_dyn) { af_set_error("af_genome_align: filter of %u buckets does not fit shared memory", pi.nb); return AF_ERR_ARG; }
    const int64_t slots = (int64_t)pi.tmask + 1;
    if (grow(&g->d_table, &g->table_cap, slots)) return AF_ERR_CUDA;
    AF_CUDA(cudaMemcpyAsync(g->d_filter, pi.filter.data(), (size_t)pi.nb * 4, cudaMemcpyHostToDevice, st));
    AF_CUDA(cudaMemcpyAsync(g->d_table, pi.table.data(), (size_t)slots * 8, cudaMemcpyHostToDevice, st));
    for (;;) {
        AF_CUDA(cudaMemsetAsync(g->d_counters, 0, 16 * 4, st));
        if (pi.n_entries) {
            AF_CUDA(cudaEventRecord(g->ev[0], st));
            k_genome_scan<<<g->num_sms, 1024, smem, st>>>((const uint4 *)g->d_pk, n_iters, g->d_filter, pi.fmul, pi.nb, g->d_cand,
                                                           (uint32_t)g->cand_cap, g->d_counters);
            AF_CUDA(cudaEventRecord(g->ev[1], st));
            k_genome_seed<<<g->num_sms * 8, 256, 0, st>>>(g->d_pk, g->d_nm, g->d_cand, (uint32_t)g->cand_cap, g->d_table, pi.tmask, g->d_qpk,
                                                           g->d_qnm, g->d_qlen, g->d_seeds, (uint32_t)g->seed_cap, g->d_counters);
            k_genome_extend<<<g->num_sms * 8, 256, 0, st>>>(g->d_pk, g->d_nm, g->d_seeds, (uint32_t)g->seed_cap, g->d_qpk, g->d_qnm, g->d_qlen,
                                                             EP, g->d_best, g->d_counters);
            af_note_launches(3);
            AF_CUDA(cudaGetLastError());
        }
        uint32_t cnt[2] = {0, 0};
        AF_CUDA(cudaMemcpyAsync(cnt, g->d_counters, 8, cudaMemcpyDeviceToHost, st));
        AF_CUDA(cudaStreamSynchronize(st));
        if (pi.n_entries) { float ms = 0; AF_CUDA(cudaEventElapsedTime(&ms, g->ev[0], g->ev[1])); S.scan_ms += ms; }
        if ((int64_t)cnt[0] > g->cand_cap || (int64_t)cnt[1] > g->seed_cap) {      // nothing is dropped: the pass runs again with room
            if (grow(&g->d_cand, &g->cand_cap, std::max<int64_t>(g->cand_cap, (int64_t)cnt[0] + (cnt[0] >> 2))) ||
                grow(&g->d_seeds, &g->seed_cap, std::max<int64_t>(g->seed_cap, (int64_t)cnt[1] + (cnt[1] >> 2)))) return AF_ERR_CUDA;
            S.n_retries++;
            continue;
        }
        S.n_candidates += cnt[0]; S.n_seeds += cnt[1];
        break;
    }
    S.n_passes++;
    return AF_OK;
}

extern "C" int af_genome_align(af_genome_t *g, const char *reads, const int64_t *offs, int64_t n_reads, const af_params_t *params,
                               int32_t reads_per_pass, af_genome_hit_t *hits_out, int64_t *n_hits_out, af_genome_stats_t *stats) {
    if (!g || !offs || n_reads < 0 || !n_hits_out || (n_reads && (!reads || !hits_out)) || n_reads >= (1 << 23)) { af_set_error("af_genome_align: bad argument"); return AF_ERR_ARG; }
    af_params_t Pp;
    if (params) Pp = *params; else af_default_params(&Pp);
    if (Pp.k != 19 || Pp.A <= 0 || Pp.B < 0 || Pp.X < 0 || (int64_t)2 * Pp.A * AF_GENOME_MAX_READ_LEN + 1 > 65535 || Pp.clip5 < 0 || Pp.clip3 < 0) {
        af_set_error("af_genome_align: unsupported parameters (k must be 19, 2*A*%d+1 must fit 16 bits)", AF_GENOME_MAX_READ_LEN);
        return AF_ERR_ARG;
    }
    af_genome_stats_t S;
    memset(&S, 0, sizeof(S));
    S.genome_bases = g->G;
    *n_hits_out = 0;
    if (stats) *stats = S;
    if (n_reads == 0) return AF_OK;
    std::lock_guard<std::mutex> call_lock(g->call_mu);
    AF_CUDA(cudaSetDevice(g->device));

    // oriented reads: 2j = the read, 2j+1 = its reverse complement; N packed as A with its mask bit set
    const int64_t nq = 2 * n_reads;
    std::vector<uint32_t> qpk((size_t)nq * GQ_WORDS, 0u), qnm((size_t)nq * GQ_NWORDS, 0u);
    std::vector<uint16_t> qlen((size_t)nq, 0);
    std::vector<uint8_t> codes(AF_GENOME_MAX_READ_LEN);
    for (int64_t j = 0; j < n_reads; j++) {
        const int64_t L = offs[j + 1] - offs[j];
        if (L < 0 || L > AF_GENOME_MAX_READ_LEN) { af_set_error("af_genome_align: read %lld has %lld bases, the limit is %d", (long long)j, (long long)L, AF_GENOME_MAX_READ_LEN); return AF_ERR_ARG; }
        for (int64_t i = 0; i < L; i++) codes[(size_t)i] = af_code_of(reads[offs[j] + i]);
        for (int s = 0; s < 2; s++) {
            const size_t q = (size_t)(2 * j + s);
            qlen[q] = (uint16_t)L;
            for (int64_t i = 0; i < L; i++) {
                uint8_t c = s ? codes[(size_t)(L - 1 - i)] : codes[(size_t)i];
                if (c > 3) { qnm[q * GQ_NWORDS + (size_t)(i >> 5)] |= 1u << (i & 31); continue; }
                if (s) c = (uint8_t)(3 - c);
                qpk[q * GQ_WORDS + (size_t)(i >> 4)] |= (uint32_t)c << (2 * (i & 15));
            }
        }
    }
    if (g->q_cap < n_reads) {
        cudaFree(g->d_qpk); cudaFree(g->d_qnm); cudaFree(g->d_qlen); cudaFree(g->d_best); cudaFree(g->d_recs);
        g->d_qpk = g->d_qnm = nullptr; g->d_qlen = nullptr; g->d_best = nullptr; g->d_recs = nullptr; g->q_cap = 0;
        AF_CUDA(cudaMalloc(&g->d_qpk, qpk.size() * 4));
        AF_CUDA(cudaMalloc(&g->d_qnm, qnm.size() * 4));
        AF_CUDA(cudaMalloc(&g->d_qlen, qlen.size() * 2));
        AF_CUDA(cudaMalloc(&g->d_best, (size_t)n_reads * 8));
        AF_CUDA(cudaMalloc(&g->d_recs, (size_t)n_reads * sizeof(af_genome_hit_t)));
        g->q_cap = n_reads;
    }
    cudaStream_t st = g->st;
    AF_CUDA(cudaEventRecord(g->ev[2], st));
    AF_CUDA(cudaMemcpyAsync(g->d_qpk, qpk.data(), qpk.size() * 4, cudaMemcpyHostToDevice, st));
    AF_CUDA(cudaMemcpyAsync(g->d_qnm, qnm.data(), qnm.size() * 4, cudaMemcpyHostToDevice, st));
    AF_CUDA(cudaMemcpyAsync(g->d_qlen, qlen.data(), qlen.size() * 2, cudaMemcpyHostToDevice, st));
    AF_CUDA(cudaMemsetAsync(g->d_best, 0, (size_t)n_reads * 8, st));
    {   // first capacities (AF_GENOME_TEST_CAP: tests make them tiny to exercise the grow-and-repeat path)
        const char *tc = getenv("AF_GENOME_TEST_CAP");
        const int64_t c0 = tc && atoll(tc) > 0 ? atoll(tc) : (int64_t)1 << 22, s0 = tc && atoll(tc) > 0 ? atoll(tc) : (int64_t)1 << 20;
        if (grow(&g->d_cand, &g->cand_cap, c0) || grow(&g->d_seeds, &g->seed_cap, s0)) return AF_ERR_CUDA;
    }

    size_t max_dyn = 0;
    if (allow_full_smem(k_genome_scan, &max_dyn)) return AF_ERR_CUDA;
    ExtParams EP{Pp.k, Pp.A, Pp.B, Pp.clip5, Pp.clip3, Pp.T, Pp.X};
    const long long n_iters = g->G_scan / G_BLOCK;

    // passes: consecutive reads up to the key budget (or the caller's count)
    std::vector<PassIndex> passes;
    for (int64_t j0 = 0; j0 < n_reads;) {
        int64_t j1 = j0, n_ent = 0;
        while (j1 < n_reads) {
            const int L = qlen[(size_t)(2 * j1)];
            const int add = L >= G_KP ? 2 * (L - G_KP + 1) : 0;
            if (j1 > j0 && ((reads_per_pass > 0 && j1 - j0 >= reads_per_pass) || (reads_per_pass <= 0 && n_ent + add > G_KEY_BUDGET))) break;
            n_ent += add;
            j1++;
        }
        passes.emplace_back();
        passes.back().j0 = j0; passes.back().j1 = j1;
        j0 = j1;
    }
    // their filters and tables are built by a few host threads running ahead of the GPU (one pass costs ~2 ms of
    // host time against ~0.3 ms of device time)
    const int n_pass = (int)passes.size();
    std::mutex mu;
    std::condition_variable cv;
    int next_build = 0, consumed = 0;
    bool stop = false;
    const int n_workers = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, n_pass}));
    std::vector<std::thread> workers;
    for (int t = 0; t < n_workers; t++)
        workers.emplace_back([&]() {
            for (;;) {
                int k;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    k = next_build++;
                    if (k >= n_pass) return;
                    cv.wait(lk, [&] { return stop || k < consumed + 16; });   // bounded run-ahead: 16 passes of ~1.3 MB
                    if (stop) return;
                }
                build_pass(qpk, qnm, qlen, passes[(size_t)k]);
                { std::lock_guard<std::mutex> lk(mu); passes[(size_t)k].ready = true; }
                cv.notify_all();
            }
        });
    auto finish_workers = [&]() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (std::thread &w : workers) w.join();
    };
    int rc = AF_OK;
    for (int k = 0; k < n_pass && rc == AF_OK; k++) {
        PassIndex &pi = passes[(size_t)k];
        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return pi.ready; }); }
        S.host_index_ms += pi.ms;
        rc = run_pass(g, pi, EP, n_iters, max_dyn, S);
        std::vector<uint32_t>().swap(pi.filter);
        std::vector<uint32_t>().swap(pi.table);
        { std::lock_guard<std::mutex> lk(mu); consumed = k + 1; }
        cv.notify_all();
    }
    finish_workers();
    if (rc != AF_OK) return rc;
    k_genome_finish<<<(unsigned)((n_reads + 7) / 8), 256, 0, st>>>(g->d_pk, g->d_nm, g->d_qpk, g->d_qnm, g->d_qlen, EP, g->d_best, n_reads, g->d_recs);
    af_note_launches(1);
    AF_CUDA(cudaGetLastError());
    std::vector<af_genome_hit_t> recs((size_t)n_reads);
    AF_CUDA(cudaMemcpyAsync(recs.data(), g->d_recs, (size_t)n_reads * sizeof(af_genome_hit_t), cudaMemcpyDeviceToHost, st));
    AF_CUDA(cudaEventRecord(g->ev[3], st));
    AF_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    AF_CUDA(cudaEventElapsedTime(&ms, g->ev[2], g->ev[3]));
    S.total_ms = ms;
    int64_t n = 0;
    for (int64_t j = 0; j < n_reads; j++) if (recs[(size_t)j].m_len) hits_out[n++] = recs[(size_t)j];
    *n_hits_out = n;
    if (stats) *stats = S;
    return AF_OK;
}
