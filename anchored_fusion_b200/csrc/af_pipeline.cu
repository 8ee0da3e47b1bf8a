// af_pipeline.cu -- host->device streaming executor for the anchoring path.
//
// Takes a HOST batch (pinned buffers filled by af_pack_pairs / af_fastq_next), cuts it into
// tile-aligned chunks, and for each chunk issues cudaMemcpyAsync H2D -> af_anchor_batch ->
// D2H of the counts on the chunk slot's own stream.  N slots are in flight, so the PCIe copy
// of chunk i+1 overlaps the kernels of chunk i.  This is the end-to-end entry the e2e
// measurement and the CLI use; it is the stand-in for the `bwa mem | samtools view | samtools
// sort` process pipe (Anchored_Fusion.py:182) with the sort reduced to the hit list.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "af_common.h"

#define AF_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            af_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return AF_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)


// ---- wire format -> tiles, on the device (include/anchored_fusion.h, "Wire format") -----------------------------
// One thread per pair: word j of the pair's bit stream sits at wire[(tile * NW + j) * 32 + lane] (coalesced), every
// tile word is two neighbouring stream words funnel-shifted; bases past max_read_len in a mate's last word get the
// pad pattern back, the words that pad a pair to whole quads are zero.  Writes one 128-bit quad per store, 512
// contiguous bytes per warp.  Reads 76 and writes 80 bytes per 2 x 150 bp pair: ~0.25 ms per 10 M pairs, behind a
// 14 ms copy.
__global__ void __launch_bounds__(256)
k_wire_expand(const uint32_t *__restrict__ wire, long long n_tiles, int L, int W, int Q, int NW, uint32_t padw, uint4 *__restrict__ packed) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x, tile = p >> 5;
    if (tile >= n_tiles) return;
    const int lane = (int)(p & 31);
    const uint32_t *src = wire + tile * NW * 32 + lane;
    for (int q = 0; q < Q; q++) {
        uint32_t o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int wi = 4 * q + c;
            uint32_t word = 0u;
            if (wi < 2 * W) {
                const int m = wi >= W, t = wi - m * W, nb = min(32, 2 * L - 32 * t), bit = m * 2 * L + 32 * t;
                word = padw;                                // nb <= 0: a word past the read (W is rounded up for long reads)
                if (nb > 0) {
                    const int a = bit >> 5, sh = bit & 31;
                    const uint32_t lo = __ldg(src + a * 32), hi = a + 1 < NW ? __ldg(src + (a + 1) * 32) : 0u;
                    const uint32_t v = __funnelshift_r(lo, hi, sh), mask = nb >= 32 ? 0xFFFFFFFFu : (1u << nb) - 1u;
                    word = (v & mask) | (padw & ~mask);
                }
            }
            o[c] = word;
        }
        packed[(tile * Q + q) * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

static uint32_t pad_word_of(int32_t pad_byte) {
    uint32_t padw = 0;
    for (int i = 0; i < 16; i++) padw |= (uint32_t)((pad_byte >> (2 * (i & 3))) & 3) << (2 * i);
    return padw;
}

static long long g_wire_launches = 0;

static int wire_expand_launch(const void *d_wire, int32_t L, int64_t n_pairs, int32_t pad_byte, void *d_packed, cudaStream_t st) {
    af_layout_t lay;
    int rc = af_layout(L, n_pairs, &lay);
    if (rc) return rc;
    if (lay.n_tiles == 0) return AF_OK;
    const long long threads = lay.n_tiles * 32;
    k_wire_expand<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const uint32_t *)d_wire, lay.n_tiles, L, lay.words_per_read,
                                                                     lay.quads_per_pair, (4 * L + 31) / 32, pad_word_of(pad_byte), (uint4 *)d_packed);
    g_wire_launches++;
    AF_CUDA(cudaGetLastError());
    return AF_OK;
}

extern "C" int af_wire_expand_device(const void *d_wire, int32_t max_read_len, int64_t n_pairs, int32_t pad_byte, void *d_packed, void *stream) {
    if ((!d_wire || !d_packed) && n_pairs > 0) { af_set_error("af_wire_expand_device: null"); return AF_ERR_ARG; }
    return wire_expand_launch(d_wire, max_read_len, n_pairs, pad_byte, d_packed, (cudaStream_t)stream);
}

static const int MAX_GENES = 64;     // anchor indexes one pipeline can scan a resident chunk for

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    void *d_packed = nullptr, *d_ws = nullptr;
    void *d_wire = nullptr;         // wire-format staging of the chunk (af_pipeline_run_wire), allocated on first use
    uint16_t *d_lens = nullptr;
    uint32_t *d_nids = nullptr, *d_nmask = nullptr;
    std::vector<uint32_t *> d_counts;   // per anchor index
    std::vector<af_hit_t *> d_hits;     // per anchor index
    uint32_t *h_counts = nullptr;   // pinned, MAX_GENES x AF_N_COUNTS
    uint32_t *h_nids = nullptr;     // pinned staging of rebased N-read ids
    af_hit_t *h_hits = nullptr;     // pinned
    bool busy = false;
    int n_genes = 0;                // anchor indexes the chunk in flight was scanned for
    int64_t first_pair = 0, n_pairs = 0;
};

struct af_pipeline {
    const af_dev_index_t *d = nullptr;
    int device = 0;
    int64_t slot_pairs = 0, cand_cap = 0, hits_cap = 0, ncap = 0;
    int32_t max_read_len = 0;
    size_t ws_bytes = 0;
    std::vector<Slot> slots;
    long long launches0 = 0, wire_launches0 = 0;
    long long h2d_bytes = 0;        // bytes copied host -> device so far (tiles, lengths, N lists)
};


extern "C" void af_pipeline_free(af_pipeline_t *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        cudaFree(s.d_packed); cudaFree(s.d_wire); cudaFree(s.d_ws); cudaFree(s.d_lens); cudaFree(s.d_nids); cudaFree(s.d_nmask);
        for (uint32_t *c : s.d_counts) cudaFree(c);
        for (af_hit_t *h : s.d_hits) cudaFree(h);
        cudaFreeHost(s.h_counts); cudaFreeHost(s.h_nids); cudaFreeHost(s.h_hits);
        if (s.done) cudaEventDestroy(s.done);
        if (s.st) cudaStreamDestroy(s.st);
    }
    delete p;
}

extern "C" int af_pipeline_create(const af_dev_index_t *d, int64_t slot_pairs, int32_t max_read_len, int32_t n_slots,
                                  af_pipeline_t **out) {
    if (!d || !out || slot_pairs <= 0 || n_slots <= 0 || n_slots > 16) { af_set_error("af_pipeline_create: bad argument"); return AF_ERR_ARG; }
    slot_pairs = (slot_pairs + 31) & ~31ll;  // chunks are whole tiles
    af_layout_t lay;
    int rc = af_layout(max_read_len, slot_pairs, &lay);
    if (rc) return rc;
    af_pipeline *p = new af_pipeline();
    p->d = d;
    p->device = af_dev_index_device(d);
    p->slot_pairs = slot_pairs;
    p->max_read_len = max_read_len;
    p->cand_cap = 2 * slot_pairs;      // every read may be flagged: nothing is ever dropped
    p->hits_cap = 2 * slot_pairs;
    p->ncap = 2 * slot_pairs;
    p->ws_bytes = af_workspace_bytes_len(slot_pairs, p->cand_cap, max_read_len);
    p->slots.resize((size_t)n_slots);
    p->launches0 = af_kernel_launches();
    p->wire_launches0 = g_wire_launches;
    cudaError_t e = cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_packed, (size_t)lay.packed_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ws, p->ws_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_lens, (size_t)slot_pairs * 2 * sizeof(uint16_t));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_nids, (size_t)p->ncap * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_nmask, (size_t)p->ncap * AF_NMASK_WORDS * 4);
        s.d_counts.assign(1, nullptr); s.d_hits.assign(1, nullptr);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_counts[0], AF_N_COUNTS * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_hits[0], (size_t)p->hits_cap * sizeof(af_hit_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&s.h_counts, (size_t)MAX_GENES * AF_N_COUNTS * 4, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(&s.h_nids, (size_t)p->ncap * 4, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(&s.h_hits, (size_t)p->hits_cap * sizeof(af_hit_t), cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        af_set_error("af_pipeline_create: %s", cudaGetErrorString(e));
        af_pipeline_free(p);
        return AF_ERR_CUDA;
    }
    *out = p;
    return AF_OK;
}

extern "C" int64_t af_pipeline_launches(const af_pipeline_t *p) { return p ? af_kernel_launches() - p->launches0 + (g_wire_launches - p->wire_launches0) : 0; }
extern "C" int64_t af_pipeline_h2d_bytes(const af_pipeline_t *p) { return p ? p->h2d_bytes : 0; }

// wait for a slot, append its hits (rebased to batch read ids) to each anchor index's list
static int collect(af_pipeline *p, Slot &s, af_hit_t *const *h_hits, const int64_t *hits_cap, int64_t *n_hits, int64_t *n_flagged) {
    if (!s.busy) return AF_OK;
    AF_CUDA(cudaEventSynchronize(s.done));
    s.busy = false;
    for (int g = 0; g < s.n_genes; g++) {
        const uint32_t *c = s.h_counts + (size_t)g * AF_N_COUNTS;
        const uint32_t status = c[AF_CNT_STATUS], nh = c[AF_CNT_HITS];
        n_flagged[g] += c[AF_CNT_FLAGGED];
        if (status) { af_set_error("af_pipeline_run: device capacity overflow (status %u)", status); return AF_ERR_CAPACITY; }
        if (nh == 0) continue;
        if (n_hits[g] + nh > hits_cap[g]) { af_set_error("af_pipeline_run: hit buffer holds %lld records, need more", (long long)hits_cap[g]); return AF_ERR_CAPACITY; }
        AF_CUDA(cudaMemcpyAsync(s.h_hits, s.d_hits[g], (size_t)nh * sizeof(af_hit_t), cudaMemcpyDeviceToHost, s.st));
        AF_CUDA(cudaStreamSynchronize(s.st));
        const uint32_t base = (uint32_t)(2 * s.first_pair);
        af_hit_t *dst = h_hits[g] + n_hits[g];
        for (uint32_t i = 0; i < nh; i++) { dst[i] = s.h_hits[i]; dst[i].read_id += base; }
        n_hits[g] += nh;
    }
    (void)p;
    return AF_OK;
}

// After an error some slots may still have copies and kernels in flight that read the caller's host
// buffers and would otherwise be collected -- with the old run's read-id rebase -- by the next run.
// Wait for them and forget their results.
static void drain_all(af_pipeline *p) {
    for (Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        s.busy = false;
    }
}

static int pipeline_run_impl(af_pipeline_t *p, int n_genes, const af_dev_index_t *const *idx, const af_batch_t *hb,
                             af_hit_t *const *h_hits, const int64_t *hits_cap, int64_t *n_hits_out, int64_t *n_flagged_out,
                             bool wire = false, int32_t pad_byte = 0);

extern "C" int af_pipeline_run_wire(af_pipeline_t *p, const af_batch_t *hb, int32_t pad_byte, af_hit_t *h_hits, int64_t hits_cap,
                                    int64_t *n_hits_out, int64_t *n_flagged_out) {
    if (!p || !n_hits_out) { af_set_error("af_pipeline_run_wire: null argument"); return AF_ERR_ARG; }
    int64_t nf = 0;
    const int rc = pipeline_run_impl(p, 1, &p->d, hb, &h_hits, &hits_cap, n_hits_out, &nf, true, pad_byte);
    if (rc != AF_OK) drain_all(p);
    else if (n_flagged_out) *n_flagged_out = nf;
    return rc;
}

extern "C" int af_pipeline_run(af_pipeline_t *p, const af_batch_t *hb, af_hit_t *h_hits, int64_t hits_cap,
                               int64_t *n_hits_out, int64_t *n_flagged_out) {
    if (!p || !n_hits_out) { af_set_error("af_pipeline_run: null argument"); return AF_ERR_ARG; }
    int64_t nf = 0;
    const int rc = pipeline_run_impl(p, 1, &p->d, hb, &h_hits, &hits_cap, n_hits_out, &nf);
    if (rc != AF_OK) drain_all(p);             // keeps af_last_error() of the failing call
    else if (n_flagged_out) *n_flagged_out = nf;
    return rc;
}

// One pass over a HOST batch for several anchor indexes (all uploaded to the pipeline's device): every
// chunk is copied to the GPU ONCE and scanned for each index in turn while it is resident, so the
// host->device traffic does not depend on the number of anchored genes (the reference re-reads both
// FASTQ files once per gene, Anchored_Fusion.py:126,182).
extern "C" int af_pipeline_run_multi(af_pipeline_t *p, int32_t n_indexes, const af_dev_index_t *const *indexes,
                                     const af_batch_t *hb, af_hit_t *const *h_hits, const int64_t *hits_cap,
                                     int64_t *n_hits_out, int64_t *n_flagged_out) {
    if (!p || !indexes || !h_hits || !hits_cap || !n_hits_out || !n_flagged_out || n_indexes < 1 || n_indexes > MAX_GENES) {
        af_set_error("af_pipeline_run_multi: bad argument (1..%d indexes)", MAX_GENES);
        return AF_ERR_ARG;
    }
    for (int g = 0; g < n_indexes; g++)
        if (!indexes[g] || af_dev_index_device(indexes[g]) != p->device) { af_set_error("af_pipeline_run_multi: index %d is not on device %d", g, p->device); return AF_ERR_ARG; }
    const int rc = pipeline_run_impl(p, n_indexes, indexes, hb, h_hits, hits_cap, n_hits_out, n_flagged_out);
    if (rc != AF_OK) drain_all(p);
    return rc;
}

static int pipeline_run_impl(af_pipeline_t *p, int n_genes, const af_dev_index_t *const *idx, const af_batch_t *hb,
                             af_hit_t *const *h_hits, const int64_t *hits_cap, int64_t *n_hits_out, int64_t *n_flagged_out,
                             bool wire, int32_t pad_byte) {
    if (!p || !hb || !n_hits_out) { af_set_error("af_pipeline_run: null argument"); return AF_ERR_ARG; }
    for (int g = 0; g < n_genes; g++) if (hits_cap[g] > 0 && !h_hits[g]) { af_set_error("af_pipeline_run: null hit buffer"); return AF_ERR_ARG; }
    if (hb->max_read_len != p->max_read_len) { af_set_error("af_pipeline_run: batch max_read_len %d, pipeline built for %d", hb->max_read_len, p->max_read_len); return AF_ERR_ARG; }
    if (hb->n_pairs >= (1ll << 31)) { af_set_error("af_pipeline_run: a host batch holds at most 2^31-1 pairs"); return AF_ERR_ARG; }
    af_layout_t lay;
    int rc = af_layout(hb->max_read_len, hb->n_pairs, &lay);
    if (rc) return rc;
    AF_CUDA(cudaSetDevice(p->device));
    for (Slot &s : p->slots)                      // per-index result buffers, grown on first use
        while ((int)s.d_counts.size() < n_genes) {
            uint32_t *c = nullptr; af_hit_t *h = nullptr;
            AF_CUDA(cudaMalloc(&c, AF_N_COUNTS * 4));
            s.d_counts.push_back(c);
            s.d_hits.push_back(nullptr);
            AF_CUDA(cudaMalloc(&h, (size_t)p->hits_cap * sizeof(af_hit_t)));
            s.d_hits.back() = h;
        }
    std::vector<int64_t> n_hits((size_t)n_genes, 0), n_flagged((size_t)n_genes, 0);
    const size_t tile_bytes = (size_t)lay.quads_per_pair * 512;
    size_t si = 0;
    for (int64_t first = 0; first < hb->n_pairs; first += p->slot_pairs) {
        Slot &s = p->slots[si];
        si = (si + 1) % p->slots.size();
        rc = collect(p, s, h_hits, hits_cap, n_hits.data(), n_flagged.data());   // results come back in chunk order
        if (rc) return rc;
        const int64_t n = std::min<int64_t>(p->slot_pairs, hb->n_pairs - first);
        const int64_t tiles = (n + 31) / 32;
        s.first_pair = first;
        s.n_pairs = n;
        if (wire) {                                  // 4 L bits per pair cross the link; the tiles are rebuilt here
            const size_t wtile = (size_t)((4 * hb->max_read_len + 31) / 32) * 128;
            if (!s.d_wire) AF_CUDA(cudaMalloc(&s.d_wire, (size_t)(p->slot_pairs / 32) * wtile));
            AF_CUDA(cudaMemcpyAsync(s.d_wire, (const char *)hb->packed + (size_t)(first / 32) * wtile, (size_t)tiles * wtile,
                                    cudaMemcpyHostToDevice, s.st));
            p->h2d_bytes += (long long)((size_t)tiles * wtile);
            rc = wire_expand_launch(s.d_wire, hb->max_read_len, n, pad_byte, s.d_packed, s.st);
            if (rc) return rc;
        } else {
            AF_CUDA(cudaMemcpyAsync(s.d_packed, (const char *)hb->packed + (size_t)(first / 32) * tile_bytes,
                                    (size_t)tiles * tile_bytes, cudaMemcpyHostToDevice, s.st));
            p->h2d_bytes += (long long)((size_t)tiles * tile_bytes);
        }
        af_batch_t db;
        db.packed = s.d_packed;
        db.n_pairs = n;
        db.max_read_len = hb->max_read_len;
        db.uniform_len = hb->uniform_len;
        db.lens = nullptr;
        if (hb->uniform_len <= 0) {
            AF_CUDA(cudaMemcpyAsync(s.d_lens, hb->lens + 2 * first, (size_t)n * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice, s.st));
            p->h2d_bytes += (long long)((size_t)n * 2 * sizeof(uint16_t));
            db.lens = s.d_lens;
        }
        // the chunk's slice of the sorted N-read list, ids rebased to the chunk
        db.nread_ids = nullptr; db.nmask = nullptr; db.n_nreads = 0;
        if (hb->n_nreads > 0) {
            const uint32_t lo_id = (uint32_t)(2 * first), hi_id = (uint32_t)(2 * (first + n));
            const uint32_t *b = hb->nread_ids, *e = hb->nread_ids + hb->n_nreads;
            const uint32_t *lo = std::lower_bound(b, e, lo_id), *hi = std::lower_bound(b, e, hi_id);
            const int64_t cnt = hi - lo;
            if (cnt > 0) {
                for (int64_t i = 0; i < cnt; i++) s.h_nids[i] = lo[i] - lo_id;
                AF_CUDA(cudaMemcpyAsync(s.d_nids, s.h_nids, (size_t)cnt * 4, cudaMemcpyHostToDevice, s.st));
                AF_CUDA(cudaMemcpyAsync(s.d_nmask, hb->nmask + (size_t)(lo - b) * AF_NMASK_WORDS,
                                        (size_t)cnt * AF_NMASK_WORDS * 4, cudaMemcpyHostToDevice, s.st));
                p->h2d_bytes += (long long)((size_t)cnt * (4 + AF_NMASK_WORDS * 4));
                db.nread_ids = s.d_nids; db.nmask = s.d_nmask; db.n_nreads = cnt;
            }
        }
        for (int g = 0; g < n_genes; g++) {        // the chunk stays resident: one scan per anchor index
            rc = af_anchor_batch(idx[g], &db, s.d_ws, p->ws_bytes, p->cand_cap, s.d_hits[g], p->hits_cap, s.d_counts[g], s.st);
            if (rc) return rc;
            AF_CUDA(cudaMemcpyAsync(s.h_counts + (size_t)g * AF_N_COUNTS, s.d_counts[g], AF_N_COUNTS * 4, cudaMemcpyDeviceToHost, s.st));
        }
        s.n_genes = n_genes;
        AF_CUDA(cudaEventRecord(s.done, s.st));
        s.busy = true;
    }
    // drain in chunk order: continue round-robin from the oldest slot
    for (size_t k = 0; k < p->slots.size(); k++) {
        Slot &s = p->slots[(si + k) % p->slots.size()];
        rc = collect(p, s, h_hits, hits_cap, n_hits.data(), n_flagged.data());
        if (rc) return rc;
    }
    for (int g = 0; g < n_genes; g++) { n_hits_out[g] = n_hits[g]; n_flagged_out[g] = n_flagged[g]; }
    return AF_OK;
}
