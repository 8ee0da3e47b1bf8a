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


struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    void *d_packed = nullptr, *d_ws = nullptr;
    uint16_t *d_lens = nullptr;
    uint32_t *d_nids = nullptr, *d_nmask = nullptr, *d_counts = nullptr;
    af_hit_t *d_hits = nullptr;
    uint32_t *h_counts = nullptr;   // pinned
    uint32_t *h_nids = nullptr;     // pinned staging of rebased N-read ids
    af_hit_t *h_hits = nullptr;     // pinned
    bool busy = false;
    int64_t first_pair = 0, n_pairs = 0;
};

struct af_pipeline {
    const af_dev_index_t *d = nullptr;
    int device = 0;
    int64_t slot_pairs = 0, cand_cap = 0, hits_cap = 0, ncap = 0;
    int32_t max_read_len = 0;
    size_t ws_bytes = 0;
    std::vector<Slot> slots;
    long long launches0 = 0;
};


extern "C" void af_pipeline_free(af_pipeline_t *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        cudaFree(s.d_packed); cudaFree(s.d_ws); cudaFree(s.d_lens); cudaFree(s.d_nids); cudaFree(s.d_nmask);
        cudaFree(s.d_counts); cudaFree(s.d_hits);
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
    p->ws_bytes = af_workspace_bytes(slot_pairs, p->cand_cap);
    p->slots.resize((size_t)n_slots);
    p->launches0 = af_kernel_launches();
    cudaError_t e = cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_packed, (size_t)lay.packed_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ws, p->ws_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_lens, (size_t)slot_pairs * 2 * sizeof(uint16_t));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_nids, (size_t)p->ncap * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_nmask, (size_t)p->ncap * AF_NMASK_WORDS * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_counts, AF_N_COUNTS * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_hits, (size_t)p->hits_cap * sizeof(af_hit_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&s.h_counts, AF_N_COUNTS * 4, cudaHostAllocDefault);
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

extern "C" int64_t af_pipeline_launches(const af_pipeline_t *p) { return p ? af_kernel_launches() - p->launches0 : 0; }

// wait for a slot, append its hits (rebased to batch read ids)
static int collect(af_pipeline *p, Slot &s, af_hit_t *h_hits, int64_t hits_cap, int64_t &n_hits, int64_t &n_flagged) {
    if (!s.busy) return AF_OK;
    AF_CUDA(cudaEventSynchronize(s.done));
    s.busy = false;
    uint32_t status = s.h_counts[AF_CNT_STATUS], nh = s.h_counts[AF_CNT_HITS];
    n_flagged += s.h_counts[AF_CNT_FLAGGED];
    if (status) { af_set_error("af_pipeline_run: device capacity overflow (status %u)", status); return AF_ERR_CAPACITY; }
    if (nh == 0) return AF_OK;
    if (n_hits + nh > hits_cap) { af_set_error("af_pipeline_run: hit buffer holds %lld records, need more", (long long)hits_cap); return AF_ERR_CAPACITY; }
    AF_CUDA(cudaMemcpyAsync(s.h_hits, s.d_hits, (size_t)nh * sizeof(af_hit_t), cudaMemcpyDeviceToHost, s.st));
    AF_CUDA(cudaStreamSynchronize(s.st));
    const uint32_t base = (uint32_t)(2 * s.first_pair);
    for (uint32_t i = 0; i < nh; i++) { h_hits[n_hits + i] = s.h_hits[i]; h_hits[n_hits + i].read_id += base; }
    n_hits += nh;
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

static int pipeline_run_impl(af_pipeline_t *p, const af_batch_t *hb, af_hit_t *h_hits, int64_t hits_cap,
                             int64_t *n_hits_out, int64_t *n_flagged_out);

extern "C" int af_pipeline_run(af_pipeline_t *p, const af_batch_t *hb, af_hit_t *h_hits, int64_t hits_cap,
                               int64_t *n_hits_out, int64_t *n_flagged_out) {
    const int rc = pipeline_run_impl(p, hb, h_hits, hits_cap, n_hits_out, n_flagged_out);
    if (rc != AF_OK && p) drain_all(p);        // keeps af_last_error() of the failing call
    return rc;
}

static int pipeline_run_impl(af_pipeline_t *p, const af_batch_t *hb, af_hit_t *h_hits, int64_t hits_cap,
                             int64_t *n_hits_out, int64_t *n_flagged_out) {
    if (!p || !hb || !n_hits_out || (hits_cap > 0 && !h_hits)) { af_set_error("af_pipeline_run: null argument"); return AF_ERR_ARG; }
    if (hb->max_read_len != p->max_read_len) { af_set_error("af_pipeline_run: batch max_read_len %d, pipeline built for %d", hb->max_read_len, p->max_read_len); return AF_ERR_ARG; }
    if (hb->n_pairs >= (1ll << 31)) { af_set_error("af_pipeline_run: a host batch holds at most 2^31-1 pairs"); return AF_ERR_ARG; }
    af_layout_t lay;
    int rc = af_layout(hb->max_read_len, hb->n_pairs, &lay);
    if (rc) return rc;
    AF_CUDA(cudaSetDevice(p->device));
    int64_t n_hits = 0, n_flagged = 0;
    const size_t tile_bytes = (size_t)lay.quads_per_pair * 512;
    size_t si = 0;
    for (int64_t first = 0; first < hb->n_pairs; first += p->slot_pairs) {
        Slot &s = p->slots[si];
        si = (si + 1) % p->slots.size();
        rc = collect(p, s, h_hits, hits_cap, n_hits, n_flagged);   // results come back in chunk order
        if (rc) return rc;
        const int64_t n = std::min<int64_t>(p->slot_pairs, hb->n_pairs - first);
        const int64_t tiles = (n + 31) / 32;
        s.first_pair = first;
        s.n_pairs = n;
        AF_CUDA(cudaMemcpyAsync(s.d_packed, (const char *)hb->packed + (size_t)(first / 32) * tile_bytes,
                                (size_t)tiles * tile_bytes, cudaMemcpyHostToDevice, s.st));
        af_batch_t db;
        db.packed = s.d_packed;
        db.n_pairs = n;
        db.max_read_len = hb->max_read_len;
        db.uniform_len = hb->uniform_len;
        db.lens = nullptr;
        if (hb->uniform_len <= 0) {
            AF_CUDA(cudaMemcpyAsync(s.d_lens, hb->lens + 2 * first, (size_t)n * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice, s.st));
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
                db.nread_ids = s.d_nids; db.nmask = s.d_nmask; db.n_nreads = cnt;
            }
        }
        rc = af_anchor_batch(p->d, &db, s.d_ws, p->ws_bytes, p->cand_cap, s.d_hits, p->hits_cap, s.d_counts, s.st);
        if (rc) return rc;
        AF_CUDA(cudaMemcpyAsync(s.h_counts, s.d_counts, AF_N_COUNTS * 4, cudaMemcpyDeviceToHost, s.st));
        AF_CUDA(cudaEventRecord(s.done, s.st));
        s.busy = true;
    }
    // drain in chunk order: continue round-robin from the oldest slot
    for (size_t k = 0; k < p->slots.size(); k++) {
        Slot &s = p->slots[(si + k) % p->slots.size()];
        rc = collect(p, s, h_hits, hits_cap, n_hits, n_flagged);
        if (rc) return rc;
    }
    *n_hits_out = n_hits;
    if (n_flagged_out) *n_flagged_out = n_flagged;
    return AF_OK;
}
