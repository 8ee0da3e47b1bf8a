// af_exchange.cu -- multi-GPU hit exchange over NVLink peer memory.
//
// One process per GPU; each owns a buffer of world x n_slots log regions (cudaMalloc, exported
// with CUDA IPC, opened by every other rank).  The hit-compaction kernel of the anchoring path
// (k_hit_scatter<true>, af_kernels.cu) stores each record into region (rank, slot) of EVERY
// rank's buffer, so after the kernels of a batch have run the batch's records are already on all
// GPUs: no collective launch, no per-batch rendezvous between ranks.  This file holds the host
// side: allocation, handle exchange, reset and read-back.  The reference has no counterpart (one
// process; `samtools view` writes one BAM, Anchored_Fusion.py:194).
#include <cuda_runtime.h>

#include <cstring>

#include "af_common.h"

#define AF_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            af_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return AF_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

static_assert(sizeof(cudaIpcMemHandle_t) == AF_IPC_HANDLE_BYTES, "IPC handle size");
static_assert(sizeof(af_log_header) <= AF_LOG_HEADER_BYTES, "log header");

static inline char *region_of(const af_exchange *ex, char *buffer, int src, int slot) {
    return buffer + ((size_t)src * ex->n_slots + slot) * ex->region_bytes;
}

extern "C" int af_exchange_create(int device, int32_t rank, int32_t world, int32_t n_slots, int64_t log_cap, af_exchange_t **out) {
    if (!out || world < 1 || world > AF_MAX_PEERS || rank < 0 || rank >= world || n_slots < 1 || n_slots > 64 || log_cap < 1 || log_cap >= (1ll << 31)) {
        af_set_error("af_exchange_create: bad argument (world <= %d, log_cap < 2^31)", AF_MAX_PEERS);
        return AF_ERR_ARG;
    }
    AF_CUDA(cudaSetDevice(device));
    af_exchange *ex = new af_exchange();
    ex->device = device; ex->rank = rank; ex->world = world; ex->n_slots = n_slots; ex->log_cap = log_cap;
    ex->region_bytes = ((size_t)AF_LOG_HEADER_BYTES + (size_t)log_cap * sizeof(af_hit_t) + 255) & ~(size_t)255;
    ex->total_bytes = ex->region_bytes * (size_t)world * (size_t)n_slots;
    ex->local = nullptr; ex->state = nullptr; ex->connected = world == 1;
    memset(ex->seq, 0, sizeof ex->seq);
    for (int r = 0; r < AF_MAX_PEERS; r++) ex->peer[r] = nullptr;
    cudaError_t e = cudaMalloc((void **)&ex->local, ex->total_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ex->state, (size_t)n_slots * 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ex->state, 0, (size_t)n_slots * 4 * sizeof(unsigned long long));
    // only the headers need to start out zero
    for (int i = 0; e == cudaSuccess && i < world * n_slots; i++) e = cudaMemset(ex->local + (size_t)i * ex->region_bytes, 0, AF_LOG_HEADER_BYTES);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        af_set_error("af_exchange_create: %s", cudaGetErrorString(e));
        cudaFree(ex->local); cudaFree(ex->state);
        delete ex;
        return AF_ERR_CUDA;
    }
    ex->peer[rank] = ex->local;
    *out = ex;
    return AF_OK;
}

extern "C" void af_exchange_free(af_exchange_t *ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < ex->world; r++)
        if (r != ex->rank && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
    cudaFree(ex->local); cudaFree(ex->state);
    delete ex;
}

extern "C" int af_exchange_handle(const af_exchange_t *ex, void *handle_out) {
    if (!ex || !handle_out) { af_set_error("af_exchange_handle: null"); return AF_ERR_ARG; }
    AF_CUDA(cudaSetDevice(ex->device));
    cudaIpcMemHandle_t h;
    AF_CUDA(cudaIpcGetMemHandle(&h, ex->local));
    memcpy(handle_out, &h, sizeof h);
    return AF_OK;
}

extern "C" int af_exchange_connect(af_exchange_t *ex, const void *handles) {
    if (!ex || (!handles && ex->world > 1)) { af_set_error("af_exchange_connect: null"); return AF_ERR_ARG; }
    AF_CUDA(cudaSetDevice(ex->device));
    for (int r = 0; r < ex->world; r++) {
        if (r == ex->rank || ex->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * AF_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            af_set_error("af_exchange_connect: cannot open rank %d's buffer over CUDA IPC: %s", r, cudaGetErrorString(e));
            return AF_ERR_CUDA;
        }
        ex->peer[r] = (char *)p;
    }
    ex->connected = true;
    return AF_OK;
}

int af_exchange_sink(af_exchange *ex, int slot, int64_t pair_base, af_sink *out) {
    if (!ex || !ex->connected) { af_set_error("hit exchange is not connected"); return AF_ERR_ARG; }
    if (slot < 0 || slot >= ex->n_slots || pair_base < 0) { af_set_error("hit exchange: slot %d of %d", slot, ex->n_slots); return AF_ERR_ARG; }
    out->world = ex->world;
    out->log_cap = (uint32_t)ex->log_cap;
    out->state = ex->state + 4 * slot;
    out->seq = ex->seq[slot];
    out->pair_base = (unsigned long long)pair_base;
    for (int r = 0; r < AF_MAX_PEERS; r++) out->region[r] = r < ex->world ? region_of(ex, ex->peer[r], ex->rank, slot) : nullptr;
    return AF_OK;
}

// empties log (rank, slot) on every rank: one thread per (slot, peer)
__global__ void k_exchange_reset(af_sink s0, int n_slots, size_t region_bytes) {
    const int slot = blockIdx.x, r = threadIdx.x;
    if (slot >= n_slots) return;
    if (r < s0.world) {
        af_log_header *h = (af_log_header *)(s0.region[r] + (size_t)slot * region_bytes);
        h->tail = 0; h->status = 0; h->n_batches = 0;
    }
    if (r < 4) s0.state[4 * slot + r] = 0;
}

extern "C" int af_exchange_reset(af_exchange_t *ex, void *stream) {
    af_sink s0;
    int rc = af_exchange_sink(ex, 0, 0, &s0);
    if (rc) return rc;
    AF_CUDA(cudaSetDevice(ex->device));
    k_exchange_reset<<<ex->n_slots, 32, 0, (cudaStream_t)stream>>>(s0, ex->n_slots, ex->region_bytes);
    AF_CUDA(cudaGetLastError());
    memset(ex->seq, 0, sizeof ex->seq);
    return AF_OK;
}

extern "C" int af_exchange_read(const af_exchange_t *ex, int32_t src_rank, int32_t slot, af_hit_t *h_out, int64_t cap,
                                int64_t *n_out, uint32_t *status_out, uint32_t *n_batches_out) {
    if (!ex || !n_out || src_rank < 0 || src_rank >= ex->world || slot < 0 || slot >= ex->n_slots || cap < 0) {
        af_set_error("af_exchange_read: bad argument");
        return AF_ERR_ARG;
    }
    AF_CUDA(cudaSetDevice(ex->device));
    const char *reg = region_of(ex, ex->local, src_rank, slot);
    af_log_header h;
    AF_CUDA(cudaMemcpy(&h, reg, sizeof h, cudaMemcpyDeviceToHost));
    if (h.tail > (unsigned long long)ex->log_cap) { af_set_error("af_exchange_read: corrupt log header"); return AF_ERR_CUDA; }
    *n_out = (int64_t)h.tail;
    if (status_out) *status_out = h.status;
    if (n_batches_out) *n_batches_out = h.n_batches;
    if ((int64_t)h.tail > cap) { af_set_error("af_exchange_read: log holds %lld records, buffer %lld", (long long)h.tail, (long long)cap); return AF_ERR_CAPACITY; }
    if (h.tail) {
        if (!h_out) { af_set_error("af_exchange_read: null output"); return AF_ERR_ARG; }
        AF_CUDA(cudaMemcpy(h_out, reg + AF_LOG_HEADER_BYTES, (size_t)h.tail * sizeof(af_hit_t), cudaMemcpyDeviceToHost));
    }
    return AF_OK;
}
