"""Multi-GPU sharding of the anchoring pass: one process per GPU, read pairs dealt to ranks,
no data-path collective; the only exchange is the small hit-record lists (the reference has no
collective at all -- SURVEY.md 2, 8e).  Two ways to exchange them:

* HitExchange (GPUs of one box): the hit-compaction kernel stores every record into a log on
  every rank through NVLink peer memory (CUDA IPC), so nothing is launched or awaited per batch;
* gather_hits_tensor: one all-gather per batch; backend "nccl" (device tensors) or "gloo" (CPU
  tensors, tests)."""
import ctypes

import numpy as np

from ._lib import HIT_DTYPE

LOG_MARKER = 0xFFFFFFFF
IPC_HANDLE_BYTES = 64
STATUS_LOG_OVERFLOW = 4


def bind_near_gpu(device_index):
    """Pin this process to the CPUs NVML reports as local to the GPU (same socket / NUMA node), so that
    the pinned staging buffers it allocates next are first-touched on that node and the H2D copies of
    eight ranks do not all cross one socket's memory controller.  Returns the CPU set, or None when
    NVML cannot tell (then nothing is changed)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device_index)
        bus = "%08X:%02X:%02X.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def shard_range(n_pairs, rank, world):
    """Contiguous, tile-aligned share of [0, n_pairs) for `rank`."""
    tiles = (n_pairs + 31) // 32
    lo = (tiles * rank // world) * 32
    hi = (tiles * (rank + 1) // world) * 32
    return min(lo, n_pairs), min(hi, n_pairs)


def gather_hits_tensor(counts_and_hits, cap, group=None):
    """One all-gather of (counters, first `cap` records) per rank.  counts_and_hits: the engine's
    [2 + capacity, 4] int32 tensor (Anchorer.counts_and_hits): rows 0..1 hold the 8 counters (hit count
    at flat index 1, AF_CNT_HITS), rows 2.. the records -- contiguous, so nothing is packed first.
    Returns (all_counts [world] int32, all_hits [world, cap, 4] int32) tensors, asynchronously on
    the current stream for NCCL."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = counts_and_hits[: 2 + cap]
    out = torch.empty((world, 2 + cap, 4), dtype=torch.int32, device=src.device)
    dist.all_gather_into_tensor(out.view(-1, 4), src, group=group)
    return out[:, 0, 1], out[:, 2:]


def merge_gathered(all_counts, all_hits, pair_offsets):
    """Host side: concatenate the ranks' records with read_ids made global
    (read_id += 2 * first pair of the rank's shard); raises if a rank overflowed the gather cap."""
    counts = all_counts.cpu().numpy()
    recs = all_hits.cpu().numpy()
    cap = recs.shape[1]
    parts = []
    for r, c in enumerate(counts):
        if c > cap:
            raise RuntimeError("rank %d holds %d hits, gather cap is %d" % (r, c, cap))
        h = np.ascontiguousarray(recs[r, :c]).view(np.uint8).reshape(-1).view(HIT_DTYPE).copy()
        h["read_id"] += np.uint32(2 * pair_offsets[r])
        parts.append(h)
    return np.concatenate(parts) if parts else np.zeros(0, HIT_DTYPE)


def parse_log(records):
    """Split one log region (HIT_DTYPE array: markers + records, include/anchored_fusion.h) into
    [(pair_base, hits)] in append order.  A marker is read_id 0xFFFFFFFF, pos / clip_l|m_len the low /
    high 32 bits of pair_base, last word the record count."""
    raw = np.ascontiguousarray(records).view(np.uint32).reshape(-1, 4)
    out, i, n = [], 0, len(raw)
    while i < n:
        if raw[i, 0] != LOG_MARKER:
            raise ValueError("hit log: record %d should be a batch marker" % i)
        base = int(raw[i, 1]) | (int(raw[i, 2]) << 32)
        cnt = int(raw[i, 3])
        if i + 1 + cnt > n:
            raise ValueError("hit log: batch at %d claims %d records, %d left (log overflow?)" % (i, cnt, n - i - 1))
        out.append((base, records[i + 1: i + 1 + cnt]))
        i += 1 + cnt
    return out


def globalise(batches):
    """[(pair_base, hits)] -> one HIT_DTYPE-like array with int64 read_ids (read_id + 2 * pair_base),
    ordered by read_id."""
    dt = np.dtype([("read_id", "<i8")] + [(n, HIT_DTYPE[n]) for n in HIT_DTYPE.names if n != "read_id"])
    parts = []
    for base, h in batches:
        g = np.zeros(len(h), dt)
        for name in HIT_DTYPE.names:
            g[name] = h[name]
        g["read_id"] += 2 * base
        parts.append(g)
    allh = np.concatenate(parts) if parts else np.zeros(0, dt)
    return allh[np.argsort(allh["read_id"], kind="stable")]


class HitExchange:
    """Per-rank log buffers shared over CUDA IPC (include/anchored_fusion.h, af_exchange_*).

        ex = HitExchange(rank, world, n_slots, log_cap, device)     # collective: swaps IPC handles
        eng.enqueue(batch, slot=s, exchange=ex, pair_base=first_pair_of_batch)   # any number of times
        batches = ex.collect()     # collective: stream sync + barrier, then [(src_rank, pair_base, hits)]
    """

    def __init__(self, rank, world, n_slots, log_cap, device, group=None):
        import torch.distributed as dist
        from ._lib import check, lib
        self._h = ctypes.c_void_p()
        self.rank, self.world, self.n_slots, self.log_cap, self.group = rank, world, n_slots, int(log_cap), group
        dev = device.index if hasattr(device, "index") else int(device)
        self.device = dev
        check(lib().af_exchange_create(dev, rank, world, n_slots, self.log_cap, ctypes.byref(self._h)))
        mine = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        check(lib().af_exchange_handle(self._h, mine))
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, mine.raw, group=group)
            check(lib().af_exchange_connect(self._h, b"".join(handles)))
            dist.barrier(group=group)

    def reset(self, stream=None):
        """Empty this rank's logs on every rank.  Collective: the ranks meet before and after."""
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(lib().af_exchange_reset(self._h, ctypes.c_void_p(st.cuda_stream)))
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def read(self, src_rank, slot):
        """(records incl. markers, status, n_batches) of log (src_rank, slot) as this rank holds it."""
        from ._lib import check, lib
        n, status, nb = ctypes.c_int64(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
        out = np.zeros(self.log_cap, HIT_DTYPE)
        check(lib().af_exchange_read(self._h, src_rank, slot, out.ctypes.data, self.log_cap, ctypes.byref(n),
                                     ctypes.byref(status), ctypes.byref(nb)))
        return out[: n.value], status.value, nb.value

    def collect(self):
        """Wait for every rank's kernels, then return [(src_rank, pair_base, hits)] of all logs."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        out = []
        for r in range(self.world):
            for s in range(self.n_slots):
                recs, status, _ = self.read(r, s)
                if status & STATUS_LOG_OVERFLOW:
                    raise RuntimeError("hit log (rank %d, slot %d) overflowed its %d records" % (r, s, self.log_cap))
                out += [(r, base, h) for base, h in parse_log(recs)]
        return out

    def close(self):
        from . import _lib
        if self._h and _lib._lib is not None:
            _lib._lib.af_exchange_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
