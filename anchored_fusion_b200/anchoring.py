"""Host objects of the anchoring path: anchor index, packed batches, the per-GPU anchorer.

Everything computes in libafb200.so (CUDA, sm_100a) through the C ABI of
include/anchored_fusion.h; torch is used for device memory and streams only.  This replaces
the reference's `bwa index` + `bwa mem -M | samtools view -F 772` stage
(Anchored_Fusion.py:167-194; Anchored_Fusion_singlecell.py:185-231).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import HIT_DTYPE, AnchoredFusionError, Batch, IndexInfo, Layout, Params, Synth, check, lib, ptr


def default_params(**over):
    p = Params()
    lib().af_default_params(ctypes.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def layout(max_read_len, n_pairs):
    lay = Layout()
    check(lib().af_layout(max_read_len, n_pairs, ctypes.byref(lay)))
    return lay


class AnchorIndex:
    """k'-mer index of one anchored CDS, both strands (stands in for `bwa index`,
    Anchored_Fusion.py:167-172).  Host side; `upload(device)` copies it into a GPU's HBM."""

    def __init__(self, anchor_seq, params=None, kp=0):
        if isinstance(anchor_seq, str):
            anchor_seq = anchor_seq.encode()
        self.seq = bytes(anchor_seq)
        self.params = params or default_params()
        h = ctypes.c_void_p()
        check(lib().af_index_build(self.seq, len(self.seq), ctypes.byref(self.params), kp, ctypes.byref(h)))
        self._h = h
        self.info = IndexInfo()
        check(lib().af_index_info(self._h, ctypes.byref(self.info)))

    @property
    def pad_byte(self):
        return self.info.pad_byte

    @property
    def bloom(self):
        """True when the filter words hold Bloom bits (long anchor) instead of fingerprint buckets."""
        return bool(lib().af_index_filter_kind(self._h))

    def filter_words(self):
        p = lib().af_index_filter(self._h)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), (self.info.n_buckets,)).copy()

    def table_words(self):
        p = lib().af_index_table(self._h)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)),
                                     (self.info.table_slots, 2)).copy()

    def upload(self, device=0):
        return DeviceIndex(self, device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.af_index_free(h)


class DeviceIndex:
    def __init__(self, index, device):
        self.index = index
        self.device = int(device)
        h = ctypes.c_void_p()
        check(lib().af_index_upload(index._h, self.device, ctypes.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.af_dev_index_free(h)


class PackedBatch:
    """A batch of read pairs in the 2-bit tile layout (host numpy arrays or device tensors)."""

    def __init__(self, packed, n_pairs, max_read_len, uniform_len=0, lens=None, nread_ids=None, nmask=None):
        self.packed, self.n_pairs, self.max_read_len = packed, int(n_pairs), int(max_read_len)
        self.uniform_len, self.lens, self.nread_ids, self.nmask = int(uniform_len), lens, nread_ids, nmask
        self.n_nreads = 0 if nread_ids is None else int(len(nread_ids))

    def c_struct(self):
        return Batch(ptr(self.packed), self.n_pairs, self.max_read_len, self.uniform_len, ptr(self.lens),
                     ptr(self.nread_ids) if self.n_nreads else None, ptr(self.nmask) if self.n_nreads else None,
                     self.n_nreads)

    def read_len(self, read_id):
        return self.uniform_len if self.uniform_len > 0 else int(self.lens[read_id])

    def to_device(self, device):
        import torch
        dev = torch.device("cuda", device)

        def up(a):
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        nm = self.nmask.view(np.int32) if self.nmask is not None else None
        ni = self.nread_ids.view(np.int32) if self.nread_ids is not None else None
        ln = self.lens.view(np.int16) if self.lens is not None else None
        pk = self.packed.view(np.int32) if isinstance(self.packed, np.ndarray) else self.packed
        return PackedBatch(up(pk), self.n_pairs, self.max_read_len, self.uniform_len, up(ln),
                           up(ni) if self.n_nreads else None, up(nm) if self.n_nreads else None)


def pack_pairs(seqs1, seqs2, max_read_len=None, pad_byte=0xE4, out=None):
    """2-bit pack two lists of ASCII reads (bytes/str) into the tile layout (host arrays)."""
    n = len(seqs1)
    if len(seqs2) != n:
        raise ValueError("the two mates' lists differ in length")
    b1 = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs1]
    b2 = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs2]
    if max_read_len is None:
        max_read_len = max([len(s) for s in b1] + [len(s) for s in b2] + [1])
    lay = layout(max_read_len, n)
    off1 = np.zeros(n + 1, dtype=np.int64)
    off2 = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(s) for s in b1], out=off1[1:])
        np.cumsum([len(s) for s in b2], out=off2[1:])
    cat1, cat2 = b"".join(b1), b"".join(b2)
    packed = out if out is not None else np.zeros(max(lay.packed_bytes // 4, 1), dtype=np.uint32)
    lens = np.zeros(max(2 * n, 1), dtype=np.uint16)
    nids = np.zeros(max(2 * n, 1), dtype=np.uint32)
    nmask = np.zeros((max(2 * n, 1), _lib.NMASK_WORDS), dtype=np.uint32)
    nn, ulen = ctypes.c_int64(0), ctypes.c_int32(0)
    check(lib().af_pack_pairs(ctypes.cast(ctypes.c_char_p(cat1), ctypes.c_void_p), off1.ctypes.data,
                              ctypes.cast(ctypes.c_char_p(cat2), ctypes.c_void_p), off2.ctypes.data, n, max_read_len,
                              pad_byte, packed.ctypes.data, lens.ctypes.data, nids.ctypes.data, nmask.ctypes.data,
                              len(nids), ctypes.byref(nn), ctypes.byref(ulen)))
    k = nn.value
    return PackedBatch(packed, n, max_read_len, ulen.value, lens[: 2 * n],
                       nids[:k].copy() if k else None, nmask[:k].copy() if k else None)


def wire_bytes(max_read_len, n_pairs):
    return lib().af_wire_bytes(max_read_len, n_pairs)


def wire_from_packed(packed, max_read_len, n_pairs, out=None):
    """Packed tiles (uint32 array, host) -> wire format (4 * max_read_len bits per pair, include/anchored_fusion.h)."""
    packed = np.ascontiguousarray(packed).view(np.uint32)
    nbytes = wire_bytes(max_read_len, n_pairs)
    wire = out if out is not None else np.zeros(nbytes // 4, dtype=np.uint32)
    assert wire.nbytes >= nbytes
    check(lib().af_wire_from_packed(packed.ctypes.data, max_read_len, n_pairs, wire.ctypes.data))
    return wire


def wire_to_packed(wire, max_read_len, n_pairs, pad_byte):
    """Host twin of the device expansion: wire format -> packed tiles."""
    wire = np.ascontiguousarray(wire).view(np.uint32)
    packed = np.zeros(layout(max_read_len, n_pairs).packed_bytes // 4, dtype=np.uint32)
    check(lib().af_wire_to_packed(wire.ctypes.data, max_read_len, n_pairs, pad_byte, packed.ctypes.data))
    return packed


def unpack_read(batch, read_id, length=None):
    """codes 0..3 of one read of a HOST batch (test helper)."""
    length = batch.read_len(read_id) if length is None else length
    out = np.zeros(length, dtype=np.uint8)
    check(lib().af_unpack_read(ptr(batch.packed), batch.max_read_len, read_id, length, out.ctypes.data))
    return out


class Anchorer:
    """The hot path on one GPU: seed scan -> compaction -> verify/extend -> compaction.

    Device-resident entry (`anchor`) takes a PackedBatch whose arrays are CUDA tensors;
    `anchor_host` streams a host batch through the C++ pipeline (pinned staging,
    cudaMemcpyAsync, several chunks in flight).
    """

    def __init__(self, index, device=0):
        import torch
        if not torch.cuda.is_available():
            raise AnchoredFusionError("no CUDA device: the anchoring path has no CPU fallback")
        self.torch = torch
        self.index = index
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        self.dindex = index.upload(self.device)
        self._ws = {}
        self._pipe = None
        self._pipe_key = None

    # -- device-resident path -----------------------------------------------------------
    def _workspace(self, n_pairs, cand_cap, hits_cap, slot=0, max_read_len=256):
        key = (n_pairs, cand_cap, hits_cap, max_read_len)
        cur = self._ws.get(slot) if isinstance(self._ws, dict) else None
        if cur is None or cur[0] != key:
            torch = self.torch
            if not isinstance(self._ws, dict):
                self._ws = {}
            nbytes = lib().af_workspace_bytes_len(n_pairs, cand_cap, max_read_len)
            # counts (8 x int32 = 2 rows) and the hit records share one tensor, so that the multi-GPU
            # gather can ship (count, first records) as one contiguous slice with no packing kernel
            hc = torch.zeros((2 + max(hits_cap, 1), 4), dtype=torch.int32, device=self.dev)
            cur = (key, torch.empty(nbytes, dtype=torch.uint8, device=self.dev), hc[2:], hc[:2].view(-1),
                   torch.cuda.Stream(device=self.dev) if slot else None, hc)
            self._ws[slot] = cur
        return cur[1:5]

    def counts_and_hits(self, slot=0):
        """The slot's [2 + cap, 4] int32 tensor: rows 0..1 are the counters, rows 2.. the hit records."""
        return self._ws[slot][5]

    def slot_stream(self, slot):
        """The side stream of workspace slot `slot` (> 0), created on first use."""
        return self._ws[slot][4]

    def enqueue(self, batch, cand_cap=None, hits_cap=None, stream=None, slot=0, exchange=None, pair_base=0):
        """Launch the whole path on the current (or given) stream; no host sync.
        `slot` selects an independent workspace so that consecutive batches can be in flight on
        different streams (slot > 0 owns a side stream, used when `stream` is None).
        exchange: a dist.HitExchange -- the batch's records are also appended (by the hit-compaction
        kernel itself, NVLink peer stores) to this rank's log `slot` on every rank, tagged `pair_base`.
        Returns (hits tensor [cap,4] int32 raw records, counts tensor)."""
        torch = self.torch
        n = batch.n_pairs
        cand_cap = int(cand_cap or 2 * max(n, 1))
        hits_cap = int(hits_cap or cand_cap)
        ws, hits, counts, side = self._workspace(n, cand_cap, hits_cap, slot, batch.max_read_len)
        st = stream if stream is not None else (side if side is not None else torch.cuda.current_stream(self.dev))
        cb = batch.c_struct()
        if exchange is not None:
            check(lib().af_anchor_batch_exchange(self.dindex._h, ctypes.byref(cb), ws.data_ptr(), ws.numel(), cand_cap,
                                                 hits.data_ptr(), hits_cap, counts.data_ptr(), exchange._h, slot,
                                                 int(pair_base), ctypes.c_void_p(st.cuda_stream)))
        else:
            check(lib().af_anchor_batch(self.dindex._h, ctypes.byref(cb), ws.data_ptr(), ws.numel(), cand_cap,
                                        hits.data_ptr(), hits_cap, counts.data_ptr(), ctypes.c_void_p(st.cuda_stream)))
        return hits, counts

    def anchor(self, batch, cand_cap=None, hits_cap=None):
        """Anchor a device-resident batch; returns (hits ndarray HIT_DTYPE ordered by read_id, stats)."""
        hits, counts = self.enqueue(batch, cand_cap, hits_cap)
        c = counts.cpu().numpy().view(np.uint32)
        if c[_lib.CNT_STATUS]:
            raise AnchoredFusionError("device capacity overflow (status %d): raise cand_cap/hits_cap" % c[_lib.CNT_STATUS])
        nh = int(c[_lib.CNT_HITS])
        out = hits[:nh].cpu().numpy().view(np.uint8).reshape(-1).view(HIT_DTYPE).copy() if nh else np.zeros(0, HIT_DTYPE)
        return out, {"flagged": int(c[_lib.CNT_FLAGGED]), "hits": nh}

    def seed_scan(self, batch, stream=None):
        """Only the seed-scan kernel; returns the flag words tensor [n_tiles, 2] (int32)."""
        torch = self.torch
        lay = layout(batch.max_read_len, batch.n_pairs)
        flags = torch.zeros((max(lay.n_tiles, 1), 2), dtype=torch.int32, device=self.dev)
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        cb = batch.c_struct()
        check(lib().af_seed_scan(self.dindex._h, ctypes.byref(cb), flags.data_ptr(), ctypes.c_void_p(st.cuda_stream)))
        return flags

    # -- host-buffer path (end to end) ---------------------------------------------------
    def pipeline(self, max_read_len, slot_pairs=1 << 20, n_slots=3):
        key = (max_read_len, slot_pairs, n_slots)
        if self._pipe_key != key:
            self.close_pipeline()
            h = ctypes.c_void_p()
            check(lib().af_pipeline_create(self.dindex._h, slot_pairs, max_read_len, n_slots, ctypes.byref(h)))
            self._pipe, self._pipe_key = h, key
        return self._pipe

    def pipeline_h2d_bytes(self):
        """bytes the current host-buffer pipeline has copied to the GPU so far"""
        return lib().af_pipeline_h2d_bytes(self._pipe) if self._pipe else 0

    def close_pipeline(self):
        if self._pipe and _lib._lib is not None:
            _lib._lib.af_pipeline_free(self._pipe)
        self._pipe, self._pipe_key = None, None

    def anchor_host(self, batch, slot_pairs=1 << 20, n_slots=3, hits_out=None, wire=False):
        """Anchor a HOST batch (numpy / pinned arrays): H2D copies, kernels and the D2H of the
        hit list all happen inside this call.  Returns (hits ndarray, stats).  wire=True: batch.packed holds
        the wire format (`wire_from_packed`), 4 L bits per pair instead of whole tiles; it is expanded on the GPU."""
        pipe = self.pipeline(batch.max_read_len, slot_pairs, n_slots)
        cap = 2 * max(batch.n_pairs, 1) if hits_out is None else len(hits_out)
        out = hits_out if hits_out is not None else np.zeros(cap, dtype=HIT_DTYPE)
        nh, nf = ctypes.c_int64(0), ctypes.c_int64(0)
        cb = batch.c_struct()
        if wire:
            check(lib().af_pipeline_run_wire(pipe, ctypes.byref(cb), self.index.pad_byte, out.ctypes.data, cap, ctypes.byref(nh), ctypes.byref(nf)))
        else:
            check(lib().af_pipeline_run(pipe, ctypes.byref(cb), out.ctypes.data, cap, ctypes.byref(nh), ctypes.byref(nf)))
        return out[: nh.value], {"flagged": nf.value, "hits": nh.value}

    def __del__(self):
        try:
            self.close_pipeline()
        except Exception:
            pass


# ---- synthetic reads (measurement inputs) ---------------------------------------
def synth_spec(seed=1, ref_len=10_000_000, anchor_start=1_000_000, anchor_len=6783, read_len=150, frag_mean=300,
               frag_sd=30, sub_ppm=0, fusion_ppm=0, n_ppm=0):
    return Synth(seed, ref_len, anchor_start, anchor_len, read_len, frag_mean, frag_sd, sub_ppm, fusion_ppm, n_ppm, 0)


def synth_anchor(spec):
    buf = ctypes.create_string_buffer(spec.anchor_len)
    check(lib().af_synth_anchor(ctypes.byref(spec), ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw[: spec.anchor_len]


def synth_pairs_host(spec, first_pair, n_pairs):
    """(mate1, mate2) uint8 code arrays [n_pairs, read_len], codes 0..4."""
    m1 = np.zeros((n_pairs, spec.read_len), dtype=np.uint8)
    m2 = np.zeros((n_pairs, spec.read_len), dtype=np.uint8)
    check(lib().af_synth_pairs_host(ctypes.byref(spec), first_pair, n_pairs, m1.ctypes.data, m2.ctypes.data))
    return m1, m2


def synth_pairs_device(spec, first_pair, n_pairs, pad_byte, device=0, stream=None):
    """Generate packed tiles straight into HBM; returns a device PackedBatch."""
    import torch
    dev = torch.device("cuda", device)
    lay = layout(spec.read_len, n_pairs)
    packed = torch.empty(max(lay.packed_bytes // 4, 1), dtype=torch.int32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    check(lib().af_synth_pairs_device(ctypes.byref(spec), first_pair, n_pairs, pad_byte, packed.data_ptr(),
                                      ctypes.c_void_p(st.cuda_stream)))
    return PackedBatch(packed, n_pairs, spec.read_len, spec.read_len)


def codes_to_ascii(codes):
    return np.frombuffer(b"ACGTN", dtype=np.uint8)[codes].tobytes()
