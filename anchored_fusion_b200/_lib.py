"""ctypes binding of libafb200.so (the C ABI declared in include/anchored_fusion.h).

There is no fallback: if the CUDA library cannot be loaded the import of the product fails
loudly.  The library is built in-tree by anchored_fusion_b200/build.py (nvcc, sm_100a).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafb200.so")

c_i32, c_i64, c_u32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p


class Params(ctypes.Structure):
    """af_params_t: bwa-mem defaults the reference relies on (Anchored_Fusion.py:182)."""
    _fields_ = [(n, c_i32) for n in ("k", "A", "B", "clip5", "clip3", "T", "X")]


class IndexInfo(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("anchor_len", "k", "kp", "stride", "n_keys", "n_entries", "n_buckets",
                                      "n_overflow", "table_slots")] + [("filter_mul", c_u32), ("pad_byte", c_i32)]


class Layout(ctypes.Structure):
    _fields_ = [("max_read_len", c_i32), ("words_per_read", c_i32), ("quads_per_pair", c_i32), ("reserved", c_i32),
                ("n_pairs", c_i64), ("n_tiles", c_i64), ("packed_bytes", c_i64)]


class Batch(ctypes.Structure):
    _fields_ = [("packed", c_vp), ("n_pairs", c_i64), ("max_read_len", c_i32), ("uniform_len", c_i32),
                ("lens", c_vp), ("nread_ids", c_vp), ("nmask", c_vp), ("n_nreads", c_i64)]


class GenomeStats(ctypes.Structure):
    """af_genome_stats_t"""
    _fields_ = [("genome_bases", c_i64), ("n_candidates", c_i64), ("n_seeds", c_i64), ("n_passes", c_i32), ("n_retries", c_i32),
                ("scan_ms", ctypes.c_double), ("total_ms", ctypes.c_double), ("host_index_ms", ctypes.c_double)]


class Synth(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("ref_len", c_i64), ("anchor_start", c_i64), ("anchor_len", c_i32),
                ("read_len", c_i32), ("frag_mean", c_i32), ("frag_sd", c_i32), ("sub_ppm", c_u32),
                ("fusion_ppm", c_u32), ("n_ppm", c_u32), ("reserved", c_u32)]


HIT_DTYPE = np.dtype([("read_id", "<u4"), ("pos", "<i4"), ("clip_l", "<u2"), ("m_len", "<u2"),
                      ("clip_r", "<u2"), ("score_strand", "<u2")])
GENOME_HIT_DTYPE = np.dtype([("pos", "<i8"), ("read_id", "<u4"), ("clip_l", "<u2"), ("m_len", "<u2"), ("clip_r", "<u2"),
                             ("score_strand", "<u2"), ("reserved", "<u4")])
GENOME_SEP = 256
CNT_FLAGGED, CNT_HITS, CNT_STATUS, CNT_SEEDED, N_COUNTS = 0, 1, 2, 3, 8
NMASK_WORDS = 16
MAX_READ_LEN = 512
GENOME_MAX_READ_LEN = 256

# every symbol include/anchored_fusion.h declares: name -> (restype, argtypes)
P = ctypes.POINTER
SIGNATURES = {
    "af_last_error": (ctypes.c_char_p, []),
    "af_abi_version": (ctypes.c_int, []),
    "af_default_params": (None, [P(Params)]),
    "af_index_build": (ctypes.c_int, [ctypes.c_char_p, c_i64, P(Params), c_i32, P(c_vp)]),
    "af_index_free": (None, [c_vp]),
    "af_index_info": (ctypes.c_int, [c_vp, P(IndexInfo)]),
    "af_index_filter_kind": (ctypes.c_int, [c_vp]),
    "af_index_filter": (c_vp, [c_vp]),
    "af_index_table": (c_vp, [c_vp]),
    "af_index_upload": (ctypes.c_int, [c_vp, ctypes.c_int, P(c_vp)]),
    "af_dev_index_free": (None, [c_vp]),
    "af_dev_index_device": (ctypes.c_int, [c_vp]),
    "af_layout": (ctypes.c_int, [c_i32, c_i64, P(Layout)]),
    "af_pack_pairs": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64,
                                     P(c_i64), P(c_i32)]),
    "af_unpack_read": (ctypes.c_int, [c_vp, c_i32, c_i64, c_i32, c_vp]),
    "af_fastq_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, P(c_vp)]),
    "af_fastq_open_threads": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, c_i32, P(c_vp)]),
    "af_fastq_open_multi": (ctypes.c_int, [P(ctypes.c_char_p), P(ctypes.c_char_p), c_i32, c_i32, P(c_vp)]),
    "af_fastq_threads": (ctypes.c_int, [c_vp]),
    "af_fastq_skip": (ctypes.c_int, [c_vp, c_i64, P(c_i64)]),
    "af_fastq_records": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, P(c_i64)]),
    "af_fastq_file_starts": (ctypes.c_int, [c_vp, c_vp, c_i32]),
    "af_fastq_batch_first_pair": (c_i64, [c_vp]),
    "af_fastq_peek": (ctypes.c_int, [ctypes.c_char_p, c_i32, P(c_i32)]),
    "af_fastq_close": (None, [c_vp]),
    "af_debug_crc32": (c_u32, [c_vp, c_i64]),
    "af_debug_gunzip_chunks": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i64, P(c_i64), P(c_i32)]),
    "af_fastq_next": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, P(c_i64), P(c_i32),
                                     P(c_i64)]),
    "af_fastq_record": (ctypes.c_int, [c_vp, c_i64, P(c_vp), P(c_i32), P(c_vp), P(c_vp), P(c_i32)]),
    "af_workspace_bytes": (ctypes.c_size_t, [c_i64, c_i64]),
    "af_workspace_bytes_len": (ctypes.c_size_t, [c_i64, c_i64, c_i32]),
    "af_anchor_batch": (ctypes.c_int, [c_vp, P(Batch), c_vp, ctypes.c_size_t, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "af_seed_scan": (ctypes.c_int, [c_vp, P(Batch), c_vp, c_vp]),
    "af_debug_scan_pair": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, P(c_i32), P(c_i32)]),
    "af_seed_scan_config": (ctypes.c_int, [c_i32, c_i32]),
    "af_kernel_launches": (c_i64, []),
    "af_profile_begin": (None, []),
    "af_profile_end": (ctypes.c_int, [c_vp, c_vp]),
    "af_pipeline_create": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, P(c_vp)]),
    "af_pipeline_free": (None, [c_vp]),
    "af_pipeline_run": (ctypes.c_int, [c_vp, P(Batch), c_vp, c_i64, P(c_i64), P(c_i64)]),
    "af_pipeline_run_multi": (ctypes.c_int, [c_vp, c_i32, c_vp, P(Batch), c_vp, c_vp, c_vp, c_vp]),
    "af_wire_bytes": (c_i64, [c_i32, c_i64]),
    "af_wire_from_packed": (ctypes.c_int, [c_vp, c_i32, c_i64, c_vp]),
    "af_wire_to_packed": (ctypes.c_int, [c_vp, c_i32, c_i64, c_i32, c_vp]),
    "af_wire_expand_device": (ctypes.c_int, [c_vp, c_i32, c_i64, c_i32, c_vp, c_vp]),
    "af_pipeline_run_wire": (ctypes.c_int, [c_vp, P(Batch), c_i32, c_vp, c_i64, P(c_i64), P(c_i64)]),
    "af_pipeline_launches": (c_i64, [c_vp]),
    "af_pipeline_h2d_bytes": (c_i64, [c_vp]),
    "af_exchange_create": (ctypes.c_int, [ctypes.c_int, c_i32, c_i32, c_i32, c_i64, P(c_vp)]),
    "af_exchange_free": (None, [c_vp]),
    "af_exchange_handle": (ctypes.c_int, [c_vp, c_vp]),
    "af_exchange_connect": (ctypes.c_int, [c_vp, c_vp]),
    "af_exchange_reset": (ctypes.c_int, [c_vp, c_vp]),
    "af_anchor_batch_exchange": (ctypes.c_int, [c_vp, P(Batch), c_vp, ctypes.c_size_t, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_i64, c_vp]),
    "af_exchange_read": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_i64, P(c_i64), P(ctypes.c_uint32), P(ctypes.c_uint32)]),
    "af_host_alloc": (c_vp, [ctypes.c_size_t]),
    "af_host_free": (None, [c_vp]),
    "af_genome_from_fasta": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, P(c_vp)]),
    "af_genome_from_contigs": (ctypes.c_int, [P(ctypes.c_char_p), P(ctypes.c_char_p), P(c_i64), c_i32, ctypes.c_int, P(c_vp)]),
    "af_debug_genome_fasta": (ctypes.c_int, [ctypes.c_char_p, P(c_i64), P(c_i32), P(ctypes.c_uint64)]),
    "af_genome_synth": (ctypes.c_int, [ctypes.c_uint64, c_i64, ctypes.c_int, P(c_vp)]),
    "af_genome_free": (None, [c_vp]),
    "af_genome_length": (c_i64, [c_vp]),
    "af_genome_n_contigs": (c_i32, [c_vp]),
    "af_genome_contig": (ctypes.c_int, [c_vp, c_i32, P(ctypes.c_char_p), P(c_i64), P(c_i64)]),
    "af_genome_align": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, P(Params), c_i32, c_vp, P(c_i64), P(GenomeStats)]),
    "af_synth_anchor": (ctypes.c_int, [P(Synth), c_vp]),
    "af_synth_pairs_host": (ctypes.c_int, [P(Synth), c_i64, c_i64, c_vp, c_vp]),
    "af_synth_pairs_device": (ctypes.c_int, [P(Synth), c_i64, c_i64, c_i32, c_vp, c_vp]),
}


class AnchoredFusionError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libafb200.so; raise if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AnchoredFusionError(
                "libafb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python anchored_fusion_b200/build.py` (needs nvcc); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.af_abi_version() != 2:
            raise AnchoredFusionError("libafb200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise AnchoredFusionError("libafb200 error %d: %s" % (rc, lib().af_last_error().decode(errors="replace")))


def ptr(a):
    """data pointer of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()
