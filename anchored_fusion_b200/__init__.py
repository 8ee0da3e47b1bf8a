"""anchored_fusion_b200 -- B200-native read-anchoring path of Anchored-Fusion.

Drop-in for the reference's `bwa index` / `bwa mem -M | samtools view` stage
(Anchored_Fusion.py:167-194).  All compute happens in libafb200.so (hand-written sm_100a
CUDA behind the C ABI of include/anchored_fusion.h); importing this package fails loudly
if that library is not built -- there is no CPU fallback.
"""
from ._lib import HIT_DTYPE, AnchoredFusionError, lib  # noqa: F401

lib()  # load now: a missing CUDA library must not go unnoticed

from .anchoring import (Anchorer, AnchorIndex, PackedBatch, default_params, layout, pack_pairs,  # noqa: E402,F401
                        synth_anchor, synth_pairs_device, synth_pairs_host, synth_spec, unpack_read, wire_bytes,
                        wire_from_packed, wire_to_packed)
