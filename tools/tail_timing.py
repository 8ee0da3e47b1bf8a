"""Per-warp cycle breakdown of k_tail's first region (debug hook af_debug_tail_timing)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anchored_fusion_b200 as af
from anchored_fusion_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
spec = af.synth_spec(seed=1, ref_len=10_000_000, anchor_start=4_000_000, anchor_len=6783, read_len=150, sub_ppm=10_000, fusion_ppm=0)
index = af.AnchorIndex(af.synth_anchor(spec))
eng = af.Anchorer(index, 0)
batch = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
for _ in range(3):
    eng.anchor(batch, cand_cap=n // 4)
buf = torch.zeros((148 * 8 + 148 * 64 * 8,), dtype=torch.int64, device="cuda")
L = lib()
L.af_debug_tail_timing.argtypes = [ctypes.c_void_p]
L.af_debug_tail_timing(buf.data_ptr())
eng.anchor(batch, cand_cap=n // 4)
torch.cuda.synchronize()
L.af_debug_tail_timing(None)
full = buf.cpu().numpy()
b = full[:148 * 8].reshape(148, 8)
names = ["staged", "phase1a_end", "phase1b_end", "lookback_done", "cta_end", "queued"]
for i, nm in enumerate(names):
    v = b[:, i]
    print("%-14s min %9d  median %9d  max %9d" % (nm, v.min(), int(np.median(v)), v.max()))

t = full[148 * 8:].reshape(148, 64, 8)
t = t[t[:, :, 0] > 0]
if len(t):
    print("thread_extend (AF_TAIL_PROF build), %d threads sampled:" % len(t))
    for i, nm in enumerate(["total", "collect", "eval", "n_eval", "n_load", "n_iter", "t_start", "t_consume"]):
        v = t[:, i]
        print("  %-12s min %8d  median %8d  mean %10.1f  max %8d" % (nm, v.min(), int(np.median(v)), v.mean(), v.max()))
