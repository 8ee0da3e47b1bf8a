"""The drop-in boundary is a plain C ABI: examples/anchor_host_batch.c is compiled as strict C99 against
include/anchored_fusion.h and linked to libafb200.so.  Without a GPU it must stop loudly at
af_index_upload (no CPU fallback); on a B200 it prints the two anchored records of its toy pair."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path):
    exe = str(tmp_path / "anchor_host_batch")
    libdir = os.path.join(ROOT, "anchored_fusion_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "anchor_host_batch.c"), "-o", exe, "-L" + libdir, "-lafb200", "-Wl,-rpath," + libdir]
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.check_call(cmd, env=env)
    return exe


def test_c_client_compiles_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "anchor 1279 bp, k'=12" in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3 and "af_index_upload failed (-2)" in r.stderr


@pytest.mark.gpu
def test_c_client_anchors_its_toy_pair(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("read ")]
    assert lines == ["read 0 (pair 0 mate 1): POS 201  0S100M0S  strand 0 score 100",
                     "read 1 (pair 0 mate 2): POS 401  0S100M0S  strand 1 score 100"]
