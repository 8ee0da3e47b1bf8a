"""The host side of the library (index build, packer, FASTQ reader) under AddressSanitizer + UBSan on
random and deliberately damaged inputs (tools/fuzz_host.cpp).  compute-sanitizer is not available on
the GPU pool; this sanitises the code that parses files it did not write."""
import os
import subprocess

import pytest

from conftest import ROOT


def test_host_code_is_clean_under_asan_and_ubsan(tmp_path):
    exe = str(tmp_path / "fuzz_host")
    csrc = os.path.join(ROOT, "anchored_fusion_b200", "csrc")
    cmd = ["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "fuzz_host.cpp"),
           os.path.join(csrc, "af_host.cpp"), os.path.join(csrc, "af_fastq.cpp"), "-o", exe, "-lz", "-lpthread"]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    b = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert b.returncode == 0, b.stderr[-2000:]
    work = tmp_path / "w"
    work.mkdir()
    r = subprocess.run([exe, str(work), "30", "7"], capture_output=True, text=True, timeout=600,
                       env=dict(env, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert "fuzz_host: 30 iterations" in r.stdout
