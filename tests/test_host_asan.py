"""The host side of the library (index build, packer, the task-parallel FASTQ reader and its DEFLATE decoder)
under AddressSanitizer + UBSan on random and deliberately damaged inputs -- plain, gzip, multi-member gzip,
BGZF, damaged before and after compression; undamaged inputs must also come back record for record
(tools/fuzz_host.cpp).  compute-sanitizer is not available on
the GPU pool; this sanitises the code that parses files it did not write."""
import os
import subprocess

import pytest

from conftest import ROOT


def test_host_code_is_clean_under_asan_and_ubsan(tmp_path):
    exe = str(tmp_path / "fuzz_host")
    csrc = os.path.join(ROOT, "anchored_fusion_b200", "csrc")
    # AF_FASTQ_TEST_SIZES: 24 KB segments / 5 KB slices / 2 KB "small file" limit, so that the multi-segment BGZF,
    # streamed-gzip and plain-text paths of the reader all run on the small fuzz files
    cmd = ["g++", "-std=c++17", "-g", "-O1", "-DAF_FASTQ_TEST_SIZES", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "fuzz_host.cpp"),
           os.path.join(csrc, "af_host.cpp"), os.path.join(csrc, "af_fastq.cpp"), os.path.join(csrc, "af_genome_host.cpp"), "-o", exe, "-lz", "-lpthread"]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    b = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert b.returncode == 0, b.stderr[-2000:]
    work = tmp_path / "w"
    work.mkdir()
    r = subprocess.run([exe, str(work), "150", "7"], capture_output=True, text=True, timeout=600,
                       env=dict(env, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert "fuzz_host: 150 iterations" in r.stdout
