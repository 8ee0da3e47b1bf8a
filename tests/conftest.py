import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle are built in-tree (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def bundled():
    """The reference's bundled sample, decoded from tests/golden/bundled_c1.npz."""
    z = np.load(os.path.join(GOLDEN, "bundled_c1.npz"))
    L = int(z["read_len"])
    packed = z["reads_2bit"]
    codes = np.zeros((packed.shape[0], L), dtype=np.uint8)
    for i in range(L):
        codes[:, i] = (packed[:, i // 4] >> (2 * (i % 4))) & 3
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [lut[c].tobytes().decode() for c in codes]
    return {"header": str(z["header"]), "anchor": str(z["anchor"]), "names1": str(z["names1"]).split("\n"),
            "names2": str(z["names2"]).split("\n"), "codes": codes, "seqs1": seqs[0::2], "seqs2": seqs[1::2],
            "read_len": L, "qual_char": str(z["qual_char"]), "oracle_hits": z["oracle_hits"]}


def hits_equal(a, b):
    """bit-exact comparison of two hit record arrays (same dtype layout)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()
