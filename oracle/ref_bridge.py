"""Import the reference's functions.py in THIS container -- TEST INFRASTRUCTURE ONLY.

/root/reference/functions.py:4 does `from Bio import Align`, but the two PairwiseAligner
objects it builds (functions.py:772-776, :1149-1153) are never used (.align is never
called), so a stub Bio.Align makes the module importable.  Used only by
tests/golden/make_goldens.py to produce committed golden vectors; /root/reference does not
exist on the GPU box, so nothing at test/bench run time may call this.
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def load_reference_functions():
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        align = types.ModuleType("Bio.Align")

        class PairwiseAligner:  # constructed by the reference, never called
            pass

        align.PairwiseAligner = PairwiseAligner
        bio.Align = align
        sys.modules["Bio"] = bio
        sys.modules["Bio.Align"] = align
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("functions")
