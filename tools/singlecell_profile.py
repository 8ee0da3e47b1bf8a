"""cProfile of the single-cell driver over the cells tools/singlecell_e2e.py left behind."""
import cProfile
import os
import pstats
import shutil
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anchored_fusion_b200.cli import main_singlecell  # noqa: E402

root = os.path.join(tempfile.gettempdir(), "af_sc_e2e")
shutil.rmtree(os.path.join(root, "out"), ignore_errors=True)
pr = cProfile.Profile()
pr.enable()
main_singlecell(["--file_anchored_cds", os.path.join(root, "genes.fa"), "--fastq_dir", os.path.join(root, "cells"),
                 "--out_folder", os.path.join(root, "out")])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
