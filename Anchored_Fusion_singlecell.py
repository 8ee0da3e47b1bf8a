#!/usr/bin/env python
"""Per-cell driver, same flags as the reference's Anchored_Fusion_singlecell.py; runs the anchoring
stage on B200 for every <cell>_1/<cell>_2 FASTQ pair of --fastq_dir."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from anchored_fusion_b200.cli import main_singlecell  # noqa: E402

if __name__ == '__main__':
    sys.exit(main_singlecell())
