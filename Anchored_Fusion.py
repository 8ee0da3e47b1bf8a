#!/usr/bin/env python
"""Bulk driver, same flags as the reference's Anchored_Fusion.py; runs the anchoring stage on B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from anchored_fusion_b200.cli import main_bulk  # noqa: E402

if __name__ == '__main__':
    sys.exit(main_bulk())
