#!/usr/bin/env python
"""Launcher of the whole-matrix from-pixels parity check of BASELINE config 2 (implementation and oracle use live in
tests/check_config2_full.py - test infrastructure). Usage on a GPU box:
    python tools/check_config2_full.py            # 1063 frames, writes profiles/r2_config2_full_parity.json"""
import os
import runpy
import sys

sys.argv[0] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "check_config2_full.py")
runpy.run_path(sys.argv[0], run_name="__main__")
