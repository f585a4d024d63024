#!/usr/bin/env python
"""Tuning variants of the CUDA library (other -D settings of the same source), each its own in-tree .so that
CRL_B200_LIB selects at run time; the product is always the default build.
    python tools/build_variants.py w24=CRL_WARPS_N15=24,CRL_STAGE_ROWS=16 w20=CRL_WARPS_N15=20,CRL_STAGE_ROWS=16"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from combinatorial_rl_tasks_b200 import build as b
for spec in sys.argv[1:]:
    name, defs = spec.split('=', 1)
    d = dict(x.split('=') for x in defs.split(','))
    print(b.build(force=True, defines=d, out=os.path.join(b.HERE, f'libcrl_b200_{name}.so')))
