#!/usr/bin/env python
"""Registers / stack / spills / shared memory of every kernel in the CUDA build (ptxas -v), one line per kernel.
Usage: python tools/ptxas_report.py [-DNAME=VAL ...]   (compiles to /tmp, never touches the product .so)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from combinatorial_rl_tasks_b200 import build as b

def main():
    defs = [a for a in sys.argv[1:] if a.startswith('-D')]
    only = [a for a in sys.argv[1:] if not a.startswith('-D')]
    srcs = [b.SRC, b.SRC_ENCODE] if not only else [s for s in (b.SRC, b.SRC_ENCODE) if any(o in s for o in only)]
    cmd = ['/usr/local/cuda/bin/nvcc'] + b.NVCC_FLAGS + defs + ['-Xptxas', '-v', '-o', '/tmp/ptxas_report.so'] + srcs
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stderr
    name = None
    rows = {}
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(.*', '', name).replace('crl::', '').replace('void ', '')
            rows[name] = {}
            continue
        if name is None:
            continue
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m and 'stack' not in rows[name]:       # the entry's own line comes first; callees' (warp_reset) follow
            rows[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r'Used (\d+) registers', line)
        if m:
            rows[name]['regs'] = int(m.group(1))
            sm = re.search(r'(\d+) bytes smem', line)
            rows[name]['smem'] = int(sm.group(1)) if sm else 0
    for k, v in rows.items():
        print(f"{v.get('regs', '?'):>4} regs {v.get('stack', 0):>4} stack {v.get('spill_st', 0):>4}/{v.get('spill_ld', 0):<4} spill {v.get('smem', 0):>6} smem  {k}")

if __name__ == '__main__':
    main()
