#!/usr/bin/env python
"""A few dozen ZoneVecEnv.step_host calls (PointTSP 65,536, caller's actions in page-locked arrays) and nothing else: the target
of an ncu capture of the zero-copy host step kernel (PCIe bytes and duration per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import combinatorial_rl_tasks_b200 as crl

B = 65536
env = crl.ZoneVecEnv('PointTSP-v0', B); env.seed(1); env.reset()
acts = env.pinned_actions(4)
rs = np.random.RandomState(0)
for a in acts:
    np.copyto(a, rs.uniform(-1, 1, a.shape).astype(np.float32))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    env.step_host(acts[i % 4])
torch.cuda.synchronize()
print('rows moved', env.host_rows_moved())
