"""Small all-paths case for compute-sanitizer (memcheck / racecheck): every kernel, ragged
batch, fused auto-reset, host-supplied layouts, qpos/qvel import/export, host-buffer step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402

for env_id in ('PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0', 'PointTSP-v1'):
    B = 301
    env = crl.ZoneVecEnv(env_id, B)
    env.seed(11)
    env.reset()
    bits = env.aux[:, 3].view(torch.int32)
    bits.copy_((bits & ~0xffff) | (env.spec.num_steps - 6))
    for t in range(12):                      # every env finishes and is rebuilt in-kernel
        env.step_random(action_seed=5)
    env.step(torch.zeros(B, 2, device='cuda'))
    env.step_no_reset(torch.ones(B, 2, device='cuda'))
    env.step_host(np.zeros((B, 2), np.float32))
    env.reset(mask=torch.arange(B) % 3 == 0)
    N = env.spec.num_zones
    rs = np.random.RandomState(0)
    lay = {'xy0': rs.uniform(-2, 2, (5, 2)), 'rot0': rs.uniform(0, 6, 5), 'zone_xy': rs.uniform(-2, 2, (5, N, 2)),
           'zone_max_steps': rs.randint(100, 900, (5, N)), 'colours': rs.randint(0, 3, (5, N))}
    env.reset(layout=lay, env_ids=[0, 7, 33, 299, 300])
    env.set_qpos_qvel(rs.uniform(-1, 1, (5, 3)), rs.uniform(-1, 1, (5, 3)), env_ids=[0, 7, 33, 299, 300])
    qp, qv = env.get_qpos_qvel()
    env.physics_substeps(torch.zeros(B, 2, device='cuda'), 3)
    torch.cuda.synchronize()
    print(env_id, 'ok', env.counters())
