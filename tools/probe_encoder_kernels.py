"""Per-kernel device times of ZoneEncoder's forwards (torch profiler / CUPTI), 65,536 PointTSP envs, h = 185: which of
the two launches of crl_encoder_forward the time goes to."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402

B, h, N, Z = 65536, 185, 15, 6
g = torch.Generator(device='cuda').manual_seed(0)
rn = lambda *s, scale=1.0: torch.randn(*s, device='cuda', generator=g) * scale
sd = {'zone_net_.0.weight': rn(h, 8 + Z, scale=0.3), 'zone_net_.0.bias': rn(h, scale=0.1), 'zone_net_.2.weight': rn(h, h, scale=0.1),
      'zone_net_.2.bias': rn(h, scale=0.1), 'zone_net_.4.weight': rn(h, h, scale=0.1), 'zone_net_.4.bias': rn(h, scale=0.1),
      'combine_net_.weight': rn(h, 8 + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
enc = crl.ZoneEncoder(sd, num_zones=N)
env = crl.ZoneVecEnv('PointTSP-v0', B)
env.seed(1)
obs = env.reset()
reps = [(obs['obs'].clone(), obs['zone_obs'].clone()) for _ in range(11)]
for i in range(5):
    enc(*reps[i]); enc.forward_from_state(env); enc._head(enc.packed_head, reps[i][0], enc.pooled(*reps[i]))
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(22):
        enc(*reps[i % 11])
    for i in range(22):
        enc.forward_from_state(env)
    for i in range(22):
        enc._head(enc.packed_head, reps[i % 11][0], enc.pooled(*reps[i % 11]))
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f'{e.key[:90]:92s} n={e.count:4d} avg {e.device_time_total / e.count:8.1f} us')
