"""Diagnostic: host time per ZoneVecEnv.step / Rollout.step call (the GPU queue absorbs the kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import combinatorial_rl_tasks_b200 as crl
from combinatorial_rl_tasks_b200.rollout import Rollout
B = 1024
env = crl.ZoneVecEnv('PointTSP-v0', B); env.seed(1); env.reset()
a = torch.zeros(B, 2, device='cuda')
def t(f, n=3000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize(); return dt
print('env.step(a)                 %.1f us' % t(lambda: env.step(a)))
print('env.step_random()           %.1f us' % t(lambda: env.step_random()))
ro = Rollout(env, 64); ro.begin()
v = torch.zeros(B, device='cuda')
k = [0]
def rstep():
    ro.step(k[0] % 64, a, v); k[0] += 1
print('Rollout.step(t, a, v)       %.1f us' % t(rstep))
def rstep2():
    ro.step(k[0] % 64, ro.actions[k[0] % 64]); k[0] += 1
print('Rollout.step(t, in place)   %.1f us' % t(rstep2))
