"""Diagnostics: cost of the in-kernel auto-reset (latency of a single env reset inside a
step) for each task.  Not a bench line."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402
from combinatorial_rl_tasks_b200 import _lib  # noqa: E402


def timed(fn, n):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    for _ in range(n):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) * 1e3 / n


for env_id in ('PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0'):
    B = 262144
    env = crl.ZoneVecEnv(env_id, B)
    env.seed(5); env.reset()
    a = torch.rand(B, 2, device='cuda') * 2 - 1
    for _ in range(5):
        env._step(a, 0)
    base = timed(lambda: env._step(a, 0), 50)
    print(f'{env_id}: step without auto-reset {base:8.1f} us')
    for every in (0, 4096, 512, 64, 32, 1):
        mask = torch.zeros(B, dtype=torch.uint8, device='cuda')
        if every:
            mask[::every] = 1
        n = int(mask.sum())
        env.reset(mask=mask)
        t = timed(lambda: env.reset(mask=mask), 5)
        print(f'   reset kernel, {n:7d} envs reset (1 per {every:5d}): {t:9.1f} us')
    # steps with a controlled fraction of envs finishing: set steps so that 1/every envs end
    for every in (4096, 512, 64):
        def prep():
            bits = env.aux[:, 3].view(torch.int32)
            bits.copy_(bits & ~0xffff)
            bits[::every] |= (env.spec.num_steps - 1)
        prep(); torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); env._step(a, _lib.STEP_AUTO_RESET); t1.record(); torch.cuda.synchronize()
        print(f'   step with auto-reset, {B // every:6d} envs finishing: {t0.elapsed_time(t1) * 1e3:9.1f} us')
