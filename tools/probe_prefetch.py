"""Diagnostics (not a bench line): what do in-step resets and the background prefetch cost?
For one task/batch: (a) 16-step windows with every next-layout slot pre-filled and NO prefetch
running, (b) the prefetch kernel alone after such a window, (c) windows with the prefetch
launched concurrently on the side stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402
from combinatorial_rl_tasks_b200 import _lib  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else 'PointTTSP-v0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
R = 2
envs = []
for r in range(R):
    e = crl.ZoneVecEnv(env_id, B, prefetch_every=0, env_offset=r * B)
    e.seed(1 + r * B); e.reset(); envs.append(e)
acts = [torch.rand(B, 2, device='cuda') * 2 - 1 for _ in range(R)]


def cycle():
    for e, a in zip(envs, acts):
        e._step(a, _lib.STEP_AUTO_RESET)


cycle(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    cycle()
side = torch.cuda.Stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


def prefetch_all(stream=None):
    for e in envs:
        e.prefetch(stream)


# warm-up into steady state (resets happening), slots kept filled
for i in range(600):
    g.replay()
    if i % 8 == 7:
        prefetch_all(torch.cuda.current_stream())
torch.cuda.synchronize()
for W in (8, 32):
    ta, tb, tc = [], [], []
    for rep in range(12):
        prefetch_all(torch.cuda.current_stream()); torch.cuda.synchronize()
        a0, a1 = ev(), ev()
        a0.record()
        for _ in range(W):
            g.replay()
        a1.record(); torch.cuda.synchronize()
        ta.append(a0.elapsed_time(a1) * 1e3 / (W * R))
        b0, b1 = ev(), ev()
        b0.record(); prefetch_all(torch.cuda.current_stream()); b1.record(); torch.cuda.synchronize()
        tb.append(b0.elapsed_time(b1) * 1e3 / R)
        # concurrent: prefetch of the previous window's resets runs on the side stream during this window
        for _ in range(W):
            g.replay()
        torch.cuda.synchronize()
        c0, c1 = ev(), ev()
        c0.record()
        side.wait_stream(torch.cuda.current_stream())
        prefetch_all(side)
        for _ in range(W):
            g.replay()
        c1.record(); torch.cuda.synchronize()
        tc.append(c0.elapsed_time(c1) * 1e3 / (W * R))
        side.synchronize()
    med = lambda v: sorted(v)[len(v) // 2]
    c = envs[0].counters()
    print(f'{env_id} B={B} window={W} cycles: step, slots pre-filled, no prefetch running: {med(ta):7.2f} us (min {min(ta):.2f}) | '
          f'prefetch kernel alone after the window: {med(tb):8.1f} us | step with prefetch concurrent: {med(tc):7.2f} us (min {min(tc):.2f}) '
          f'| inline so far {c["resets_inline"]:.0f}')
