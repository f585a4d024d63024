"""How fast can ANY one-launch-per-step kernel move the step's bytes at this batch size?
Graph-replayed torch device copies with the same bytes per launch as one env-step batch
(read+write = bytes), cycling a ring of buffers larger than 2x L2, exactly like bench.py.
Reported next to the step kernel's number in profiles/ as the practical ceiling of the
launch granularity (it is NOT the roofline denominator; that stays MEASURED_PEAKS.json)."""
import sys
import torch

def probe(total_bytes, ring_bytes=2 * 126 * 1024 * 1024 + 1, steps=4000):
    half = total_bytes // 2 // 16 * 16
    R = max(2, -(-ring_bytes // total_bytes))
    src = [torch.empty(half, dtype=torch.uint8, device='cuda').random_() for _ in range(R)]
    dst = [torch.empty(half, dtype=torch.uint8, device='cuda') for _ in range(R)]
    def cycle():
        for a, b in zip(src, dst):
            b.copy_(a)
    cycle(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cycle()
    for _ in range(20):
        g.replay()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    n = max(1, steps // R)
    for _ in range(n):
        g.replay()
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1e3 / (n * R)
    return us, 2 * half / us / 1e3

if __name__ == '__main__':
    for envs in (65536, 262144, 1048576):
        for per in (592,):
            us, gbs = probe(envs * per)
            print(f'copy of {envs * per / 1e6:7.1f} MB traffic per launch ({envs} envs x {per} B): {us:8.2f} us/launch, {gbs:7.0f} GB/s')
