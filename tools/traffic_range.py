#!/usr/bin/env python
"""DRAM traffic of a WHOLE run of steps, launches overlapping as in the bench (per-kernel ncu counters serialise the launches
and miss what the write-back L2 drains between two kernels): bench.py's ring stepped for --steps steps between
cudaProfilerStart / Stop, to be run under

    ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none \
        --clock-control none --csv --log-file out.csv python tools/traffic_range.py PointTSP-v0:65536 --steps 700

Prints the steps taken and their algorithmic bytes; traffic / algorithmic = the range's read + write sum over that."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('case')
    ap.add_argument('--steps', type=int, default=700)
    a = ap.parse_args()
    import torch
    import combinatorial_rl_tasks_b200 as crl
    from combinatorial_rl_tasks_b200 import _lib
    env_id, B = a.case.split(':')[0], int(a.case.split(':')[1])
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    args = argparse.Namespace(bank=0, prefetch_every=32, prefetch_warps=0, cfg=[], no_auto_reset=False, env=env_id, min_replicas=0)
    ring = bench.Ring(crl, _lib, args, env_id, B, dev, 0, chained=None, streams=0)
    steps = (a.steps // ring.R) * ring.R                      # whole ring cycles
    ring.run(ring.prepare(4 * ring.R * 8))                    # warm: graphs built, caches in their steady state
    plan = ring.prepare(steps)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ring.run(plan)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({'case': a.case, 'steps': steps, 'ring_replicas': ring.R, 'streams': ring.S,
                      'bytes_per_env_step': ring.step_bytes, 'algorithmic_bytes': steps * B * ring.step_bytes}), flush=True)


if __name__ == '__main__':
    main()
