#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/probe_step_overhead.py 2>&1 | tail -6
timeout 600 python tools/bench_rollout.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['env'], d['envs'], 'collect %.3e env-steps/s' % d['collect_env_steps_per_s'], 'gae %.1f us frac %.2f' % (d['gae_us'], d['gae_frac_of_hbm_peak']))"
