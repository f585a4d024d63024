#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --env PointTTSP-v0 --envs 262144 --steps 600 --warmup 1200 --no-cpu-baseline --e2e-steps 2"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_kernel|prefetch|reset_kernel|gather" -s 1300 -c 500 --csv --log-file gpurun_out/r01ab_launches_ttsp_262144.csv $CMD > gpurun_out/ab_ncu.log 2>&1; echo "ncu rc=$?"
