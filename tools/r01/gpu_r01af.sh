#!/bin/bash
set -u
mkdir -p gpurun_out
CMD_T="python bench.py --env PointTTSP-v0 --envs 262144 --steps 200 --warmup 1000 --no-cpu-baseline --e2e-steps 2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1100 -c 2 -f -o gpurun_out/r01af_step_ttsp_262144 $CMD_T > gpurun_out/af_ncu.log 2>&1
echo "ncu ttsp rc=$?"
