#!/bin/bash
# 4-GPU line at the HEAD of round 1
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 4 --no-cpu-baseline --e2e-steps 50 > gpurun_out/bl_n4.json 2> gpurun_out/bl_n4.err; echo "n4 rc=$?"; wc -l gpurun_out/bl_n4.json; cut -c1-200 gpurun_out/bl_n4.json
