#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r_n2.json 2> gpurun_out/r_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r_n2.err; cut -c1-400 gpurun_out/r_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --env PointTTSP-v0 --envs 1048576 --e2e-steps 5 > gpurun_out/r_n2_ttsp_1m.json 2> gpurun_out/r_n2b.err; echo "n2 ttsp rc=$?"; cut -c1-300 gpurun_out/r_n2_ttsp_1m.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r_n2_ref.json 2> gpurun_out/r_n2c.err; echo "n2 ref rc=$?"; cut -c1-200 gpurun_out/r_n2_ref.json
