#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n > gpurun_out/u_n$n.json 2> gpurun_out/u_n$n.err; echo "n$n rc=$?"; tail -n 1 gpurun_out/u_n$n.json | cut -c1-330
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --env ColourMatch-v0 --envs 1048576 --e2e-steps 5 > gpurun_out/u_n8_cm_1m.json 2> gpurun_out/u_n8b.err; echo "n8 cm rc=$?"; tail -n 1 gpurun_out/u_n8_cm_1m.json | cut -c1-330
