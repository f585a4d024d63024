#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/av_n2.json 2> gpurun_out/av_n2.err; echo "n2 rc=$?"; wc -l gpurun_out/av_n2.json; cut -c1-260 gpurun_out/av_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/av_n2_ref.json 2> gpurun_out/av_n2b.err; echo "n2 ref rc=$?"; wc -l gpurun_out/av_n2_ref.json
