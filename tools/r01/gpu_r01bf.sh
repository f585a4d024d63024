#!/bin/bash
# bench lines at the HEAD of round 1 for BASELINE configs[4]'s per-GPU size (1,048,576 envs) and the encoder sweep
set -u
mkdir -p gpurun_out
for env in PointTSP-v0 PointTTSP-v0 ColourMatch-v0; do
  timeout 300 python bench.py --env $env --envs 1048576 --steps 8000 --warmup 800 --no-cpu-baseline > gpurun_out/bf_${env}_1m.json 2>> gpurun_out/bf_err.log; echo "$env rc=$?"; cut -c1-100 gpurun_out/bf_${env}_1m.json
done
timeout 200 python tools/bench_encode.py --envs 262144 --env ColourMatch-v0 --hidden 185 > gpurun_out/bf_enc_cm.json 2>> gpurun_out/bf_err.log; echo "enc cm rc=$?"; cut -c1-330 gpurun_out/bf_enc_cm.json
timeout 200 python tools/bench_encode.py --envs 262144 --env PointTTSP-v0 --hidden 185 > gpurun_out/bf_enc_ttsp.json 2>> gpurun_out/bf_err.log; echo "enc ttsp rc=$?"; cut -c1-330 gpurun_out/bf_enc_ttsp.json
timeout 200 python tools/bench_encode.py --envs 65536 --hidden 64 > gpurun_out/bf_enc_h64.json 2>> gpurun_out/bf_err.log; echo "enc h64 rc=$?"; cut -c1-330 gpurun_out/bf_enc_h64.json
