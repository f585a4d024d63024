#!/bin/bash
# round 2, call q (4 GPUs): does binding each rank to its GPU's NUMA node fix the end-to-end scaling? + sampler with packed math
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02q_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r02q_lscpu.txt
for mode in bind nobind; do
  extra=""; [ $mode = nobind ] && extra="--no-numa-bind"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra $extra > gpurun_out/r02q_bench_n4_$mode.json 2> gpurun_out/r02q_bench_n4_$mode.err; echo "bench n4 $mode rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02q_bench_n4_$mode.json').read().strip().splitlines()[-1])
print('$mode: n_gpus %d value %.4g frac %.4f e2e %.4g pageable %.4g' % (d['n_gpus'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['pageable_actions_value']), d['e2e']['host_placement'])
PY
done
timeout 600 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_hard.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02q_pytest.log
timeout 600 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:65536 PointTTSP-v0:1048576 --seconds 0.6 2> gpurun_out/r02q_err.log | cut -c1-48,80-200
cat gpurun_out/r02q_lscpu.txt; head -12 gpurun_out/r02q_topo.txt
