#!/bin/bash
# round 2, call ay (8 GPUs): the bench exactly as the driver launches it for N = 8 (SCALE's last point)
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02ay_bench_n8.json 2> gpurun_out/r02ay_bench_n8.err; echo "bench n8 rc=$? wall $(( $(date +%s) - t0 )) s"
tail -n 4 gpurun_out/r02ay_bench_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ay_bench_n8.json').read().strip().splitlines()[-1])
print('n_gpus %d value %.4g frac %.4f e2e %.4g' % (d['n_gpus'], d['value'], d['roofline']['frac'], d['e2e']['value']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g' % (v['frac'], v['value']), v['episode_stats']['reduction'])
PY
nvidia-smi topo -m > gpurun_out/r02ay_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/r02ay_lscpu.txt
