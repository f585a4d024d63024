#!/bin/bash
# round 2, call av (4 GPUs): the bench as the driver launches it for N = 4 at the HEAD of round 2
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02av_bench_n4.json 2> gpurun_out/r02av_bench_n4.err; echo "bench n4 rc=$?"
tail -n 4 gpurun_out/r02av_bench_n4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02av_bench_n4.json').read().strip().splitlines()[-1])
print('n_gpus %d value %.4g frac %.4f e2e %.4g' % (d['n_gpus'], d['value'], d['roofline']['frac'], d['e2e']['value']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g' % (v['frac'], v['value']), v['episode_stats']['reduction'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02av_ref_n4.json 2> gpurun_out/r02av_ref_n4.err; echo "ref n4 rc=$?"; cut -c1-300 gpurun_out/r02av_ref_n4.json
