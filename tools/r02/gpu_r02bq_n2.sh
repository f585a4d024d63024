#!/bin/bash
# round 2, call bq (2 GPUs, result records as deltas): the driver's launch line at HEAD (prepared host call, copy-engine floor in the line)
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02bq_bench_n2.json 2> gpurun_out/r02bq_bench_n2.err; echo "bench n2 rc=$? wall $(( $(date +%s) - t0 )) s"
tail -n 3 gpurun_out/r02bq_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02bq_bench_n2.json').read().strip().splitlines()[-1])
print('n_gpus %d value %.4g frac %.4f e2e %.4g floor %.4g' % (d['n_gpus'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['copy_engine_floor_value']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g' % (v['frac'], v['value']), v['episode_stats']['reduction'])
PY
