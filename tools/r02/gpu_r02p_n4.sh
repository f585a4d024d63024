#!/bin/bash
# round 2, call p (4 GPUs): the bench as the driver launches it for N = 4; smoke() on one of them
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err; echo "bench n2 rc=$?"
tail -n 4 gpurun_out/r02p_bench_n4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02p_bench_n4.json').read().strip().splitlines()[-1])
print('n_gpus %d value %.4g frac %.4f e2e %.4g' % (d['n_gpus'], d['value'], d['roofline']['frac'], d['e2e']['value']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g' % (v['frac'], v['value']), v['episode_stats']['reduction'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02p_ref_n4.json 2> gpurun_out/r02p_ref_n4.err; echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/r02p_ref_n4.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/r02p_smoke.log
