#!/bin/bash
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d launches %s" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], d['gpu_launches_detail']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 16000 --warmup 1600 > gpurun_out/at_cm_16k.json 2>>gpurun_out/at_err.log; show gpurun_out/at_cm_16k.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 64000 --warmup 6400 --prefetch-warps 4 > gpurun_out/at_cm_64k_w4.json 2>>gpurun_out/at_err.log; show gpurun_out/at_cm_64k_w4.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 64000 --warmup 6400 --prefetch-every 128 > gpurun_out/at_cm_64k_pe128.json 2>>gpurun_out/at_err.log; show gpurun_out/at_cm_64k_pe128.json
