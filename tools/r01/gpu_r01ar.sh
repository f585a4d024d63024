#!/bin/bash
# the reference's training configuration: 100 fixed maps in a layout bank (resets are copies, no sampler)
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d episodes %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s['episodes']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTTSP-v0:262144 PointTTSP-v0:1048576 PointTTSP-v0:65536 PointTSP-v0:65536 ColourMatch-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --bank 100 --no-cpu-baseline --e2e-steps 2 --steps 16000 --warmup 1600 > gpurun_out/ar_${env}_${n}_bank.json 2>>gpurun_out/ar_err.log; show gpurun_out/ar_${env}_${n}_bank.json
done
tail -3 gpurun_out/ar_err.log
