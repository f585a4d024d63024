#!/bin/bash
# TimedTSP with (almost) no timeouts: what does the step cost without its resets?
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-40s %.3e frac %.3f %.2f us/step pf %d inl %d episodes %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s['episodes']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 3000 --warmup 300 --cfg beta_a=400 --cfg beta_b=0.5 > gpurun_out/aa_ttsp_notimeouts.json 2>>gpurun_out/aa_err.log; show gpurun_out/aa_ttsp_notimeouts.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 3000 --warmup 300 > gpurun_out/aa_ttsp_base.json 2>>gpurun_out/aa_err.log; show gpurun_out/aa_ttsp_base.json
timeout 300 python bench.py --env PointTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 3000 --warmup 300 > gpurun_out/aa_tsp_base.json 2>>gpurun_out/aa_err.log; show gpurun_out/aa_tsp_base.json
tail -3 gpurun_out/aa_err.log
