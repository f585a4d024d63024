#!/bin/bash
# Re-entry check of HEAD: GPU tests, smoke, the three single-GPU configs, reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/i_smoke.log
timeout 400 python bench.py > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/i_bench_default.json
for spec in PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:1048576 PointTTSP-v0:1048576 ColourMatch-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 5 --repeats 2 > gpurun_out/i_bench_${env}_${n}.json 2>> gpurun_out/i_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/i_bench_${env}_${n}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-30s %.3e  frac %.3f  %.2f us/step  pf %d inl %d to %s e2e %.3e"%("${env}:${n}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts'), d['e2e']['value']))
except Exception as e:
    print("${env}:${n} FAILED", e)
PY
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_reference.json 2>&1; cut -c1-600 gpurun_out/i_bench_reference.json
