#!/bin/bash
# delta host step (e2e), TimedTSP diagnosis
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j_pytest.log
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=d['episode_stats']; e=d['e2e']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d | e2e %.3e (full %.3e) d2h %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], e['value'], e.get('full_copy_value',0), e['d2h_bytes_per_step']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/j_tsp.json 2>gpurun_out/j_tsp.err; show gpurun_out/j_tsp.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline > gpurun_out/j_cm.json 2>>gpurun_out/j_err.log; show gpurun_out/j_cm.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 20 > gpurun_out/j_ttsp.json 2>>gpurun_out/j_err.log; show gpurun_out/j_ttsp.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 5 --no-auto-reset --prefetch-every 0 > gpurun_out/j_ttsp_noreset.json 2>>gpurun_out/j_err.log; show gpurun_out/j_ttsp_noreset.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 5 --chained 1 > gpurun_out/j_ttsp_chained.json 2>>gpurun_out/j_err.log; show gpurun_out/j_ttsp_chained.json
timeout 300 python bench.py --env PointTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 5 > gpurun_out/j_tsp_262144.json 2>>gpurun_out/j_err.log; show gpurun_out/j_tsp_262144.json
CMD_T="python bench.py --env PointTTSP-v0 --envs 262144 --steps 1400 --warmup 100 --no-cpu-baseline --e2e-steps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1300 -c 2 -f -o gpurun_out/r01j_step_ttsp_262144 $CMD_T > gpurun_out/j_ncu.log 2>&1
echo "ncu ttsp rc=$?"
