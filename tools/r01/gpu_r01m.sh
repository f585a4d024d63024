#!/bin/bash
# fence dedupe in chain_release, compat.ParallelEnv test, launch list of our kernels only
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/m_pytest.log
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
timeout 400 python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/m_tsp.json 2>gpurun_out/m_tsp.err; show gpurun_out/m_tsp.json
timeout 400 python bench.py --no-cpu-baseline --chained 0 > gpurun_out/m_tsp_unchained.json 2>>gpurun_out/m_err.log; show gpurun_out/m_tsp_unchained.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 65536 --no-cpu-baseline --e2e-steps 20 > gpurun_out/m_cm_65536.json 2>>gpurun_out/m_err.log; show gpurun_out/m_cm_65536.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 65536 --no-cpu-baseline --e2e-steps 20 > gpurun_out/m_ttsp_65536.json 2>>gpurun_out/m_err.log; show gpurun_out/m_ttsp_65536.json
CMD="python bench.py --steps 210 --warmup 21 --no-cpu-baseline --e2e-steps 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_kernel|prefetch|reset_kernel|gather" -c 700 --csv --log-file gpurun_out/r01m_launches_pointtsp_65536.csv $CMD > gpurun_out/m_ncu1.log 2>&1; echo "ncu launches rc=$?"
