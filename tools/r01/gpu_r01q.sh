#!/bin/bash
# consolidated round-1 numbers at HEAD: tests, smoke, every config, reference arm
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/q_smoke.log
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=d['episode_stats']; e=d['e2e']
    print("%-40s %.3e frac %.3f %.2f us/step pf %d inl %d | e2e %.3e (full %.3e) d2h %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], e['value'], e.get('full_copy_value',0), e['d2h_bytes_per_step']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 400 python bench.py > gpurun_out/q_default.json 2>gpurun_out/q_default.err; echo "bench rc=$?"; show gpurun_out/q_default.json
for spec in PointTSP-v0:262144 PointTSP-v0:1048576 PointTTSP-v0:65536 PointTTSP-v0:262144 PointTTSP-v0:1048576 ColourMatch-v0:65536 ColourMatch-v0:262144 ColourMatch-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 20 > gpurun_out/q_${env}_${n}.json 2>>gpurun_out/q_err.log; show gpurun_out/q_${env}_${n}.json
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/q_reference.json 2>&1; cut -c1-300 gpurun_out/q_reference.json
