#!/bin/bash
# branch-free TimedTSP row build; launch list + full ncu capture of the default config at HEAD
set -u
mkdir -p gpurun_out
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
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 20 > gpurun_out/l_ttsp.json 2>>gpurun_out/l_err.log; show gpurun_out/l_ttsp.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 1048576 --no-cpu-baseline --e2e-steps 5 > gpurun_out/l_ttsp_1m.json 2>>gpurun_out/l_err.log; show gpurun_out/l_ttsp_1m.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 20 > gpurun_out/l_cm.json 2>>gpurun_out/l_err.log; show gpurun_out/l_cm.json
CMD="python bench.py --steps 210 --warmup 21 --no-cpu-baseline --e2e-steps 3"
timeout 300 $CMD > gpurun_out/l_short.json 2>>gpurun_out/l_err.log; show gpurun_out/l_short.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01l_launches_pointtsp_65536.csv $CMD > gpurun_out/l_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -f -o gpurun_out/r01l_step_tsp_65536 $CMD > gpurun_out/l_ncu2.log 2>&1; echo "ncu full rc=$?"
CMD_C="python bench.py --env ColourMatch-v0 --envs 262144 --steps 300 --warmup 30 --no-cpu-baseline --e2e-steps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 200 -c 2 -f -o gpurun_out/r01l_step_cm_262144 $CMD_C > gpurun_out/l_ncu3.log 2>&1; echo "ncu cm rc=$?"
