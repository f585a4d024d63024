#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c_pytest.log
python tools/probe_prefetch.py PointTTSP-v0 262144 2>&1 | tee gpurun_out/probe_prefetch.log
python tools/probe_prefetch.py PointTSP-v0 262144 2>&1 | tee -a gpurun_out/probe_prefetch.log
for i in 1 2 3; do bash -c 'python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --repeats 3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"TTSP\", d[\"roofline\"][\"frac\"], d[\"all_reps_ms\"])"'; done
for spec in PointTSP-v0:65536 ColourMatch-v0:262144; do env=${spec%%:*}; n=${spec##*:}; python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', d['roofline']['frac'], d['all_reps_ms'])"; done
