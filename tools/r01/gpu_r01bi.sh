#!/bin/bash
# A/B: reset path with both slot flags requested together with the episode counter (-DCRL_RESET_BOTH_FLAGS)
set -u
mkdir -p gpurun_out
V=combinatorial_rl_tasks_b200/libcrl_b200_bothflags.so
CRL_B200_LIB=$V timeout 300 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_hard.py -x -q -k "prefetch or auto_reset or twin or chained or inline" > gpurun_out/bi_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/bi_pytest.log
for lib in base bothflags; do
  if [ $lib = base ]; then L=combinatorial_rl_tasks_b200/libcrl_b200.so; else L=$V; fi
  CRL_B200_LIB=$L timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bi_ttsp_$lib.json 2>> gpurun_out/bi_err.log; echo "$lib ttsp rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bi_ttsp_$lib.json')); print('$lib ttsp', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
CRL_B200_LIB=$V timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bi_cm_bothflags.json 2>> gpurun_out/bi_err.log; python -c "
import json; d=json.load(open('gpurun_out/bi_cm_bothflags.json')); print('bothflags cm', d['value'], d['ms_per_step'], d['roofline']['frac'])"
CRL_B200_LIB=$V timeout 300 python bench.py --env PointTTSP-v0 --envs 65536 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bi_ttsp64k_bothflags.json 2>> gpurun_out/bi_err.log; python -c "
import json; d=json.load(open('gpurun_out/bi_ttsp64k_bothflags.json')); print('bothflags ttsp 65536', d['value'], d['ms_per_step'], d['roofline']['frac'])"
