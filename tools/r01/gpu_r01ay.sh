#!/bin/bash
# ABI v5 with the fixed-placement support as a template parameter: GPU suite + headline configs
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/ay_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ay_pytest.log
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline > gpurun_out/ay_bench_ttsp.json 2>> gpurun_out/ay_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ay_bench_ttsp.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/ay_bench_tsp.json 2> gpurun_out/ay_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ay_bench_tsp.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline > gpurun_out/ay_bench_cm.json 2>> gpurun_out/ay_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ay_bench_cm.json
timeout 300 python bench.py --env PointTSP-v4 --envs 65536 --no-cpu-baseline > gpurun_out/ay_bench_v4.json 2>> gpurun_out/ay_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ay_bench_v4.json
timeout 300 python bench.py --env PointTSP-v5 --envs 65536 --no-cpu-baseline > gpurun_out/ay_bench_v5.json 2>> gpurun_out/ay_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ay_bench_v5.json
