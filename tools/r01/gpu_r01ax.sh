#!/bin/bash
# ABI v5 (hard instances PointTSP-v4 / v5): whole GPU suite, then the three headline configs
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/ax_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/ax_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/ax_bench_tsp.json 2> gpurun_out/ax_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ax_bench_tsp.json
timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline > gpurun_out/ax_bench_ttsp.json 2>> gpurun_out/ax_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ax_bench_ttsp.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline > gpurun_out/ax_bench_cm.json 2>> gpurun_out/ax_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ax_bench_cm.json
timeout 300 python bench.py --env PointTSP-v4 --envs 65536 --no-cpu-baseline > gpurun_out/ax_bench_v4.json 2>> gpurun_out/ax_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/ax_bench_v4.json
