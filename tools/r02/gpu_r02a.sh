#!/bin/bash
# round 2, first GPU call: tests, the driver's bench command, default bench, and a sweep of launch strategies
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench_driver.json 2> gpurun_out/r02a_bench_driver.err; echo "bench driver-args rc=$?"
tail -3 gpurun_out/r02a_bench_driver.err
timeout 600 python tools/sweep.py PointTSP-v0:65536 PointTSP-v0:65536:c0 PointTSP-v0:65536:c0:s2 PointTSP-v0:65536:c0:s3 PointTSP-v0:65536:c1:s2 \
   PointTTSP-v0:262144 PointTTSP-v0:262144:c1 PointTTSP-v0:262144:c0:s2 PointTTSP-v0:262144:c1:s2 PointTTSP-v0:65536 PointTTSP-v0:65536:c0:s2 PointTTSP-v0:65536:c0:s3 \
   ColourMatch-v0:262144 ColourMatch-v0:262144:c1 ColourMatch-v0:262144:c0:s2 ColourMatch-v0:262144:c0:s3 \
   PointTSP-v0:1048576 PointTTSP-v0:1048576 ColourMatch-v0:1048576 > gpurun_out/r02a_sweep.jsonl 2> gpurun_out/r02a_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02a_sweep.jsonl | cut -c1-330
tail -5 gpurun_out/r02a_sweep.err
timeout 300 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:262144:c0:s2 --cfg beta_a=400 --cfg beta_b=0.5 > gpurun_out/r02a_sweep_noreset.jsonl 2>> gpurun_out/r02a_sweep.err; echo "sweep2 rc=$?"
cat gpurun_out/r02a_sweep_noreset.jsonl | cut -c1-330
