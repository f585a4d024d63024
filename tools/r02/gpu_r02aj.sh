#!/bin/bash
# round 2, call aj: ncu at HEAD -- launch list of the driver's bench command, full sets of the pipelined zone encoder,
# of both head kernels and of the headline step kernel (each after the same command ran plainly)
set -u
mkdir -p gpurun_out
CMD="python tools/bench_encode.py --iters 6"
timeout 300 $CMD > gpurun_out/r02aj_plain.json 2> gpurun_out/r02aj_err.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"zone_encode_kernel|encoder_head_kernel" -s 10 -c 6 -f -o gpurun_out/r02aj_encode $CMD > gpurun_out/r02aj_ncu.log 2>&1; echo "ncu encode rc=$?"
tail -n 3 gpurun_out/r02aj_ncu.log
BENCH="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.2 --extra-seconds 0.1 --e2e-seconds 0.2"
timeout 600 $BENCH > gpurun_out/r02aj_bench_short.json 2>> gpurun_out/r02aj_err.log &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02aj_launches.csv $BENCH > gpurun_out/r02aj_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel" -s 40 -c 2 -f -o gpurun_out/r02aj_step $BENCH > gpurun_out/r02aj_ncu_step.log 2>&1; echo "ncu step rc=$?"
ls -la gpurun_out/ | tail -n 12
