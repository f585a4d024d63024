#!/bin/bash
# round 2, call b: where do TimedTSP's resets go?  full ncu set with source on step_kernel<TTSP,15>, and the
# DRAM read+write bytes of 70 consecutive PointTSP launches with the caches left alone (single pass, no replay)
set -u
mkdir -p gpurun_out

CMD="python tools/sweep.py PointTTSP-v0:262144 --seconds 0.3"
timeout 300 $CMD > gpurun_out/r02b_plain_ttsp.jsonl 2> gpurun_out/r02b_err.log &&
timeout 900 ncu --graph-profiling node --set full --clock-control none --import-source on -k regex:step_kernel -s 1500 -c 2 -f -o gpurun_out/r02b_step_ttsp_262144 $CMD > gpurun_out/r02b_ncu1.log 2>&1; echo "ncu ttsp rc=$?"
cat gpurun_out/r02b_plain_ttsp.jsonl | cut -c1-250
CMD2="python tools/sweep.py PointTSP-v0:65536 --seconds 0.2"
timeout 300 $CMD2 > gpurun_out/r02b_plain_tsp.jsonl 2>> gpurun_out/r02b_err.log &&
timeout 900 ncu --graph-profiling node --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --cache-control none --clock-control none -k regex:step_kernel -s 1000 -c 70 --csv --log-file gpurun_out/r02b_traffic_tsp_65536.csv $CMD2 > gpurun_out/r02b_ncu2.log 2>&1; echo "ncu traffic rc=$?"
tail -n 3 gpurun_out/r02b_ncu1.log; tail -n 3 gpurun_out/r02b_ncu2.log
