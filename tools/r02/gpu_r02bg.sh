#!/bin/bash
# round 2, call bg: ncu full set of the HEAD step kernels (TimedTSP 262,144, ColourMatch 262,144, PointTSP 65,536), each after its plain run
set -u
mkdir -p gpurun_out
for c in PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:65536; do
  tag=$(echo $c | tr ':' '_' | tr -d '-')
  CMD="python tools/sweep.py $c --seconds 0.3"
  timeout 300 $CMD > gpurun_out/r02bg_plain_$tag.jsonl 2>> gpurun_out/r02bg_err.log &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1500 -c 2 -f -o gpurun_out/r02bg_step_$tag $CMD > gpurun_out/r02bg_ncu_$tag.log 2>&1; echo "ncu $c rc=$?"
  cut -c1-140 gpurun_out/r02bg_plain_$tag.jsonl
done
tail -n 3 gpurun_out/r02bg_err.log; ls -la gpurun_out/r02bg_*.ncu-rep
