#!/bin/bash
# round 2, call bm: DRAM read + write of a whole run of overlapping launches (ncu app-range replay, caches left alone)
set -u
mkdir -p gpurun_out
for c in PointTSP-v0:65536:700 PointTTSP-v0:262144:200 ColourMatch-v0:262144:300; do
  case=$(echo $c | cut -d: -f1,2); steps=$(echo $c | cut -d: -f3); tag=$(echo $case | tr ':' '_' | tr -d '-')
  timeout 200 python tools/traffic_range.py $case --steps $steps > gpurun_out/r02bm_plain_$tag.json 2>> gpurun_out/r02bm_err.log &&
  timeout 400 ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r02bm_range_$tag.csv python tools/traffic_range.py $case --steps $steps > gpurun_out/r02bm_ncu_$tag.log 2>&1; echo "ncu $case rc=$?"
  cat gpurun_out/r02bm_plain_$tag.json; tail -n 4 gpurun_out/r02bm_range_$tag.csv | cut -c1-300
done
tail -n 3 gpurun_out/r02bm_err.log
