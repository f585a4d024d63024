#!/bin/bash
# round 2, call e: counter issued first / consumed late; launch strategies; look-ahead distance; new TTSP profile
set -u
mkdir -p gpurun_out
echo "== default build"
timeout 900 python tools/sweep.py PointTSP-v0:65536 PointTSP-v0:65536:c0:s2 PointTSP-v0:65536:c0:s3 PointTSP-v0:262144 PointTSP-v0:262144:c0:s2 \
  PointTTSP-v0:262144 PointTTSP-v0:262144:c1 PointTTSP-v0:262144:c0:s2 PointTTSP-v0:65536 PointTTSP-v0:65536:c0:s2 \
  ColourMatch-v0:262144 ColourMatch-v0:262144:c1 ColourMatch-v0:262144:c0:s2 ColourMatch-v0:262144:c0:s3 \
  PointTSP-v0:1048576 PointTTSP-v0:1048576 ColourMatch-v0:1048576 --seconds 0.6 2> gpurun_out/r02e_err.log | tee gpurun_out/r02e_sweep.jsonl | cut -c1-48,80-200
for a in 0 600 2400; do
echo "== look-ahead $a CTAs"
CRL_PF_AHEAD=$a timeout 600 python tools/sweep.py PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:262144 PointTSP-v0:1048576 --seconds 0.5 2>> gpurun_out/r02e_err.log | tee gpurun_out/r02e_sweep_pf$a.jsonl | cut -c1-48,80-200
done
echo "== TTSP no resets"; timeout 300 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:262144:c0:s2 --seconds 0.5 --cfg beta_a=400 --cfg beta_b=0.5 2>> gpurun_out/r02e_err.log | cut -c1-48,80-200
CMD="python tools/sweep.py PointTTSP-v0:262144 --seconds 0.3"
timeout 300 $CMD > gpurun_out/r02e_plain_ttsp.jsonl 2>> gpurun_out/r02e_err.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1500 -c 2 -f -o gpurun_out/r02e_step_ttsp_262144 $CMD > gpurun_out/r02e_ncu1.log 2>&1; echo "ncu ttsp rc=$?"
tail -n 3 gpurun_out/r02e_err.log
