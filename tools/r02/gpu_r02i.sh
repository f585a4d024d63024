#!/bin/bash
# round 2, call i: ring replicas x streams x sampler warps
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep.py PointTTSP-v0:262144:c0:s2 PointTTSP-v0:262144:c0:s2:w3:p64 PointTTSP-v0:262144:c0:s2:w4 PointTTSP-v0:262144:c0:s3:r3:w3 PointTTSP-v0:262144:c0:s4:r4:w3 PointTTSP-v0:262144:c0:s4:r4:w4 \
   PointTTSP-v0:65536:c0:s3:w3 PointTTSP-v0:65536:c0:s6:w3 PointTTSP-v0:1048576:c0:s2:w3 PointTTSP-v0:1048576:c0:s2:w4 \
   PointTSP-v0:262144:c0:s2 PointTSP-v0:262144:c0:s3:r3 PointTSP-v0:1048576:c0:s2 PointTSP-v0:65536:c0:s7 PointTSP-v0:65536:c0:s4:r8 PointTSP-v0:65536:c1:s1:r8 \
   ColourMatch-v0:262144:c0:s3 ColourMatch-v0:262144:c0:s4:r4 ColourMatch-v0:1048576:c0:s2 ColourMatch-v0:1048576:c0:s3:r3 ColourMatch-v0:65536 ColourMatch-v0:65536:c0:s4 --seconds 0.6 2> gpurun_out/r02i_err.log | tee gpurun_out/r02i_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-36s frac %.4f best %.4f mean %.4f  %.2f us  inline %d R %d' % (d['case'], d['frac'], d['best'], d['mean'], d['us_per_step'], d['resets_inline'], d['R']))"
tail -n 3 gpurun_out/r02i_err.log
