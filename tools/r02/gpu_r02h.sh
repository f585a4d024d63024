#!/bin/bash
# round 2, call h: 112-register step kernels, two-tries-per-iteration sampler, bank fast path; sampler cadence / warps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r02h_pytest.log
timeout 900 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:262144:w3 PointTTSP-v0:262144:w4 PointTTSP-v0:262144:p64:w3 PointTTSP-v0:262144:p16:w3 PointTTSP-v0:262144:c0:s2:w3 \
   PointTTSP-v0:262144:b100 PointTTSP-v0:262144:b100:c0:s2 PointTTSP-v0:65536:w3 PointTTSP-v0:65536:c0:s2:w3 PointTTSP-v0:65536:b100 PointTTSP-v0:1048576:w3 \
   PointTSP-v0:65536 PointTSP-v0:262144 ColourMatch-v0:262144 ColourMatch-v0:262144:c0:s3 ColourMatch-v0:262144:w3 --seconds 0.6 2> gpurun_out/r02h_err.log | tee gpurun_out/r02h_sweep.jsonl | cut -c1-50,80-140,215-330
tail -n 3 gpurun_out/r02h_err.log
