#!/bin/bash
# round 2, call y: where epilogue 2's ~2,000 cycles go (stamps inside it)
set -u
mkdir -p gpurun_out
CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_tl.so timeout 300 python tools/enc_timeline.py > gpurun_out/r02y_timeline.txt 2>&1; echo "rc=$?"
grep "steady\|mean over\|inside" gpurun_out/r02y_timeline.txt
