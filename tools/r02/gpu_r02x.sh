#!/bin/bash
# round 2, call x: is the encoder's layer 2 paced by shared-memory bandwidth?  (an SS-mode 128 x 128 x 16 MMA reads 8 KB of
# operands per 64 cycles = all 128 B/clk of the SM's shared memory)  timing-only variants: the waiting warps sleep between
# polls / epilogue 1 does not store H1 / both
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02x_pytest.log
for v in tl_sleep tl_noh1 tl_both; do
  echo "== $v"
  CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_$v.so timeout 300 python tools/enc_timeline.py > gpurun_out/r02x_timeline_$v.txt 2>&1; echo "rc=$?"
  grep "steady\|mean over" gpurun_out/r02x_timeline_$v.txt | head -2
done
