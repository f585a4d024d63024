#!/bin/bash
# round 2, call u: which operand paces the encoder's layer-2 MMAs (~300 cycles each in call t's timeline): timing-only
# variants with a contiguous A slab per K step / a K-major B descriptor / both
set -u
mkdir -p gpurun_out
for v in tl_a tl_b tl_ab; do
  echo "== $v"
  CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_$v.so timeout 300 python tools/enc_timeline.py > gpurun_out/r02u_timeline_$v.txt 2>&1; echo "rc=$?"
  grep -A12 "CTA 0" gpurun_out/r02u_timeline_$v.txt | tail -n 5; grep "steady\|mean over" gpurun_out/r02u_timeline_$v.txt | head -2
done
