#!/bin/bash
# round 2, call as: encoder tests at the final build; insert point re-tuned on the final kernel
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q 2>&1 | tail -n 2
for v in "" _ins6 _ins18 _ins24; do
  if [ -n "$v" ]; then export CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200$v.so; fi
  timeout 300 python tools/bench_encode.py > gpurun_out/r02as_encode$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02as_encode$v.json')); print('$v fused %.1f state %.1f fwd %.1f frac %.3f' % (d['fused_us'], d['fused_from_state_us'], d['forward_us'], d['roofline']['frac']))"
done
