#!/bin/bash
# what bounds the zone encoder: timing-only variants (results are wrong by construction)
set -u
mkdir -p gpurun_out
for v in "" _crl_enc_diag_half_drain _crl_enc_diag_half_mma; do
  CRL_B200_LIB=combinatorial_rl_tasks_b200/libcrl_b200$v.so timeout 200 python tools/bench_encode.py --iters 30 > gpurun_out/bj_enc$v.json 2>> gpurun_out/bj_err.log
  python -c "
import json; d=json.load(open('gpurun_out/bj_enc$v.json')); print('variant [$v] fused_us', d['fused_us'], 'err', d['max_abs_err_vs_torch_fp32'])"
done
