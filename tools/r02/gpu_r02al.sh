#!/bin/bash
# round 2, call al: zone encoder with FOUR tile slots of 64 rows; the L1 that rides in an L2 is the slot half a ring away; variants
# insert-point variants beside it
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/r02al_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02al_pytest.log
timeout 300 python tools/bench_encode.py > gpurun_out/r02al_encode_65536.json 2> gpurun_out/r02al_encode.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02al_encode.err
for v in ring3 ins4 ins20; do
  CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_$v.so timeout 300 python tools/bench_encode.py > gpurun_out/r02al_encode_65536_$v.json 2>> gpurun_out/r02al_encode.err; echo "$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02al_encode_65536*.json')):
    try:
        d = json.load(open(f))
        print(f, 'healthy', d['healthy'], 'fused %.1f state %.1f head %.1f fwd %.1f (two calls %.1f) us  frac %.3f same %s err %.2e' % (d['fused_us'], d['fused_from_state_us'], d['head_us'], d['forward_us'], d['forward_two_calls_us'], d['roofline']['frac'], d['from_state_bit_identical'], d['max_abs_err_vs_torch_fp32']))
    except Exception as e:
        print(f, 'unreadable', e)
PY
