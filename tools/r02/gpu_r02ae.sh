#!/bin/bash
# round 2, call ae: encoder, fixed insert point (12) + M-block 1 drained on lane quadrants 2, 3; insert variants
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02ae_pytest.log
CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_tl.so timeout 300 python tools/enc_timeline.py > gpurun_out/r02ae_timeline.txt 2>&1; echo "timeline rc=$?"
grep "steady\|mean over\|inside" gpurun_out/r02ae_timeline.txt | head -3
timeout 300 python tools/bench_encode.py > gpurun_out/r02ae_encode_65536.json 2> gpurun_out/r02ae_encode.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02ae_encode.err
for v in ins8 ins18; do
  CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_$v.so timeout 300 python tools/bench_encode.py > gpurun_out/r02ae_encode_65536_$v.json 2>> gpurun_out/r02ae_encode.err; echo "$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02ae_encode_65536*.json')):
    try:
        d = json.load(open(f))
        print(f, 'healthy', d['healthy'], 'fused %.1f state %.1f head %.1f fwd %.1f us  frac %.3f same %s err %.2e' % (d['fused_us'], d['fused_from_state_us'], d['head_us'], d['forward_us'], d['roofline']['frac'], d['from_state_bit_identical'], d['max_abs_err_vs_torch_fp32']))
    except Exception as e:
        print(f, 'unreadable', e)
PY
