#!/bin/bash
# round 2, call ai: crl_encoder_forward; head kernel launched programmatically behind the zone kernel (prologue under its tail)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02ai_pytest.log
timeout 300 python tools/bench_encode.py > gpurun_out/r02ai_encode_65536.json 2> gpurun_out/r02ai_encode.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02ai_encode.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02ai_encode_65536.json'))
print('healthy', d['healthy'], 'fused %.1f state %.1f head %.1f emb %.1f fwd %.1f (two calls %.1f) us  frac %.3f same %s err %.2e fwd err %.2e' % (d['fused_us'], d['fused_from_state_us'], d['head_us'], d['zone_embedding_us'], d["forward_us"], d["forward_two_calls_us"], d["roofline"]['frac'], d['from_state_bit_identical'], d['max_abs_err_vs_torch_fp32'], d['max_abs_err_forward_vs_torch_fp32']))
PY
