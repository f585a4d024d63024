#!/bin/bash
# round 2, call s: the zone encoder as a warp-specialised pipeline (one issuing warp, mbarriers): parity, timing,
# and where the other slot's layer 1 is inserted into a layer 2 (CRL_ENC_INSERT_AFTER variants)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q -x > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02s_pytest.log
timeout 300 python tools/bench_encode.py > gpurun_out/r02s_encode_65536.json 2> gpurun_out/r02s_encode.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02s_encode.err
for v in 2 4 10 16; do
  CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_ins$v.so timeout 300 python tools/bench_encode.py > gpurun_out/r02s_encode_65536_ins$v.json 2>> gpurun_out/r02s_encode.err; echo "ins$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02s_encode_65536*.json')):
    try:
        d = json.load(open(f))
        print(f, 'healthy', d['healthy'], 'fused %.1f state %.1f head %.1f fwd %.1f us  frac %.3f same %s err %.2e' % (d['fused_us'], d['fused_from_state_us'], d['head_us'], d['forward_us'], d['roofline']['frac'], d['from_state_bit_identical'], d['max_abs_err_vs_torch_fp32']))
    except Exception as e:
        print(f, 'unreadable', e)
PY
