#!/bin/bash
# round 2, call bc: the prepared host call (crl_host_call_step, ABI 8): parity tests, per-call time against the eleven-argument call
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_walls.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02bc_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02bc_pytest.log
timeout 300 python tools/probe_e2e.py > gpurun_out/r02bc_probe.txt 2> gpurun_out/r02bc_probe.err; echo "probe rc=$?"; tail -n 3 gpurun_out/r02bc_probe.err; cat gpurun_out/r02bc_probe.txt
