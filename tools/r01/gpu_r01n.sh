#!/bin/bash
# rollout + GAE tests, e2e probe
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n_pytest.log
timeout 300 python tools/probe_e2e.py 2>&1 | tail -14
