#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -s > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/k_pytest.log
