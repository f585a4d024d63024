#!/bin/bash
# round 2, call ag: walls tests with the per-substep bar and the grazing allowance after ten fused substeps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_walls.py -m gpu -q -s > gpurun_out/r02ag_walls.log 2>&1; echo "walls rc=$?"; grep -v "^$" gpurun_out/r02ag_walls.log | grep "grazing\|worst\|passed\|failed\|Error\|FAILED" | tail -n 24
