#!/bin/bash
# round 2, call bo: PCIe bytes and duration of the zero-copy host step kernel (per call), after its plain run
set -u
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i "^pcie" | head -20 > gpurun_out/r02bo_pcie_metrics.txt; cat gpurun_out/r02bo_pcie_metrics.txt | cut -c1-100
timeout 200 python tools/e2e_calls.py 40 > gpurun_out/r02bo_plain.txt 2>> gpurun_out/r02bo_err.log && cat gpurun_out/r02bo_plain.txt &&
timeout 400 ncu --metrics pcie__read_bytes.sum,pcie__write_bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:step_kernel -s 20 -c 8 --csv --log-file gpurun_out/r02bo_pcie.csv python tools/e2e_calls.py 40 > gpurun_out/r02bo_ncu.log 2>&1; echo "ncu rc=$?"
cut -d, -f5,11,12,13 gpurun_out/r02bo_pcie.csv | tail -n 42 | cut -c1-120
tail -n 3 gpurun_out/r02bo_err.log
