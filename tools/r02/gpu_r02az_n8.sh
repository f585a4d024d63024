#!/bin/bash
# round 2, call az (8 GPUs): why is the end-to-end call 3x slower with eight ranks?  zero-copy vs staged vs copy engine, k of 8 active
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/probe_e2e_multi.py > gpurun_out/r02az_probe_n8.jsonl 2> gpurun_out/r02az_probe_n8.err; echo "probe rc=$?"
tail -n 3 gpurun_out/r02az_probe_n8.err; cat gpurun_out/r02az_probe_n8.jsonl | cut -c1-230
lspci -tv 2>/dev/null | head -60 > gpurun_out/r02az_lspci.txt; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/r02az_pcie.txt 2>&1; cat gpurun_out/r02az_pcie.txt
