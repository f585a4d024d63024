#!/bin/bash
# HEAD of round 1: whole GPU suite, smoke(), default bench, then the example with the fused encoder at full size
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/bc_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/bc_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bc_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/bc_smoke.log
timeout 300 python bench.py > gpurun_out/bc_bench.json 2> gpurun_out/bc_bench.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bc_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bc_bench_ref.json 2>> gpurun_out/bc_bench.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bc_bench_ref.json
timeout 200 python examples/collect_ppo.py --envs 65536 --frames 32 --updates 3 > gpurun_out/bc_example_torch.log 2>&1; echo "example rc=$?"; tail -2 gpurun_out/bc_example_torch.log
timeout 200 python examples/collect_ppo.py --envs 65536 --frames 32 --updates 3 --fused-encoder > gpurun_out/bc_example_fused.log 2>&1; echo "example fused rc=$?"; tail -2 gpurun_out/bc_example_fused.log
