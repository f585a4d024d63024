#!/bin/bash
# re-entry check at HEAD: the whole GPU suite, smoke(), the default bench
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/aw_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/aw_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aw_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/aw_smoke.log
timeout 300 python bench.py > gpurun_out/aw_bench.json 2> gpurun_out/aw_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/aw_bench.json
