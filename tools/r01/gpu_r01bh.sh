#!/bin/bash
# final check of round 1 at HEAD: whole GPU suite + smoke()
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/bh_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/bh_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bh_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/bh_smoke.log
