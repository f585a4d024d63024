#!/bin/bash
# TimedTSP step regression: A/B of build variants with the probe (slots pre-filled), then ncu
set -u
mkdir -p gpurun_out
P=$PWD/combinatorial_rl_tasks_b200
for v in libcrl_b200.so libcrl_b200_crl_inline_reset.so libcrl_b200_crl_no_minblocks.so; do
  echo "== $v"
  CRL_B200_LIB=$P/$v timeout 300 python tools/probe_prefetch.py PointTTSP-v0 262144 2>&1 | head -1
done
CMD_T="python bench.py --env PointTTSP-v0 --envs 262144 --steps 1400 --warmup 100 --no-cpu-baseline --e2e-steps 2 --prefetch-every 32"
$CMD_T > gpurun_out/plain_t2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1300 -c 2 -f -o gpurun_out/r01b_step_ttsp_262144 $CMD_T > gpurun_out/ncu_t2.log 2>&1
echo "ncu ttsp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prefetch_layout -s 3 -c 1 -f -o gpurun_out/r01b_prefetch_layout_ttsp $CMD_T > gpurun_out/ncu_p2.log 2>&1
echo "ncu prefetch rc=$?"
