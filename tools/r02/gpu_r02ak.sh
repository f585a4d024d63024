#!/bin/bash
# round 2, call ak: precise mode of the zone encoder (split-bf16, three MMAs per product): parity vs the fp32 module, timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q -s -k "precise" > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest rc=$?"; grep "precise\|passed\|failed\|Error" gpurun_out/r02ak_pytest.log | tail -n 16
timeout 300 python - > gpurun_out/r02ak_time.txt 2>&1 <<'PY'
import torch, combinatorial_rl_tasks_b200 as crl
B, h, N, Z = 65536, 185, 15, 6
g = torch.Generator(device='cuda').manual_seed(0)
rn = lambda *s, scale=1.0: torch.randn(*s, device='cuda', generator=g) * scale
sd = {'zone_net_.0.weight': rn(h, 8 + Z, scale=0.3), 'zone_net_.0.bias': rn(h, scale=0.1), 'zone_net_.2.weight': rn(h, h, scale=0.1), 'zone_net_.2.bias': rn(h, scale=0.1),
      'zone_net_.4.weight': rn(h, h, scale=0.1), 'zone_net_.4.bias': rn(h, scale=0.1), 'combine_net_.weight': rn(h, 8 + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
enc = crl.ZoneEncoder(sd, num_zones=N)
obs, zobs = rn(B, 8), rn(B, N, Z)
for fn, name in ((lambda: enc.pooled_precise(obs, zobs), 'pooled_precise'), (lambda: enc.forward_precise(obs, zobs), 'forward_precise'), (lambda: enc.pooled(obs, zobs), 'pooled (fast)')):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, '%.1f us' % (e0.elapsed_time(e1) / 20 * 1e3), 'healthy', enc.healthy())
PY
cat gpurun_out/r02ak_time.txt | tail -n 5
