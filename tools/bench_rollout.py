"""Measurement of the SURVEY 8f rank-3 row (not the headline bench): on-device rollout collection
(Rollout.step: outputs written in place into the rollout's slots) and crl_gae against its HBM bytes.
Prints one JSON line per configuration."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import combinatorial_rl_tasks_b200 as crl
from combinatorial_rl_tasks_b200.rollout import Rollout

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists('MEASURED_PEAKS.json') else 6650.0
for env_id, B, T in [('PointTSP-v0', 65536, 128), ('PointTSP-v0', 262144, 128), ('PointTTSP-v3', 65536, 128)]:
    env = crl.ZoneVecEnv(env_id, B); env.seed(1)
    ro = Rollout(env, T)
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    acts = torch.rand(T, B, 2, device='cuda', generator=g) * 2 - 1
    vals = torch.randn(T, B, device='cuda', generator=g)
    nv = torch.randn(B, device='cuda', generator=g)
    goals = torch.zeros(B, dtype=torch.int32, device='cuda')
    def collect():
        ro.begin()
        for t in range(T):
            if env.spec.goals:
                env.set_goal(torch.where(env.goal < 0, goals, torch.full_like(goals, -1)))
            ro.step(t, acts[t], vals[t])
    collect(); collect(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); collect(); collect(); collect(); e1.record(); torch.cuda.synchronize()
    ms_collect = e0.elapsed_time(e1) / 3
    for _ in range(3): ro.finish(nv)
    torch.cuda.synchronize()
    # GAE alone; the rollout (T+1 slots of 8 B records + values + outputs = 20 B per frame-env) is larger than L2 at these sizes
    reps = 20
    e0.record()
    for _ in range(reps): ro.finish(nv)
    e1.record(); torch.cuda.synchronize()
    ms_gae = e0.elapsed_time(e1) / reps
    bytes_gae = (8 + 4 + 4 + 4 + (4 if env.spec.goals else 0)) * T * B
    print(json.dumps({'env': env_id, 'envs': B, 'frames': T,
                      'collect_env_steps_per_s': T * B / (ms_collect * 1e-3), 'collect_ms': ms_collect,
                      'gae_us': ms_gae * 1e3, 'gae_algorithmic_bytes': bytes_gae,
                      'gae_GBps': bytes_gae / (ms_gae * 1e-3) / 1e9, 'gae_frac_of_hbm_peak': bytes_gae / (ms_gae * 1e-3) / 1e9 / peak,
                      'rollout_bytes_MB': (T + 1) * B * (32 + 4 * env.spec.num_zones * env.spec.zone_dim + 8) / 1e6}))
    del ro, env
    torch.cuda.empty_cache()
