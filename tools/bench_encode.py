"""Times crl_zone_encode (ZoneEnvModel's zone_net_ + mean-pool, fused, tcgen05) on the observations of a
PointTSP batch, next to the same op in plain torch (fp32 eager, and bf16 autocast as a library-GEMM
baseline).  CUDA events on the launching stream, warm-up, L2 flushed between timed calls by cycling
through input replicas larger than L2.  Prints one JSON line.

    python tools/bench_encode.py [--envs 65536] [--hidden 185] [--iters 50]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=65536)
    ap.add_argument('--hidden', type=int, default=185)
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--env', default='PointTSP-v0')
    args = ap.parse_args()
    B, h = args.envs, args.hidden
    spec = crl.ENV_SPECS[args.env]
    N, Z = spec.num_zones, spec.zone_dim
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8 + Z, h), torch.nn.ReLU(), torch.nn.Linear(h, h), torch.nn.ReLU(),
                              torch.nn.Linear(h, h)).cuda()
    comb = torch.nn.Linear(8 + h, h).cuda()
    sd = {f'zone_net_.{k}': v for k, v in net.state_dict().items()}
    sd.update({f'combine_net_.{k}': v for k, v in comb.state_dict().items()})
    enc = crl.ZoneEncoder(sd, num_zones=N)
    # real observations: a few random-action steps of the env; replicas so that inputs exceed L2
    env = crl.ZoneVecEnv(args.env, B)
    env.seed(1)
    obs = env.reset()
    for _ in range(8):
        obs, *_ = env.step_random(action_seed=4)
    in_bytes = B * (8 + N * Z) * 4
    reps = max(2, int(2 * 126e6 / in_bytes) + 1)
    obs_r = [obs['obs'].clone() for _ in range(reps)]
    zobs_r = [obs['zone_obs'].clone() for _ in range(reps)]
    out = torch.empty(B, h, device='cuda')

    def time_it(fn, iters):
        for i in range(5):
            fn(i % reps)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for i in range(iters):
            fn(i % reps)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / iters * 1e-3

    def torch_ref(i, dtype=None):
        o, z = obs_r[i], zobs_r[i]
        with torch.no_grad(), torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            x = torch.cat([o.view(B, 1, 8).expand(B, N, 8), z], dim=-1)
            return net(x).sum(dim=1) / N

    torch.backends.cuda.matmul.allow_tf32 = False
    t_fused = time_it(lambda i: enc.pooled(obs_r[i], zobs_r[i], out=out), args.iters)
    t_emb = time_it(lambda i: enc.zone_embedding(obs_r[i], zobs_r[i]), args.iters)      # + the head kernel with [0 | W3]
    t_fwd = time_it(lambda i: enc(obs_r[i], zobs_r[i]), args.iters)                      # ZoneEnvModel.forward: pooled + folded head
    t_fwd2 = time_it(lambda i: enc._head(enc.packed_head, obs_r[i], enc.pooled(obs_r[i], zobs_r[i], out=out)), args.iters)   # the two-call forward (fp32 pooled in between)
    pooled0 = enc.pooled(obs_r[0], zobs_r[0])
    t_head = time_it(lambda i: enc._head(enc.packed_head, obs_r[i], pooled0), args.iters)

    def torch_fwd(i, dtype=None):
        with torch.no_grad(), torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            return comb(torch.cat([obs_r[i], torch_ref(i, dtype)], dim=-1))
    ok = enc.healthy()
    t_fp32 = time_it(lambda i: torch_ref(i), max(3, args.iters // 10))
    t_bf16 = time_it(lambda i: torch_ref(i, torch.bfloat16), max(3, args.iters // 5))
    # the rows built from the state planes instead of read from zone_obs (crl_zone_encode_state); replicas of the env
    envs = [env] + [crl.ZoneVecEnv(args.env, B, env_offset=(k + 1) * B) for k in range(min(reps, 6) - 1)]
    for e in envs[1:]:
        e.seed(1 + e.cfg.env_offset)
        e.reset()
        for _ in range(8):
            e.step_random(action_seed=4)
    t_state = time_it(lambda i: enc.pooled_from_state(envs[i % len(envs)], out=out), args.iters)
    same = bool(torch.equal(enc.pooled_from_state(env), enc.pooled(env.obs, env.zone_obs)))
    t_fwd_fp32 = time_it(lambda i: torch_fwd(i), max(3, args.iters // 10))
    t_fwd_bf16 = time_it(lambda i: torch_fwd(i, torch.bfloat16), max(3, args.iters // 5))
    err_fwd = float((enc(obs_r[0], zobs_r[0]) - torch_fwd(0)).abs().max())
    err = float((enc.zone_embedding(obs_r[0], zobs_r[0]) - torch_ref(0)).abs().max())
    useful = B * N * 2 * ((8 + Z) * h + h * h)           # the kernel's two layers; the third runs on (B, h) in cuBLAS
    HP = (h + 31) // 32 * 32
    issued = (B + 7) // 8 * 128 * 2 * (16 * HP + HP * HP)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {}
    peak = peaks.get('bf16_tflops', 2250.0)
    print(json.dumps({
        'op': 'ZoneEnvModel.zone_net_ + mean over zones (env_model.py:56-78)', 'workload': f'{args.env}, {B} envs, N={N}, Z={Z}, h={h}',
        'healthy': ok, 'fused_us': t_fused * 1e6, 'fused_from_state_us': t_state * 1e6, 'from_state_bit_identical': same, 'head_us': t_head * 1e6, 'zone_embedding_us': t_emb * 1e6, 'forward_us': t_fwd * 1e6, 'forward_two_calls_us': t_fwd2 * 1e6,
        'torch_fp32_us': t_fp32 * 1e6, 'torch_bf16_autocast_us': t_bf16 * 1e6,
        'torch_forward_fp32_us': t_fwd_fp32 * 1e6, 'torch_forward_bf16_autocast_us': t_fwd_bf16 * 1e6,
        'forward_speedup_vs_torch_fp32': t_fwd_fp32 / t_fwd, 'forward_speedup_vs_torch_bf16_autocast': t_fwd_bf16 / t_fwd,
        'max_abs_err_forward_vs_torch_fp32': err_fwd,
        'envs_per_s': B / t_emb, 'speedup_vs_torch_fp32': t_fp32 / t_emb, 'speedup_vs_torch_bf16': t_bf16 / t_emb,
        'max_abs_err_vs_torch_fp32': err,
        'roofline': {'bound': 'tensor', 'achieved': useful / t_fused / 1e12, 'issued': issued / t_fused / 1e12, 'peak': peak,
                     'unit': 'TFLOP/s', 'frac': useful / t_fused / 1e12 / peak,
                     'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst)' if peaks else 'nominal'},
        'bytes': {'read': in_bytes, 'written': B * h * 4,
                  'unfused_intermediates': 'three (B N, h) fp32 activations = %d MB each' % (B * N * h * 4 // 2 ** 20)},
        'l2': f'{reps} input replicas ({reps * in_bytes >> 20} MB) cycled'}))


if __name__ == '__main__':
    main()
