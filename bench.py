#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused env.step() hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--env PointTSP-v0] [--envs 65536]

One "step" = one crl_step() launch over one batch of `--envs` envs (10 MuJoCo substeps +
task logic + full observation write per env, auto-reset on).  Default workload =
BASELINE.json configs[1]: PointTSP, 65,536 batched envs, random actions, 1 x B200.

A 65,536-env batch is 39 MB of state+observation and would sit in the 126 MB L2, so the
timed loop cycles a RING of independent replicas of the batch (each its own state,
layouts and action buffer; together > 2x L2): every launch finds its inputs in HBM.
`value` = launches x envs / device time (CUDA events, max over ranks), inputs resident
in HBM.  `e2e` = the same metric through ZoneVecEnv.step_host(): HOST numpy actions
in, HOST obs/reward/done out, copies inside the timed region.
`--impl reference` times the reference's CPU path: since mujoco-py / Safety Gym cannot be
installed here, that is the oracle's C port of it (oracle/crl_oracle.c), on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
METRIC = 'env_steps_per_sec'
UNIT = 'env-steps/s'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md, 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples taken before this call (warm-up) are dropped."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.t.join(timeout=2)
        self.rows = self.rows[max(0, self.first - 1):]
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == 'Active' for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def cpu_baseline(env_id, seconds, threads=None):
    """The oracle's C port of the reference CPU path, random actions, all host cores."""
    from oracle import c_oracle
    threads = threads or os.cpu_count() or 1
    rate, steps, wall = c_oracle.timed_rollout(env_id, threads=threads, seconds=seconds)
    return {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f'{steps} env-steps of {env_id} (fp64 C port of the reference path incl. numpy-legacy '
                      f'layout resets, random actions, {threads} threads x 1 env each, {wall:.1f} s)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = 4.0
    vals = []
    for _ in range(max(1, args.warmup if args.warmup < 3 else 1)):
        cpu_baseline(args.env, 1.0)
    k = max(1, min(args.steps, 5))
    for _ in range(k):
        vals.append(cpu_baseline(args.env, per_step))
    v = sum(x['value'] for x in vals) / len(vals)
    cb = dict(vals[-1], value=v)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': k,
        'warmup': 1, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.env} random-action rollouts on host cores; each step = {per_step:.0f} s sample'},
        'cpu_baseline': cb,
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'mujoco-py / Safety Gym are not installable here: the reference arm is the oracle C port',
    }))


def run_ours(args):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    import combinatorial_rl_tasks_b200 as crl
    from combinatorial_rl_tasks_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        # stdout carries the one JSON line and nothing else: while NCCL initialises (it prints its
        # version banner on stdout), file descriptor 1 points at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
            torch.cuda.set_device(local)
            dist.barrier()                                         # creates the communicator now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device(f'cuda:{local}')
    torch.cuda.set_device(dev)
    B = args.envs
    probe = crl.ZoneVecEnv(args.env, 32, device=dev)
    rd, wr = ctypes.c_int64(), ctypes.c_int64()
    probe.lib.crl_step_bytes(probe.cfg, ctypes.byref(rd), ctypes.byref(wr))
    step_bytes = rd.value + wr.value
    del probe
    R = max(2, -(-2 * L2_BYTES // (B * step_bytes)))             # ring > 2x L2
    K = max(1, args.steps)
    W = max(3, args.warmup)
    bank_kw = {}
    if args.bank:
        # the training configuration of the reference: a fixed set of K maps (make_train_env), here
        # K maps drawn by the device sampler itself and installed as a layout bank
        helper = crl.ZoneVecEnv(args.env, args.bank, device=dev)
        helper.seed(1)
        helper.reset()
        torch.cuda.synchronize()
        N = helper.spec.num_zones
        o = helper.origin.cpu().numpy()
        bank = {'xy0': o[:, :2], 'rot0': o[:, 2], 'zone_xy': helper.zone_xy.cpu().numpy().transpose(1, 0, 2)}
        if helper.zone_tmax is not None:
            w = helper.zone_tmax.cpu().numpy().astype(np.int64) & 0xffffffff
            bank['zone_max_steps'] = np.stack([w & 0xffff, w >> 16], 1).reshape(-1, w.shape[1])[:N].T
        if helper.cooldown is not None:
            bits = helper.aux[:, 3].view(torch.int32).cpu().numpy().astype(np.int64) & 0xffffffff
            bank['colours'] = (bits[:, None] >> 16 >> (2 * np.arange(N))) & 3
        bank_kw = dict(seed_mode='fixed_range', min_seed=1, max_seed=args.bank, layout_bank=bank)
        del helper
    envs = []
    for r in range(R):
        e = crl.ZoneVecEnv(args.env, B, device=dev, env_offset=(rank * R + r) * B,
                           prefetch_every=args.prefetch_every, prefetch_warps=args.prefetch_warps, **bank_kw)
        for kv in args.cfg:                                        # diagnostic overrides, e.g. --cfg beta_a=1000
            k, v = kv.split('=')
            setattr(e.cfg, k, type(getattr(e.cfg, k))(float(v)))
        e.seed(1 + (rank * R + r) * B)
        e.reset()
        envs.append(e)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    actions = [torch.rand(B, 2, device=dev, generator=g) * 2 - 1 for _ in range(R)]
    flags = 0 if args.no_auto_reset else _lib.STEP_AUTO_RESET
    if args.chained < 0:
        args.chained = 1 if B <= 131072 else 0

    def cycle(n=R):
        for e, a in list(zip(envs, actions))[:n]:
            e._step(a, flags, chained=args.chained)

    cycle()                                                        # eager: smem opt-in, first touch
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cycle()
    tails = {}
    for n in {K % R, W % R} - {0}:                                 # K and W need not be multiples of R
        tails[n] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(tails[n]):
            cycle(n)

    launches = {'prefetch': 0}

    def run(n_steps):
        # one graph replay = one step of every ring replica; at most every --prefetch-every steps the next-layout
        # slots are topped up on each replica's side stream (concurrent with the steps)
        for i in range(n_steps // R):
            graph.replay()
            if not args.bank:
                for e in envs:                                     # the envs' own sampler cadence (ZoneVecEnv.tick)
                    if e.tick():
                        launches['prefetch'] += 2 if e.spec.task == _lib.TASK_TSP else 3
        if n_steps % R:
            tails[n_steps % R].replay()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.5)                                                # nvidia-smi start-up
    run(W)
    torch.cuda.synchronize()
    launches['prefetch'] = 0
    if world > 1:
        dist.barrier()
    sampler.mark()
    reps = []
    n_rep = args.repeats
    for _ in range(n_rep):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0.record()
        run(K)
        t1.record()
        torch.cuda.synchronize()
        reps.append(t0.elapsed_time(t1))
    clocks = sampler.stop()
    ms = min(reps)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K * B / (ms * 1e-3)
    ms_per_step = ms / K

    # episode statistics: the path's only collective
    c = torch.zeros(8, dtype=torch.float64, device=dev)
    for e in envs:
        c += e.counters_dev
    if world > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    c = c.cpu().numpy()

    # e2e: host numpy in, host numpy out, through the public API
    e2e = None
    if rank == 0 or world > 1:
        env = envs[0]
        host_actions = [np.random.RandomState(7 + i).uniform(-1, 1, (B, 2)).astype(np.float32) for i in range(4)]
        N, Z = env.spec.num_zones, env.spec.zone_dim

        def timed_host_steps(delta):
            for i in range(3):
                env.step_host(host_actions[i % 4], delta=delta)
            ke = args.e2e_steps
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            rows = 0
            t0 = time.perf_counter()
            for i in range(ke):
                obs, rew, done, info = env.step_host(host_actions[i % 4], delta=delta)
                rows += env.delta_rows
            torch.cuda.synchronize()
            te = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([te], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                te = float(t.item())
            return world * ke * B / te, rows / ke

        full_rate, _ = timed_host_steps(False)
        rate, rows_per_step = timed_host_steps(True)
        is_delta = rows_per_step < B
        e2e = {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 8 * B,
               'd2h_bytes_per_step': int((32 + 8) * B + rows_per_step * (4 * N * Z + (4 if is_delta else 0)) + (16 if is_delta else 0)),
               'steps': args.e2e_steps,
               'api': 'ZoneVecEnv.step_host: host numpy actions in, host numpy obs/zone_obs/reward/done out, pinned '
                      'staging, copies + stream sync inside every call'
                      + ('; crl_step_host_delta, zero-copy: the step kernel reads the actions from and writes obs and result to the '
                         'pinned host buffers itself; of zone_obs only the rows that changed cross PCIe (mean %.1f of %d '
                         'rows per step); host buffers byte-identical to a full copy' % (rows_per_step, B) if is_delta
                         else '; crl_step_host: everything copied whole'),
               'full_copy_value': full_rate,
               'full_copy_d2h_bytes_per_step': (32 + 4 * N * Z + 8) * B}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    traffic = traffic_note = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_note = tj.get(f'{args.env}:{B}'), tj.get('_note')
    achieved = step_bytes * B / (ms_per_step * 1e-3) / 1e9
    canon = {'PointTSP-v0': 614, 'PointTTSP-v0': 734, 'ColourMatch-v0': 374}.get(args.env)
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.env}, {B} batched envs per launch, random actions, auto-reset {"off" if args.no_auto_reset else "on"}',
                   'prefetch_every': args.prefetch_every, 'prefetch_cadence': 'adaptive: interval doubles (up to 16x) while a round finds no empty slot', 'chained_steps': bool(args.chained),
                   'layout_bank': args.bank or None,
                   'envs_per_gpu_per_launch': B, 'ring_replicas': R,
                   'l2': f'ring of {R} independent {B}-env replicas ({R * B * step_bytes / 1e6:.0f} MB touched per '
                         f'cycle) > 2x the 126 MB L2, so every launch reads HBM',
                   'launch': 'CUDA graph of one ring cycle, replayed', 'repeats': n_rep, 'timing': 'best of repeats'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'traffic_note': traffic_note if traffic is not None else None,
                     'algorithmic_bytes_per_launch': step_bytes * B, 'peak_source': peak_src, 'kernel': 'step_kernel',
                     'bytes_per_env_step': step_bytes,
                     'canonical_bytes_per_env_step': canon,
                     'achieved_canonical': canon * B / (ms_per_step * 1e-3) / 1e9 if canon else None,
                     'frac_of_nominal_8TBs': achieved / 8000.0},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': K + launches['prefetch'],
        'gpu_launches_detail': {'step_kernel': K, 'prefetch_scan/layout/task kernels (side stream)': launches['prefetch']},
        'episode_stats': {'return_sum': c[0], 'episodes': c[1], 'successes': c[2], 'length_sum': c[3],
                          'resets_prefetched': c[4], 'resets_inline': c[5], 'chain_wait_timeouts': c[7],
                          'reduction': 'nccl all_reduce(sum)' if world > 1 else 'single rank'},
        'all_reps_ms': reps,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            out['cpu_baseline'] = cpu_baseline(args.env, args.cpu_seconds)
        except Exception as ex:  # the baseline is a reported number, never the product path
            out['cpu_baseline'] = {'error': repr(ex)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=64000)
    ap.add_argument('--warmup', type=int, default=6400)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--env', default='PointTSP-v0')
    ap.add_argument('--envs', type=int, default=65536)
    ap.add_argument('--repeats', type=int, default=1)
    ap.add_argument('--e2e-steps', type=int, default=200)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-auto-reset', action='store_true', help='diagnostic: finished envs keep stepping')
    ap.add_argument('--chained', type=int, default=-1,
                    help='1: back-to-back steps order themselves warp by warp (CRL_STEP_CHAINED); 0: whole-grid '
                         'wait; -1: chained when a launch is at most a wave or two (<= 131072 envs)')
    ap.add_argument('--prefetch-every', type=int, default=32,
                    help='top up the next-layout slots every N ring cycles (0: resets sample inline)')
    ap.add_argument('--bank', type=int, default=0,
                    help='fixed task set of K maps (make_train_env): resets copy from a layout bank, no sampler')
    ap.add_argument('--cfg', action='append', default=[], help='diagnostic: override a CrlConfig field, key=value')
    ap.add_argument('--prefetch-warps', type=int, default=0, help='background sampler warps per SM (0: default)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
