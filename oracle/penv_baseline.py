"""CPU baselines in the REFERENCE'S OWN TOPOLOGY (TEST / BENCH INFRASTRUCTURE ONLY -- never on the
product path; only bench.py's cpu_baseline / --impl reference legs and tests/ import this).

The reference runs its envs as Python objects, one per OS process, driven over pipes:

* ``eval_protocol``  -- main/scripts/evaluate.py:47-72: ONE env in ONE process,
  ``make_fixed_env(env, seed, env_seed)`` for env_seed = 1000000 .. 1000099, ``reset()``, ``step()`` until
  ``done``, success = ``info['goal_met']``.  BASELINE.json configs[0].  (The policy there is a trained
  agent; the benchmark metric is random-action rollouts, so actions are ``action_space.sample()``-like
  U(-1,1)^2 draws.)
* ``pipe_vector_env`` -- main/src/torch_ac/torch_utils/penv.py:4-59: ``ParallelEnv``: env 0 lives in the
  parent, every other env in its own ``multiprocessing.Process`` behind a ``Pipe``; ``step`` sends one
  pickled action per worker, each worker steps (and resets on done) and sends the pickled
  ``(obs dict, reward, done, info)`` back.  Same message protocol, same process count = host cores.

``impl='python'`` steps oracle/zone_env.py (numpy restatement of the reference's Python task code over the
restated physics: the closest thing to the reference's per-step Python cost that runs here -- mujoco-py /
Safety Gym are not installable); ``impl='c'`` steps the C twin through ctypes (what remains when the
Python task code costs nothing: IPC and pickling only).
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _make_env(env_id, impl, env_seed=None, train_tasks=100, rng_seed=0):
    """An env with the reference's surface: reset() -> obs dict, step(a) -> (obs, reward, done, info)."""
    if impl == 'python':
        from oracle import zone_env as ze
        if env_seed is not None:
            return ze.make_fixed_env(env_id, seed=rng_seed, env_seed=env_seed)
        return ze.make_train_env(env_id, num_training_tasks=train_tasks, rng_seed=rng_seed)
    from oracle import c_oracle

    class _CFixed:                                  # FixedSeedsWrapper(min_seed, max_seed) over the C twin
        def __init__(self):
            self.env = c_oracle.CEnv(env_id)
            self.lo, self.hi = (env_seed, env_seed) if env_seed is not None else (1, train_tasks)
            self.rng = np.random.default_rng(seed=rng_seed)

        def reset(self):
            self.env.seed(int(self.rng.integers(low=self.lo, high=self.hi + 1, size=1)[0]))
            return self.env.reset()

        def step(self, action):
            return self.env.step(action)

    return _CFixed()


def eval_protocol(env_id='PointTSP-v0', n_maps=100, first_seed=1000000, impl='c', max_seconds=20.0):
    """evaluate.py:47-72 with random actions.  Returns dict(value = env-steps/s, maps, steps, successes, wall)."""
    rs = np.random.RandomState(0)
    steps = succ = maps = 0
    t0 = time.perf_counter()
    for env_seed in range(first_seed, first_seed + n_maps):
        env = _make_env(env_id, impl, env_seed=env_seed, rng_seed=0)
        env.reset()
        while True:
            obs, reward, done, info = env.step(rs.uniform(-1, 1, 2))
            steps += 1
            if done:
                succ += bool(info.get('goal_met', False))
                break
        maps += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    wall = time.perf_counter() - t0
    return {'value': steps / wall, 'maps': maps, 'steps': steps, 'successes': succ, 'wall_s': wall, 'impl': impl,
            'sample': f'{maps} of the maps {first_seed}..{first_seed + n_maps - 1}, one process, one env, step to done'}


def _worker(conn, env_id, impl, rng_seed):
    """penv.py:4-21."""
    env = _make_env(env_id, impl, rng_seed=rng_seed)
    while True:
        cmd, data = conn.recv()
        if cmd == 'step':
            obs, reward, done, info = env.step(data)
            if done:
                obs = env.reset()
            conn.send((obs, reward, done, info))
        elif cmd == 'reset':
            conn.send(env.reset())
        elif cmd == 'kill':
            return


def pipe_vector_env(env_id='PointTSP-v0', n_envs=None, seconds=5.0, impl='c'):
    """ParallelEnv.step in the reference's process topology.  Returns dict(value = env-steps/s, ...)."""
    n_envs = n_envs or os.cpu_count() or 1
    ctx = mp.get_context('fork')
    env0 = _make_env(env_id, impl, rng_seed=1)
    locals_, procs = [], []
    for i in range(1, n_envs):
        local, remote = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(remote, env_id, impl, 1 + 10000 * i), daemon=True)
        p.start()
        remote.close()
        locals_.append(local)
        procs.append(p)
    for local in locals_:
        local.send(('reset', None))
    env0.reset()
    for local in locals_:
        local.recv()
    rs = np.random.RandomState(1)
    calls = 0
    t0 = time.perf_counter()
    while True:
        actions = rs.uniform(-1, 1, (n_envs, 2))
        for local, a in zip(locals_, actions[1:]):
            local.send(('step', a))
        obs, reward, done, info = env0.step(actions[0])
        if done:
            obs = env0.reset()
        results = list(zip(*[(obs, reward, done, info)] + [local.recv() for local in locals_]))   # noqa: F841
        calls += 1
        if calls % 16 == 0 and time.perf_counter() - t0 >= seconds:
            break
    wall = time.perf_counter() - t0
    for local in locals_:
        local.send(('kill', None))
    for p in procs:
        p.join(timeout=2)
    return {'value': calls * n_envs / wall, 'envs': n_envs, 'processes': n_envs, 'calls': calls, 'wall_s': wall,
            'impl': impl, 'sample': f'{calls} ParallelEnv.step calls over {n_envs} envs (1 in the parent, '
                                    f'{n_envs - 1} worker processes, Pipe + pickle), {wall:.1f} s'}
