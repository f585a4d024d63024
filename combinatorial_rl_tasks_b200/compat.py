"""ParallelEnv with the reference's exact call and return types, on top of ZoneVecEnv.

For unchanged reference code (main/src/torch_ac/algos/base.py:50,86,145; the hierarchical
collectors; zone-goals/src/torch_ac/algos/_hier_policy_opt.py:17-52): the constructor aside,
every method takes and returns what main/src/torch_ac/torch_utils/penv.py:23-69 and
zone-goals/src/torch_ac/torch_utils/penv.py:75-99 do --

    reset()                -> list of B obs dicts {'zone_obs': (N,Z), 'obs': (8,)}
    step(actions)          -> (obs, reward, done, info): four length-B tuples; finished envs are
                              reset inside the call (penv.py:9-10); reward float, done bool,
                              info dict with 'cost': 0, 'goal_met': True only on success
                              (Engine.step), and for the goal variants 'shaped_reward' /
                              'need_next_goal'; an env parked by WaitWrapper returns {}
    step_no_reset(actions) -> same without the reset
    set_goal(i, g), get_goal(i), needs_goal(), available_goals(i)

It costs what the reference's protocol costs -- per-env Python objects on the host every step
-- so it is the drop-in for correctness and for small batches; the tensor API of ZoneVecEnv is
the one that scales.  All arithmetic still happens in libcrl_b200.so on the GPU (step_host:
host arrays in, host arrays out); there is no CPU path here either.
"""
import numpy as np
import torch

from .vec_env import ZoneVecEnv, _EnvView


class ParallelEnv:
    def __init__(self, env_id, num_envs=None, num_training_tasks=100, hier=False, device='cuda:0', **kw):
        """``ParallelEnv([make_train_env(env_id, hier, num_training_tasks, rng_seed=...) for _ in
        range(num_envs)])`` (main/scripts/train_ppo.py:108-113, make_env.py:3-18): every reset
        re-seeds uniformly in [1, num_training_tasks]; ``hier=True`` wraps in WaitWrapper.
        The first argument may also be that very list, built with this module's make_train_env /
        make_test_env / make_fixed_env: the batch then takes its size from the list and its task,
        seeding rule and wrapper from the first element (the reference's lists are homogeneous)."""
        if isinstance(env_id, (list, tuple)):
            first, num_envs = env_id[0], len(env_id)
            env_id, hier, device = first.env_id, first.hier, first.device
            kw = dict(first.seeding, **kw)
        kw.setdefault('seed_mode', 'fixed_range')
        kw.setdefault('min_seed', 1)
        kw.setdefault('max_seed', num_training_tasks)
        first_seed = kw.pop('first_seed', None)
        self.vec = ZoneVecEnv(env_id, num_envs, device=device, wait=hier, **kw)
        if first_seed is not None:
            self.vec.seed(int(first_seed))
        self.envs = [_EnvView(self.vec, i) for i in range(num_envs)]   # the probes callers make on penv.envs[i]
        self.observation_space = self.vec.observation_space
        self.action_space = self.vec.action_space
        self._parked = np.zeros(num_envs, dtype=bool)

    # -- penv.py:46-66 ---------------------------------------------------------------
    def reset(self):
        self.vec.reset()
        self._parked[:] = False
        return self._obs_list(self.vec.obs.cpu().numpy(), self.vec.zone_obs.cpu().numpy())

    def step(self, actions):
        return self._step(actions, True)

    def step_no_reset(self, actions):
        return self._step(actions, False)

    def _step(self, actions, auto_reset):
        v = self.vec
        a = np.asarray([np.asarray(x, dtype=np.float32).reshape(2) for x in actions], dtype=np.float32)
        was_parked = self._parked.copy()
        if auto_reset:
            # a parked env is WaitWrapper's no-op followed by the worker's reset (wrappers.py:36-44,
            # penv.py:7-10): (first obs of its new episode, 0, True, {})
            obs, reward, done, info = v.step_host(a, auto_reset=True, wait=v.wait)
            self._parked[:] = False
        else:
            obs, reward, done, info = v.step_host(a, auto_reset=False, wait=v.wait)
            if v.wait:
                self._parked |= done
        goal_met, shaped, need = info['goal_met'], info.get('shaped_reward'), info.get('need_next_goal')
        infos = []
        for i in range(v.num_envs):
            if was_parked[i]:
                infos.append({})                               # WaitWrapper's no-op (wrappers.py:41-44)
                continue
            d = {'cost': 0}
            if goal_met[i]:
                d['goal_met'] = True
            if shaped is not None:
                d['shaped_reward'] = float(shaped[i])
                d['need_next_goal'] = bool(need[i])
            infos.append(d)
        return (tuple(self._obs_list(obs['obs'].copy(), obs['zone_obs'].copy())),
                tuple(float(r) for r in reward), tuple(bool(x) for x in done), tuple(infos))

    @staticmethod
    def _obs_list(obs, zone_obs):
        return [{'zone_obs': zone_obs[i], 'obs': obs[i]} for i in range(obs.shape[0])]

    # -- zone-goals penv.py:75-99 ----------------------------------------------------
    def set_goal(self, env_idx, goal):
        before = self.vec.counters()['goals_rejected']
        self.vec.set_goal_at(int(env_idx), int(np.asarray(goal)))
        assert self.vec.counters()['goals_rejected'] == before, 'set_goal: zone already visited or out of range'

    def get_goal(self, env_idx):
        g = int(self.vec.goal[env_idx].item())
        assert g >= 0, 'get_goal: no goal set'
        return self.vec.get_goal(env_idx).cpu().numpy().astype(np.float64)

    def needs_goal(self):
        return [bool(x) for x in self.vec.needs_goal().cpu().numpy()]

    def available_goals(self, env_idx):
        assert int(self.vec.goal[env_idx].item()) < 0, 'get_available_goals: a goal is set'
        return self.vec.available_goals(env_idx).cpu().numpy()

    def render(self):
        raise NotImplementedError


class GymEnv:
    """One env with the reference's single-env surface -- ``seed(s)``, ``reset() -> obs``,
    ``step(a) -> (obs, float, bool, info)`` without auto-reset (evaluate.py:48-60), the probes
    ``observation_space / action_space / unwrapped.num_cities`` -- as returned by the factories
    below.  Cheap to construct (no GPU work until the first reset), so a list of them can be
    handed to ParallelEnv exactly as train_ppo.py:108-113 does; used on its own it is a batch of
    one."""

    def __init__(self, env_id, hier, seeding, device='cuda:0'):
        from .config import ENV_SPECS
        from .spaces import Box, Dict
        self.env_id, self.hier, self.seeding, self.device = env_id, hier, dict(seeding), device
        spec = ENV_SPECS[env_id]
        N, Z = spec.num_zones, spec.zone_dim
        self.observation_space = Dict({'zone_obs': Box(-np.inf, np.inf, (N, Z)), 'obs': Box(-np.inf, np.inf, (8,))})
        self.action_space = Box(-1.0, 1.0, (2,))
        self.num_cities, self.goal_dim = N, 2
        self._pe = None

    @property
    def unwrapped(self):
        return self

    def _batch(self):
        if self._pe is None:
            self._pe = ParallelEnv([self])
        return self._pe

    def seed(self, seed):
        """Engine.seed: the next reset builds the map of seed + 1 (and, under a FixedSeedsWrapper,
        re-seeds first)."""
        self.seeding['first_seed'] = int(seed)
        if self._pe is not None:
            self._pe.vec.seed(int(seed))

    def reset(self):
        return self._batch().reset()[0]

    def step(self, action):
        obs, reward, done, info = self._batch().step_no_reset([action])
        return obs[0], reward[0], done[0], info[0]

    @property
    def goal_zone(self):
        return self._batch().envs[0].goal_zone

    def set_goal(self, goal):
        self._batch().set_goal(0, goal)

    def get_goal(self):
        return self._batch().get_goal(0)

    def get_available_goals(self):
        return self._batch().available_goals(0)

    def noop_obs(self):
        return self._batch().envs[0].noop_obs()


class TimeoutWrapper:
    """main/envs/wrappers.py:161-194 for any env with the single-env surface (``GymEnv`` here): the episode runs exactly
    ``max_timeout`` steps; once the inner env is done every further step returns its last observation and reward 0;
    ``done`` is raised only by the timer; ``info`` is ``{'timer': t}``.  Host-side bookkeeping only (the reference
    defines the class and uses it nowhere on the three tasks' training paths)."""

    def __init__(self, env, max_timeout=10000):
        self.env, self.max_timeout = env, max_timeout
        self.timer, self.inner_done, self.last_obs = 0, False, None

    def __getattr__(self, name):                    # gym.Wrapper forwards unknown attributes
        return getattr(self.env, name)

    def step(self, action):
        self.timer += 1
        if not self.inner_done:
            obs, rew, done, info = self.env.step(action)
            if done:
                self.inner_done, self.last_obs = True, obs
        else:
            obs, rew = self.last_obs, 0
        return obs, rew, self.timer == self.max_timeout, {'timer': self.timer}

    def reset(self):
        self.timer, self.inner_done, self.last_obs = 0, False, None
        return self.env.reset()


def make_train_env(env_name, hier=False, num_training_tasks=100, rng_seed=0, device='cuda:0'):
    """make_env.make_train_env (make_env.py:3-18): FixedSeedsWrapper over [1, num_training_tasks].
    ``rng_seed`` only decorrelated the per-process seed choosers; here the chooser is keyed by
    the env index."""
    return GymEnv(env_name, hier, dict(seed_mode='fixed_range', min_seed=1, max_seed=num_training_tasks), device)


def make_test_env(env_name, hier=False, seed=1000, device='cuda:0'):
    """make_env.make_test_env (:20-35): seeded once, every reset moves on to the next seed."""
    return GymEnv(env_name, False, dict(seed_mode='increment', first_seed=seed), device)


def make_fixed_env(env_name, hier=False, seed=1000, env_seed=0, device='cuda:0'):
    """make_env.make_fixed_env (:37-51): the same map (seed ``env_seed``) at every reset."""
    return GymEnv(env_name, False, dict(seed_mode='fixed_range', min_seed=env_seed, max_seed=env_seed), device)
