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
    def __init__(self, env_id, num_envs, num_training_tasks=100, hier=False, device='cuda:0', **kw):
        """``ParallelEnv([make_train_env(env_id, hier, num_training_tasks, rng_seed=...) for _ in
        range(num_envs)])`` (main/scripts/train_ppo.py:108-113, make_env.py:3-18): every reset
        re-seeds uniformly in [1, num_training_tasks]; ``hier=True`` wraps in WaitWrapper."""
        kw.setdefault('seed_mode', 'fixed_range')
        kw.setdefault('min_seed', 1)
        kw.setdefault('max_seed', num_training_tasks)
        self.vec = ZoneVecEnv(env_id, num_envs, device=device, wait=hier, **kw)
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
            obs, reward, done, info = v.step_host(a, auto_reset=True)
            self._parked[:] = False                            # a parked env runs into the step limit and restarts
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
