"""On-device rollout storage and GAE: what BaseAlgo.collect_experiences keeps per update
(main/src/torch_ac/algos/base.py:86-108 buffers, :110-249 loop), without leaving the GPU.

The reference stores each frame by copying Python lists into torch tensors
(base.py:145-166) after pickling them through a pipe.  Here a rollout owns T + 1 output slots
and, before step t, points the env's outputs at slot t + 1 (ZoneVecEnv.bind_outputs): the step
kernel writes obs, zone_obs and the result record straight into the rollout, so storing a frame
is free.  Advantages are computed by crl_gae from those records (bit-identical to
base.py:195-205).

    ro = Rollout(env, num_frames_per_proc=128)
    obs = ro.begin()                              # reset, or carry over from the previous rollout
    for t in range(ro.T):
        dist, value = acmodel(obs)                # obs: {'obs': (B,8), 'zone_obs': (B,N,Z)} views of slot t
        action = dist.sample()
        obs, reward, done, info = ro.step(t, action, value, dist.log_prob(action))
    exps = ro.finish(next_value=acmodel(obs)[1])  # advantages, returns, flattened P x T like base.py:225-231

Layout note: tensors are kept T-major, (T, B, ...), as the reference's own buffers are;
``finish`` returns both the (T, B) tensors and, on request, the reference's P x T flattening
(env-major), which is a transposing copy -- flat (non-recurrent) PPO samples frames at random and
does not need it.
"""
import ctypes

import torch

from . import _lib


class Rollout:
    def __init__(self, env, num_frames_per_proc, discount=0.99, gae_lambda=0.95):
        self.env, self.T = env, int(num_frames_per_proc)
        self.discount, self.gae_lambda = float(discount), float(gae_lambda)
        B, N, Z, T, dev = env.num_envs, env.spec.num_zones, env.spec.zone_dim, self.T, env.device
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        # slot t = what the policy sees before step t; slot t + 1 = what step t returned
        self.obs = z(T + 1, B, 8)
        self.zone_obs = z(T + 1, B, N, Z)
        self.result = z(T + 1, B, 8, dtype=torch.uint8)
        self.shaped = z(T + 1, B) if env.spec.goals else None
        self.actions = z(T, B, 2)
        self.values = z(T, B)
        self.log_probs = z(T, B, 2)
        self.advantages = z(T, B)
        self.returns = z(T, B)
        self.rewards = self.result.view(torch.float32)[1:, :, 0]         # (T, B) view: reward of step t
        self.dones = self.result[1:, :, 4].view(torch.bool)              # (T, B) view
        self._started = False
        assert B % 4 == 0, 'slots of zone_obs must start 16-byte aligned: num_envs must be a multiple of 4'
        # one validated output binding per slot, installed with a few attribute stores per step
        self._bindings = [env.prepare_outputs(self.obs[t], self.zone_obs[t], self.result[t],
                                              None if self.shaped is None else self.shaped[t]) for t in range(T + 1)]
        self._obs_views = [{'zone_obs': self.zone_obs[t], 'obs': self.obs[t]} for t in range(T + 1)]
        self._action_views = [self.actions[t] for t in range(T)]

    def _bind(self, t):
        self.env.bind_outputs(binding=self._bindings[t])

    def obs_at(self, t):
        return self._obs_views[t]

    def begin(self, reset=None):
        """Start a rollout: the first one (or ``reset=True``) resets every env into slot 0; later
        ones carry slot T of the previous rollout over to slot 0 (base.py keeps self.obs and
        self.mask across calls the same way)."""
        if reset is None:
            reset = not self._started
        if reset:
            self._bind(0)
            self.env.reset()
            self.result[0].zero_()                        # mask = 1: no env has just finished
        else:
            self.obs[0].copy_(self.obs[self.T])
            self.zone_obs[0].copy_(self.zone_obs[self.T])
            self.result[0].copy_(self.result[self.T])
        self._started = True
        return self.obs_at(0)

    def step(self, t, actions, values=None, log_probs=None):
        """ParallelEnv.step for frame t (auto-reset on), outputs written into slot t + 1."""
        a = self._action_views[t]
        if actions.data_ptr() != a.data_ptr():            # a policy may also write into ro.actions[t] directly
            a.copy_(actions)
        if values is not None:
            self.values[t].copy_(values)
        if log_probs is not None:
            self.log_probs[t].copy_(log_probs)
        self._bind(t + 1)
        _, reward, done, info = self.env.step(a)
        return self._obs_views[t + 1], reward, done, info

    def masks(self):
        """(T, B) float: masks[t] = 1 - done of the step before frame t (base.py:151-152)."""
        return 1.0 - self.result[:self.T, :, 4].float()

    def finish(self, next_value, flatten=False):
        """base.py:182-205: advantages and returns of the rollout just collected.  Rewards are the
        shaped rewards for the goal-conditioned variants (base.py:155-159)."""
        env, T, B = self.env, self.T, self.env.num_envs
        nv = next_value.to(device=env.device, dtype=torch.float32).reshape(B).contiguous()
        with torch.cuda.device(env.device):
            _lib.check(env.lib.crl_gae(self.result.data_ptr(),
                                       None if self.shaped is None else self.shaped.data_ptr(),
                                       self.values.data_ptr(), nv.data_ptr(), ctypes.c_double(self.discount),
                                       ctypes.c_double(self.gae_lambda), T, B, self.advantages.data_ptr(),
                                       self.returns.data_ptr(), env._stream()))
        env.gpu_launches += 1
        out = {'advantage': self.advantages, 'returnn': self.returns, 'value': self.values, 'action': self.actions,
               'log_prob': self.log_probs, 'reward': self.rewards if self.shaped is None else self.shaped[1:],
               'obs': {'obs': self.obs[:T], 'zone_obs': self.zone_obs[:T]}}
        if flatten:                                      # T x P -> P x T -> P * T, base.py:225-231
            pt = lambda x: x.transpose(0, 1).reshape((-1,) + tuple(x.shape[2:]))
            out = {k: ({kk: pt(vv) for kk, vv in v.items()} if isinstance(v, dict) else pt(v)) for k, v in out.items()}
        return out
