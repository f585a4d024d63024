"""Generate tests/golden/gae_base_algo.npz by running the REAL rollout collector.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden_gae.py

Imported unmodified from /root/reference/main/src: ``torch_ac/algos/base.py`` (BaseAlgo.__init__
and collect_experiences, lines 110-249: the advantage recursion is :195-205, the P x T
flattening :225-231) over ``torch_ac/torch_utils/penv.py`` (ParallelEnv, real worker processes).
Stubbed: gym / safety_gym as in gen_golden.py (import-time only).  The envs and the actor-critic
are small deterministic fakes defined here -- the recursion does not care where rewards, dones
and values come from; what is recorded is what it made of them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/main'
sys.path[:0] = [ROOT, os.path.join(HERE, 'stubs'), os.path.join(REF, 'src'), REF, os.path.join(REF, 'envs')]

from torch_ac.algos.base import BaseAlgo  # noqa: E402


class Space:
    shape = (2,)


class FakeEnv:
    """Rewards in {0, 1, 1 + bonus}, episode ends at random: the shapes a zone task produces."""
    observation_space, action_space = Space(), Space()

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def reset(self):
        return self.rs.uniform(-1, 1, 5)

    def step(self, action):
        r = float(self.rs.rand() < 0.2) + (self.rs.rand() < 0.05) * self.rs.uniform(0, 20)
        return self.rs.uniform(-1, 1, 5), r, bool(self.rs.rand() < 0.12), {}


class FakeModel(torch.nn.Module):
    recurrent = False

    def forward(self, obs):
        value = (obs * torch.arange(1, 6, dtype=torch.float)).sum(-1) * 0.7 + 0.3
        return torch.distributions.Normal(obs[:, :2], torch.ones_like(obs[:, :2])), value


class Algo(BaseAlgo):
    def update_parameters(self):
        pass


def main():
    torch.manual_seed(0)
    T, P = 24, 4
    pre = lambda obss, device=None: torch.tensor(np.array(obss), device=device, dtype=torch.float)  # noqa: E731
    algo = Algo([FakeEnv(10 + i) for i in range(P)], FakeModel(), 'cpu', T, 0.99, 1e-3, 0.95, False,
                0.01, 0.5, 0.5, 1, pre, None)
    out = {}
    for k in range(2):                        # the second pass starts from a carried-over mask
        mask_in = algo.mask.clone()
        exps, logs = algo.collect_experiences()
        with torch.no_grad():
            _, next_value = algo.acmodel(pre(algo.obs))
        out.update({f'rewards_{k}': algo.rewards.numpy().copy(), f'values_{k}': algo.values.numpy().copy(),
                    f'masks_{k}': algo.masks.numpy().copy(), f'mask_in_{k}': mask_in.numpy(),
                    f'last_mask_{k}': algo.mask.numpy().copy(), f'next_value_{k}': next_value.numpy().copy(),
                    f'advantages_{k}': algo.advantages.numpy().copy(),
                    f'exps_advantage_{k}': exps.advantage.numpy().copy(),
                    f'exps_returnn_{k}': exps.returnn.numpy().copy()})
    np.savez_compressed(os.path.join(HERE, 'gae_base_algo.npz'), discount=0.99, gae_lambda=0.95, **out)
    print('dones per pass:', [(1 - out[f'masks_{k}']).sum() for k in range(2)])


if __name__ == '__main__':
    main()
