"""Generate tests/golden/model_zone_env.npz by running the REAL ZoneEnvModel.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden_model.py

``main/src/env_model.py`` is imported unmodified (its ``import gym`` resolves to tests/golden/stubs;
the module never uses it).  Two models -- PointTSP with the reference's default width (--hidden-size 185,
scripts/train_ppo.py:66; 15 zones x 6) and ColourMatch (6 zones x 7) with width 64 --, default torch
initialisation under a fixed seed, fed observations of the recorded task fixtures.  Stored: the
state dict, the inputs, the output of ``model.forward`` and of ``model.zone_net_`` mean-pooled.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.join(HERE, 'stubs'), '/root/reference/main/src']

import env_model  # noqa: E402  (the reference's module)


class DictList(dict):
    """torch_ac.DictList: attribute access to a dict of tensors (format.py:27-28 builds one)."""
    __getattr__ = dict.__getitem__


def main():
    out = {}
    for tag, fixture, h in (('tsp', 'PointTSP_1000000_greedy.npz', 185), ('cm', 'ColourMatch_1000000_greedy.npz', 64)):
        g = np.load(os.path.join(HERE, fixture))
        obs = torch.tensor(g['obs'][::7][:48], dtype=torch.float)
        zone_obs = torch.tensor(g['zone_obs'][::7][:48], dtype=torch.float)
        torch.manual_seed(1234)
        model = env_model.ZoneEnvModel({'obs': tuple(obs.shape[1:]), 'zone_obs': tuple(zone_obs.shape[1:])}, h)
        with torch.no_grad():
            # default init gives |activations| ~ 0.1; scale the hidden layers up so that ReLUs and
            # magnitudes are exercised as after training
            for k, v in model.state_dict().items():
                if k.startswith('zone_net_.') and k.endswith('weight'):
                    v.mul_(2.0)
            y = model(DictList(obs=obs, zone_obs=zone_obs))
            bs, n = zone_obs.shape[:2]
            rep = obs.view(bs, 1, -1).expand(bs, n, obs.shape[1])
            emb = model.zone_net_(torch.cat([rep, zone_obs], dim=-1)).sum(dim=1) / n
        out[f'{tag}_obs'], out[f'{tag}_zone_obs'] = obs.numpy(), zone_obs.numpy()
        out[f'{tag}_out'], out[f'{tag}_zone_emb'] = y.numpy(), emb.numpy()
        for k, v in model.state_dict().items():
            out[f'{tag}_sd_{k}'] = v.numpy()
        print(tag, 'out', y.shape, 'abs mean', float(y.abs().mean()), 'emb abs mean', float(emb.abs().mean()),
              'emb abs max', float(emb.abs().max()))
    np.savez_compressed(os.path.join(HERE, 'model_zone_env.npz'), **out)


if __name__ == '__main__':
    main()
