"""CPU oracle, part 5: ZoneEnvModel.forward (main/src/env_model.py:48-79) restated in numpy.

TEST INFRASTRUCTURE ONLY.  PINNED by tests/golden/model_zone_env.npz, recorded by running the REAL
``ZoneEnvModel`` (tests/golden/gen_golden_model.py imports main/src/env_model.py unmodified).

    obs_repeated = obs.view(bs, 1, 8).expand(bs, N, 8)                         (:70)
    zone_emb = zone_net_(cat([obs_repeated, zone_obs], -1)).sum(dim=1) / N      (:73)
    out = combine_net_(cat([obs, zone_emb], -1))                                (:75)
zone_net_ = Linear(8 + Z, h), ReLU, Linear(h, h), ReLU, Linear(h, h)           (:57-63)
"""
import numpy as np


def zone_embedding(sd, obs, zone_obs, dtype=np.float64):
    obs, zone_obs = np.asarray(obs, dtype), np.asarray(zone_obs, dtype)
    B, N, _ = zone_obs.shape
    x = np.concatenate([np.broadcast_to(obs[:, None, :], (B, N, obs.shape[1])), zone_obs], axis=-1)
    for i in (0, 2, 4):
        x = x @ np.asarray(sd[f'zone_net_.{i}.weight'], dtype).T + np.asarray(sd[f'zone_net_.{i}.bias'], dtype)
        if i != 4:
            x = np.maximum(x, 0)
    return x.sum(axis=1) / N


def forward(sd, obs, zone_obs, dtype=np.float64):
    emb = zone_embedding(sd, obs, zone_obs, dtype)
    x = np.concatenate([np.asarray(obs, dtype), emb], axis=-1)
    return x @ np.asarray(sd['combine_net_.weight'], dtype).T + np.asarray(sd['combine_net_.bias'], dtype)
