"""oracle/zone_model.py (numpy restatement of ZoneEnvModel.forward, main/src/env_model.py:48-79) against
the fixture recorded from the REAL module (tests/golden/gen_golden_model.py)."""
import os

import numpy as np
import pytest

from oracle import zone_model as zm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'model_zone_env.npz')


def load(tag):
    g = dict(np.load(GOLDEN))
    sd = {k[len(tag) + 4:]: g[k] for k in g if k.startswith(tag + '_sd_')}
    return g, sd


@pytest.mark.parametrize('tag,h,n,z', [('tsp', 185, 15, 6), ('cm', 64, 6, 7)])
def test_numpy_restatement_matches_the_real_module(tag, h, n, z):
    g, sd = load(tag)
    assert sd['zone_net_.0.weight'].shape == (h, 8 + z) and sd['combine_net_.weight'].shape == (h, 8 + h)
    assert g[f'{tag}_zone_obs'].shape[1:] == (n, z)
    emb = zm.zone_embedding(sd, g[f'{tag}_obs'], g[f'{tag}_zone_obs'])
    out = zm.forward(sd, g[f'{tag}_obs'], g[f'{tag}_zone_obs'])
    # the module ran in float32, the restatement in float64
    assert np.max(np.abs(emb - g[f'{tag}_zone_emb'])) <= 2e-6
    assert np.max(np.abs(out - g[f'{tag}_out'])) <= 2e-6
    assert np.abs(g[f'{tag}_zone_emb']).max() > 0.5      # not a trivially small signal
