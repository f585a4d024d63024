"""oracle/gae.py against the fixture recorded from the REAL BaseAlgo.collect_experiences
(tests/golden/gen_golden_gae.py): the advantage recursion must be reproduced bit for bit."""
import os

import numpy as np

from oracle import gae as og

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'gae_base_algo.npz'))


def test_gae_matches_reference_bit_exact():
    for k in range(2):
        adv = og.gae(G[f'rewards_{k}'], G[f'values_{k}'], G[f'masks_{k}'], G[f'last_mask_{k}'], G[f'next_value_{k}'],
                     float(G['discount']), float(G['gae_lambda']))
        assert adv.dtype == np.float32
        assert np.array_equal(adv.view(np.uint32), G[f'advantages_{k}'].view(np.uint32)), k
        assert np.array_equal(og.flatten_pt(adv), G[f'exps_advantage_{k}'])
        assert np.array_equal(og.flatten_pt(G[f'values_{k}'] + adv), G[f'exps_returnn_{k}'])
        # the fixture exercises what matters: episode ends inside the rollout and a carried mask
        assert (G[f'masks_{k}'] == 0).any() and G[f'rewards_{k}'].max() > 1.5
    assert np.array_equal(G['mask_in_1'], G['last_mask_0']) and np.array_equal(G['masks_1'][0], G['mask_in_1'])
