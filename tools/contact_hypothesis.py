#!/usr/bin/env python
"""Writes profiles/r02_contact_hypothesis.json: the numbers tests/test_contact_hypothesis.py asserts on (what the
alternative reading of the sphere/floor contact, SURVEY A.3, would change)."""
import glob, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_contact_hypothesis import rollout  # noqa: E402

out = {'note': 'open-loop replays with oracle/mj_point.py under the two readings of constraint_force; reach = max distance '
               'from the start over the first 400 env steps of the recorded action sequence'}
full = np.tile(np.array([1.0, 0.0]), (300, 1))
for model in ('excluded', 'active'):
    _, v = rollout(full, model)
    out[f'terminal_speed_m_per_s_{model}'] = float(np.hypot(v[0], v[1]))
out['reference_velocity_normaliser'] = 1.5
ep = {}
for path in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'PointTSP_10000*_*.npz'))):
    g = np.load(path)
    acts = g['actions'][:400]
    ep[os.path.basename(path)] = {m: float(np.max(np.linalg.norm(rollout(acts, m)[0], axis=1))) for m in ('excluded', 'active')}
out['reach_m'] = ep
json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02_contact_hypothesis.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
