"""Generate tests/golden/hard_*.npz and hardgoals_*.npz by running the REAL hard-instance code.

Run in the build container only (needs /root/reference), once per tree (the two reference
packages both call themselves ``envs``):

    python tests/golden/gen_golden_hard.py main
    python tests/golden/gen_golden_hard.py zone-goals

``main``: PointTSP-v4 / PointTSP-v5 as registered in main/envs/__init__.py:52-81, :110-118 --
``envs/TSP_hard_env.py`` (TSPHardEnv over TSPEnv) imported unmodified, driven through
make_env.make_fixed_env exactly like the other fixtures (recorder: gen_golden.record_episode).
``zone-goals``: the goal-conditioned registrations of the same ids (zone-goals/envs/__init__.py:52-81,
zone-goals/envs/TSP_hard_env.py over TSPNextCityEnv; recorder: gen_golden_goals.record_episode);
stored under the ids 'zone-goals/PointTSP-v4' / 'zone-goals/PointTSP-v5'.
Stubs as in gen_golden.py: Safety Gym's Engine (fixed ``robot_locations`` / ``zones_locations`` /
``robot_rot`` included) is the oracle's restatement, oracle/sg_engine.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

MAIN = [('PointTSP-v4', 1000000, 'greedy', 1000), ('PointTSP-v4', 1000001, 'random', 1000),
        ('PointTSP-v5', 1000000, 'greedy', 250), ('PointTSP-v5', 1000002, 'idle', 250)]
GOALS = [('PointTSP-v4', 1000000, 'near', 1000), ('PointTSP-v5', 1000001, 'far', 300)]


def main(tree):
    if tree == 'main':
        from tests.golden import gen_golden as gg
        for env_id, seed, mode, max_len in MAIN:
            out = gg.record_episode(env_id, seed, mode, max_len)
            name = f'hard_{env_id}_{seed}_{mode}.npz'
            np.savez_compressed(os.path.join(HERE, name), **out)
            print(name, 'steps', len(out['reward']), 'return', out['reward'].sum(), 'goal_met', bool(out['goal_met'].any()),
                  'first colours', out['zone_obs'][0][:, 2].astype(int))
        # the vector-env protocol (penv.py: step, reset on done) over make_test_env envs of the 250-step instance
        # (the reference's make_train_env refuses the hard instances, make_env.py:3-18; evaluation seeds once and
        # lets Engine.reset increment the seed): every env auto-resets twice in 600 steps, each time re-sampling its
        # distractors around the fixed cities
        from envs.make_env import make_test_env
        es = [make_test_env('PointTSP-v5', seed=1000 + 50 * i) for i in range(3)]
        rs = np.random.RandomState(5)
        obs = [e.reset() for e in es]
        rec = {k: [] for k in ('actions', 'obs', 'zone_obs', 'reward', 'done', 'goal_met')}
        rec['obs'].append(np.array([o['obs'] for o in obs])); rec['zone_obs'].append(np.array([o['zone_obs'] for o in obs]))
        for t in range(600):
            acts, row = [], []
            for i, e in enumerate(es):
                a = gg.policy('PointTSP-v5', obs[i], rs, 'greedy', t)
                o, r, d, info = e.step(a)
                if d:
                    o = e.reset()
                obs[i] = o
                acts.append(a); row.append((float(r), bool(d), bool(info.get('goal_met', False))))
            rec['actions'].append(np.array(acts))
            rec['obs'].append(np.array([o['obs'] for o in obs])); rec['zone_obs'].append(np.array([o['zone_obs'] for o in obs]))
            rec['reward'].append([x[0] for x in row]); rec['done'].append([x[1] for x in row]); rec['goal_met'].append([x[2] for x in row])
        out = {k: np.array(v) for k, v in rec.items()}
        out['actions'] = out['actions'].astype(np.float32)
        out['env_id'] = np.array('PointTSP-v5')
        np.savez_compressed(os.path.join(HERE, 'hardvec_PointTSP-v5.npz'), **out)
        print('hardvec_PointTSP-v5.npz', 'resets', int(out['done'].sum()), 'return', out['reward'].sum())
    else:
        from tests.golden import gen_golden_goals as gg
        for env_id, seed, mode, max_len in GOALS:
            out = gg.record_episode(env_id, seed, mode, max_len)
            out['env_id'] = np.array('zone-goals/' + env_id)
            name = f'hardgoals_{env_id}_{seed}_{mode}.npz'
            np.savez_compressed(os.path.join(HERE, name), **out)
            print(name, 'steps', len(out['reward']), 'return', out['reward'].sum(), 'shaped', out['shaped_reward'].sum(),
                  'reached', int(out['need_next_goal'].sum()), 'goal_met', bool(out['goal_met'].any()))


if __name__ == '__main__':
    main(sys.argv[1])
