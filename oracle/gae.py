"""CPU oracle: the advantage recursion of the reference's rollout collector.

TEST INFRASTRUCTURE ONLY (see oracle/mj_point.py header): tests/ and nothing else imports this.

Restates main/src/torch_ac/algos/base.py:195-205 in numpy float32, one IEEE operation at a
time in the order torch evaluates the reference's expressions (no fused multiply-add):

    delta      = (rewards[i] + (discount * next_value) * next_mask) - values[i]
    advantages = delta + ((discount * gae_lambda) * next_advantage) * next_mask

with next_* taken from step i+1 inside the rollout and from (next_value, the mask in force
after the last step, 0) at its end; ``discount * gae_lambda`` is a Python double product that
torch rounds to float32 when it meets the tensor.  ``flatten_pt`` is the P x T flattening of
base.py:225-231 (the k-th block of T consecutive entries belongs to env k).

PINNED by tests/golden/gae_base_algo.npz, recorded by running the REAL BaseAlgo.collect_experiences
(tests/golden/gen_golden_gae.py): bit-exact.
"""
import numpy as np

f32 = np.float32


def gae(rewards, values, masks, last_mask, next_value, discount, gae_lambda):
    """rewards, values, masks: (T, P) float32; masks[i] = 1 - done of the step BEFORE step i;
    last_mask, next_value: (P,).  Returns advantages (T, P) float32."""
    rewards, values, masks = (np.asarray(x, dtype=f32) for x in (rewards, values, masks))
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    g, gl = f32(discount), f32(discount * gae_lambda)
    nm, nv, na = np.asarray(last_mask, f32), np.asarray(next_value, f32), np.zeros(rewards.shape[1], f32)
    for i in reversed(range(T)):
        delta = (rewards[i] + (g * nv) * nm) - values[i]
        adv[i] = delta + ((gl * na) * nm)
        nm, nv, na = masks[i], values[i], adv[i]
    return adv


def flatten_pt(x):
    """T x P -> P x T -> P * T (base.py:225-231)."""
    return np.ascontiguousarray(np.swapaxes(x, 0, 1)).reshape((-1,) + x.shape[2:])
