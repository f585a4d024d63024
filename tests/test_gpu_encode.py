"""crl_zone_encode (tcgen05 kernel, csrc/crl_encode.cu) + the (B, h) third Linear = ZoneEnvModel's zone_net_
+ mean-pool (main/src/env_model.py:56-78; the mean commutes with the affine third layer).  The kernel multiplies bf16 operands with fp32 accumulation, the
reference runs in fp32; bars (written here, measured values printed), relative to the largest
reference value of the batch (and at least 1):
  vs a torch reference that rounds the same operands to bf16 (same arithmetic) ... 1e-3
     (what remains are activations that round to the neighbouring bf16 because one side summed in
      fp32 and the other in fp64; measured 1e-7 .. 4e-4)
  vs the REAL module's fp32 output (fixture) and vs torch fp32 ................... 1e-2
     (bf16 operand rounding through the two wide layers; measured 3e-3 .. 5e-3)
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import zone_model as zm  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'model_zone_env.npz')
BF16_TWIN_RTOL = 1e-3
FP32_RTOL = 1e-2


@pytest.fixture(scope='module')
def crl():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import combinatorial_rl_tasks_b200 as m
    return m


def bf16_twin(sd, obs, zone_obs, head=True):
    """The kernel's arithmetic in torch: operands rounded to bf16, products and sums in fp32 (fp64 here)."""
    r = lambda t: t.to(torch.bfloat16).to(torch.float64)
    B, N, _ = zone_obs.shape
    x = r(torch.cat([obs[:, None, :].expand(B, N, obs.shape[1]), zone_obs], dim=-1))
    # the biases ride in the GEMMs as bf16 hi + lo parts (lo only where the operand has a second spare column)
    def bias(b, room_for_lo):
        hi = r(b)
        return hi + r(b.to(torch.float64) - hi) if room_for_lo else hi
    h, in_dim = sd['zone_net_.0.weight'].shape
    x = r(torch.relu(x @ r(sd['zone_net_.0.weight']).T + bias(sd['zone_net_.0.bias'], in_dim + 1 not in (16, 32))).to(torch.float32))
    x = torch.relu(x @ r(sd['zone_net_.2.weight']).T + bias(sd['zone_net_.2.bias'], True))
    pooled = x.sum(dim=1) / N                        # the fused kernel's output (fp32)
    if not head:                                     # the third Linear in fp32 on (B, h): what the pooled kernel is tested with
        return (pooled @ sd['zone_net_.4.weight'].to(torch.float64).T + sd['zone_net_.4.bias'].to(torch.float64)).to(torch.float32)
    # crl_encoder_head: [obs, pooled] and the weights rounded to bf16, the bias as hi + lo
    return (r(pooled.to(torch.float32)) @ r(sd['zone_net_.4.weight']).T + bias(sd['zone_net_.4.bias'], True)).to(torch.float32)


def fp32_ref(sd, obs, zone_obs):
    B, N, _ = zone_obs.shape
    x = torch.cat([obs[:, None, :].expand(B, N, obs.shape[1]), zone_obs], dim=-1).to(torch.float64)
    for i in (0, 2, 4):
        x = x @ sd[f'zone_net_.{i}.weight'].to(torch.float64).T + sd[f'zone_net_.{i}.bias'].to(torch.float64)
        if i != 4:
            x = torch.relu(x)
    return (x.sum(dim=1) / N).to(torch.float32)


@pytest.mark.parametrize('tag,n', [('tsp', 15), ('cm', 6)])
def test_fixture_of_the_real_module(crl, tag, n):
    g = dict(np.load(GOLDEN))
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]).cuda() for k in g if k.startswith(tag + '_sd_')}
    enc = crl.ZoneEncoder(sd, num_zones=n)
    obs, zobs = torch.from_numpy(g[f'{tag}_obs']).cuda(), torch.from_numpy(g[f'{tag}_zone_obs']).cuda()
    emb = enc.zone_embedding(obs, zobs)
    out = enc(obs, zobs)
    torch.cuda.synchronize()
    assert enc.healthy()
    e_twin = float((emb - bf16_twin(sd, obs, zobs)).abs().max())
    e_emb = float(np.abs(emb.cpu().numpy() - g[f'{tag}_zone_emb']).max())
    e_out = float(np.abs(out.cpu().numpy() - g[f'{tag}_out']).max())
    scale = max(1.0, float(np.abs(g[f'{tag}_zone_emb']).max()))
    print(f'{tag}: vs bf16 twin {e_twin:.2e}, vs real module zone_emb {e_emb:.2e}, forward {e_out:.2e}, scale {scale:.2f}')
    assert e_twin <= BF16_TWIN_RTOL * scale and e_emb <= FP32_RTOL * scale and e_out <= FP32_RTOL * scale
    # the numpy oracle agrees with what was compared against
    assert np.abs(zm.zone_embedding({k: v.cpu().numpy() for k, v in sd.items()}, g[f'{tag}_obs'], g[f'{tag}_zone_obs'])
                  - g[f'{tag}_zone_emb']).max() <= 2e-6


@pytest.mark.parametrize('B,N,Z,h,D', [(4099, 15, 6, 185, 8), (1, 15, 7, 185, 8), (777, 5, 6, 96, 8), (20000, 6, 7, 32, 8),
                                       (3001, 15, 7, 185, 10), (515, 6, 7, 64, 16)])
def test_random_batches_against_torch(crl, B, N, Z, h, D):
    """Ragged batch sizes (partial tiles, more tiles than SMs), every supported width class.  D = per-env
    features: 8 for ZoneEnvModel; 8 + goal_dim (ZoneEnvGoalModel, zone-goals/src/env_model.py) or 8 + n_skills
    (ZoneEnvSkillModel, main/src/env_model.py:81-117) when the caller concatenates goal / one-hot skill to obs --
    then the layer-1 input is 32 wide (two K steps)."""
    gen = torch.Generator(device='cuda').manual_seed(B + h)
    rn = lambda *s, scale=1.0: (torch.randn(*s, device='cuda', generator=gen) * scale)
    sd = {'zone_net_.0.weight': rn(h, D + Z, scale=0.4), 'zone_net_.0.bias': rn(h, scale=0.2),
          'zone_net_.2.weight': rn(h, h, scale=1.5 / h ** 0.5), 'zone_net_.2.bias': rn(h, scale=0.2),
          'zone_net_.4.weight': rn(h, h, scale=1.5 / h ** 0.5), 'zone_net_.4.bias': rn(h, scale=0.2),
          'combine_net_.weight': rn(h, D + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
    obs, zobs = rn(B, D), rn(B, N, Z)
    enc = crl.ZoneEncoder(sd, num_zones=N)
    out = torch.full((B + 3, h), 7.0, device='cuda')           # guard rows: nothing may be written past B
    pooled = enc.pooled(obs, zobs, out=out[:B])
    emb = torch.nn.functional.linear(pooled, sd['zone_net_.4.weight'], sd['zone_net_.4.bias'])
    torch.cuda.synchronize()
    assert enc.healthy() and bool((out[B:] == 7.0).all()) and bool((pooled >= 0).all())
    twin, ref = bf16_twin(sd, obs, zobs, head=False), fp32_ref(sd, obs, zobs)
    e_twin, e_ref = float((emb - twin).abs().max()), float((emb - ref).abs().max())
    print(f'B={B} N={N} Z={Z} h={h} D={D}: vs bf16 twin {e_twin:.2e}, vs fp32 {e_ref:.2e}, |ref| max {float(ref.abs().max()):.2f}')
    scale = max(1.0, float(ref.abs().max()))
    assert e_twin <= BF16_TWIN_RTOL * scale and e_ref <= FP32_RTOL * scale
    # the head kernel (crl_encoder_head): zone_embedding = L3(pooled), forward = combine_net_([obs, zone_emb]); guard rows
    out2 = torch.full((B + 3, h), 7.0, device='cuda')
    emb_k = enc.zone_embedding(obs, zobs, out=out2[:B])
    fwd_k = enc(obs, zobs)
    fwd_ref = torch.nn.functional.linear(torch.cat([obs, ref], -1).double(), sd['combine_net_.weight'].double(),
                                         sd['combine_net_.bias'].double()).float()
    torch.cuda.synchronize()
    e_head_twin, e_head = float((emb_k - bf16_twin(sd, obs, zobs)).abs().max()), float((emb_k - ref).abs().max())
    e_fwd = float((fwd_k - fwd_ref).abs().max())
    print(f'   head: zone_emb vs bf16 twin {e_head_twin:.2e}, vs fp32 {e_head:.2e}; forward vs fp32 {e_fwd:.2e}')
    assert enc.healthy() and bool((out2[B:] == 7.0).all())
    assert e_head_twin <= BF16_TWIN_RTOL * scale and e_head <= FP32_RTOL * scale
    assert e_fwd <= FP32_RTOL * max(1.0, float(fwd_ref.abs().max()))
    # second call on the same encoder (barrier phases, TMEM re-allocation) gives the same bits
    assert torch.equal(enc.pooled(obs, zobs), pooled)


def test_encoder_on_the_env_outputs(crl):
    """The fused step's outputs feed the encoder directly (obs (B,8), zone_obs (B,N,Z) as CrlOut lays them out)."""
    g = dict(np.load(GOLDEN))
    sd = {k[7:]: torch.from_numpy(g[k]).cuda() for k in g if k.startswith('tsp_sd_')}
    env = crl.ZoneVecEnv('PointTSP-v0', 4096)
    env.seed(3)
    obs = env.reset()
    for _ in range(5):
        obs, *_ = env.step_random(action_seed=2)
    enc = crl.ZoneEncoder(sd, num_zones=15)
    y = enc(obs)
    ref = torch.nn.functional.linear(torch.cat([obs['obs'], fp32_ref(sd, obs['obs'], obs['zone_obs'])], -1),
                                     sd['combine_net_.weight'], sd['combine_net_.bias'])
    assert enc.healthy() and y.shape == (4096, 185)
    assert float((y - ref).abs().max()) <= FP32_RTOL * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('env_id', ['PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0', 'PointTSP-v1', 'PointTTSP-v3'])
def test_rows_built_from_the_state_planes_are_the_rows_the_step_writes(crl, env_id):
    """crl_zone_encode_state builds the zone part of the encoder's input from the state planes (zone centres,
    visited / colour bits, timeouts, cooldowns, step count).  It must see bit for bit the rows the step writes to
    zone_obs -- so `pooled` is bit-identical between the two paths -- also for envs that have visited zones, run
    cooldowns, been reset; and a rollout with env.write_zone_obs = False (CRL_STEP_NO_ZONE_OBS: the step neither
    builds nor writes zone_obs) gives the same embeddings as one that materialises them."""
    B, h = 3000, 96
    N, Z = crl.ENV_SPECS[env_id].num_zones, crl.ENV_SPECS[env_id].zone_dim
    gen = torch.Generator(device='cuda').manual_seed(5)
    rn = lambda *s_, scale=1.0: (torch.randn(*s_, device='cuda', generator=gen) * scale)
    sd = {'zone_net_.0.weight': rn(h, 8 + Z, scale=0.4), 'zone_net_.0.bias': rn(h, scale=0.2),
          'zone_net_.2.weight': rn(h, h, scale=0.15), 'zone_net_.2.bias': rn(h, scale=0.2),
          'zone_net_.4.weight': rn(h, h, scale=0.15), 'zone_net_.4.bias': rn(h, scale=0.2),
          'combine_net_.weight': rn(h, 8 + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
    enc = crl.ZoneEncoder(sd, num_zones=N)
    full, lean = crl.ZoneVecEnv(env_id, B), crl.ZoneVecEnv(env_id, B)
    for e in (full, lean):
        e.seed(77)
        e.cfg.num_steps = 60                                   # short episodes: auto-resets inside the test
        e.reset()
    lean.write_zone_obs = False
    frozen = lean.zone_obs.clone()
    rs = np.random.RandomState(2)
    for t in range(90):
        a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
        if t % 7 == 3:                                         # zone events: visited bits / colour cycles / cooldowns
            z = int(rs.randint(N))
            for e in (full, lean):
                e.pose[t % 5::5, :2] = e.zone_xy[z, t % 5::5, :]
        full.step(a)
        lean.step(a)
        if t % 9 == 0 or t > 84:
            want = enc.pooled(full.obs, full.zone_obs)
            assert torch.equal(enc.pooled_from_state(full), want), (env_id, t)      # same env, both paths
            assert torch.equal(enc.pooled_from_state(lean), want), (env_id, t)      # the env that never wrote zone_obs
            assert torch.equal(enc.forward_from_state(lean), enc(full.obs, full.zone_obs)), (env_id, t)
    torch.cuda.synchronize()
    assert enc.healthy()
    assert torch.equal(lean.zone_obs, frozen)                  # CRL_STEP_NO_ZONE_OBS: untouched since the reset
    assert torch.equal(lean.obs, full.obs) and torch.equal(lean.result, full.result) and torch.equal(lean.aux, full.aux)
    assert full.counters()['episodes'] == lean.counters()['episodes'] >= B


def test_unsupported_shapes_are_refused(crl):
    from combinatorial_rl_tasks_b200 import _lib
    import ctypes
    lib = _lib.load()
    n = ctypes.c_int64()
    assert lib.crl_encoder_packed_bytes(_lib.CrlEncoderShape(8, 6, 185, 15), ctypes.byref(n)) == 0 and n.value == 256 * 192 * 2 + 256 * 32
    assert lib.crl_encoder_packed_bytes(_lib.CrlEncoderShape(8, 6, 256, 15), ctypes.byref(n)) == -4    # two resident weights do not fit
    assert lib.crl_encoder_packed_bytes(_lib.CrlEncoderShape(20, 12, 64, 15), ctypes.byref(n)) == -2     # no room for the ones column in the 32-wide input
    assert lib.crl_encoder_head_packed_bytes(_lib.CrlEncoderShape(8, 6, 185, 15), ctypes.byref(n)) == 0 and n.value == 256 * 208 * 2


@pytest.mark.parametrize('env_id,B,h', [('PointTSP-v0', 4099, 185), ('ColourMatch-v0', 1000, 185), ('PointTTSP-v0', 129, 185),
                                        ('PointTSP-v1', 77, 185), ('PointTSP-v0', 3000, 100), ('PointTTSP-v0', 2000, 64),
                                        ('ColourMatch-v0', 900, 127), ('PointTSP-v0', 1500, 33)])
def test_fused_forward_is_the_two_call_forward(crl, env_id, B, h):
    """crl_encoder_forward (the zone kernel writes the head kernel's bf16 operand image, no fp32 pooled in between, one
    bulk copy per 128 envs in the head) gives bit for bit what crl_zone_encode + crl_encoder_head give -- the head rounds
    pooled to bf16 either way -- from the materialised zone_obs and from the state planes; ragged batches; hidden widths
    whose last warp reaches beyond the image's width (100, 33), one M-block (64, 100) and two (127: the ones rows, 185)."""
    N, Z = crl.ENV_SPECS[env_id].num_zones, crl.ENV_SPECS[env_id].zone_dim
    gen = torch.Generator(device='cuda').manual_seed(B)
    rn = lambda *s_, scale=1.0: (torch.randn(*s_, device='cuda', generator=gen) * scale)
    sd = {'zone_net_.0.weight': rn(h, 8 + Z, scale=0.4), 'zone_net_.0.bias': rn(h, scale=0.2),
          'zone_net_.2.weight': rn(h, h, scale=0.1), 'zone_net_.2.bias': rn(h, scale=0.2),
          'zone_net_.4.weight': rn(h, h, scale=0.1), 'zone_net_.4.bias': rn(h, scale=0.2),
          'combine_net_.weight': rn(h, 8 + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
    enc = crl.ZoneEncoder(sd, num_zones=N)
    env = crl.ZoneVecEnv(env_id, B)
    env.seed(5)
    obs = env.reset()
    for _ in range(30):
        obs, *_ = env.step_random(action_seed=2)
    two = enc._head(enc.packed_head, obs['obs'], enc.pooled(obs['obs'], obs['zone_obs']))
    guard = torch.full((B + 2, h), 7.0, device='cuda')
    fused = enc._forward(obs['obs'], obs['zone_obs'], out=guard[:B])
    from_state = enc.forward_from_state(env)
    torch.cuda.synchronize()
    assert enc.healthy() and enc._workspace(B) is not None
    assert torch.equal(fused, two) and torch.equal(from_state, two)
    assert bool((guard[B:] == 7.0).all()) and bool(torch.isfinite(fused).all())


PRECISE_RTOL = 1e-4       # the like-for-like bar against the fp32 module; measured ~1e-5


@pytest.mark.parametrize('tag,n', [('tsp', 15), ('cm', 6)])
def test_precise_mode_against_the_real_modules_fp32_output(crl, tag, n):
    """crl_zone_encode_precise (split-bf16 operands, three MMAs per product, fp32 biases) against the fixture recorded
    from the REAL ZoneEnvModel in fp32: zone_emb and forward within 1e-4 of the largest value."""
    g = dict(np.load(GOLDEN))
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]).cuda() for k in g if k.startswith(tag + '_sd_')}
    enc = crl.ZoneEncoder(sd, num_zones=n)
    torch.backends.cuda.matmul.allow_tf32 = False
    obs, zobs = torch.from_numpy(g[f'{tag}_obs']).cuda(), torch.from_numpy(g[f'{tag}_zone_obs']).cuda()
    emb = enc.zone_embedding_precise(obs, zobs)
    out = enc.forward_precise(obs, zobs)
    torch.cuda.synchronize()
    assert enc.healthy()
    e_emb = float(np.abs(emb.cpu().numpy() - g[f'{tag}_zone_emb']).max()) / max(1.0, float(np.abs(g[f'{tag}_zone_emb']).max()))
    e_out = float(np.abs(out.cpu().numpy() - g[f'{tag}_out']).max()) / max(1.0, float(np.abs(g[f'{tag}_out']).max()))
    e_fast = float(np.abs(enc(obs, zobs).cpu().numpy() - g[f'{tag}_out']).max()) / max(1.0, float(np.abs(g[f'{tag}_out']).max()))
    print(f'{tag}: precise zone_emb {e_emb:.2e}, forward {e_out:.2e} (fast bf16 forward {e_fast:.2e})')
    assert e_emb <= PRECISE_RTOL and e_out <= PRECISE_RTOL


@pytest.mark.parametrize('B,N,Z,h', [(4099, 15, 6, 185), (1, 15, 7, 185), (777, 5, 6, 96), (20000, 6, 7, 32), (3000, 15, 7, 128)])
def test_precise_mode_random_batches(crl, B, N, Z, h):
    gen = torch.Generator(device='cuda').manual_seed(B + h + 1)
    rn = lambda *s_, scale=1.0: (torch.randn(*s_, device='cuda', generator=gen) * scale)
    sd = {'zone_net_.0.weight': rn(h, 8 + Z, scale=0.4), 'zone_net_.0.bias': rn(h, scale=0.2),
          'zone_net_.2.weight': rn(h, h, scale=1.5 / h ** 0.5), 'zone_net_.2.bias': rn(h, scale=0.2),
          'zone_net_.4.weight': rn(h, h, scale=1.5 / h ** 0.5), 'zone_net_.4.bias': rn(h, scale=0.2),
          'combine_net_.weight': rn(h, 8 + h, scale=0.1), 'combine_net_.bias': rn(h, scale=0.1)}
    obs, zobs = rn(B, 8), rn(B, N, Z)
    enc = crl.ZoneEncoder(sd, num_zones=N)
    out = torch.full((B + 3, h), 7.0, device='cuda')
    pooled = enc.pooled_precise(obs, zobs, out=out[:B])
    emb = torch.nn.functional.linear(pooled.double(), sd['zone_net_.4.weight'].double(), sd['zone_net_.4.bias'].double()).float()
    torch.cuda.synchronize()
    assert enc.healthy() and bool((out[B:] == 7.0).all()) and bool((pooled >= 0).all())
    ref = fp32_ref(sd, obs, zobs)
    err = float((emb - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    print(f'precise B={B} N={N} Z={Z} h={h}: {err:.2e} of the largest reference value')
    assert err <= PRECISE_RTOL
