"""Per-tile timeline of the zone encoder's pipeline on one SM (diagnostic build -DCRL_ENC_TIMELINE=1:
`python -c "from combinatorial_rl_tasks_b200 import build; build.build(force=True, defines={'CRL_ENC_TIMELINE': 1}, out='combinatorial_rl_tasks_b200/libcrl_b200_tl.so')"`,
then `CRL_B200_LIB=.../libcrl_b200_tl.so python tools/enc_timeline.py`).  Cycles relative to the slot's first stamp:
issuer: L1 issued, H1 seen, L2 issued; warp 0 of the slot: L1 done, epilogue 1 over, rows staged, L2 block 0 done, epilogue 2 over."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combinatorial_rl_tasks_b200 as crl  # noqa: E402
from combinatorial_rl_tasks_b200 import _lib  # noqa: E402

B, h, N, Z = 65536, 185, 15, 6
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(8 + Z, h), torch.nn.ReLU(), torch.nn.Linear(h, h), torch.nn.ReLU(), torch.nn.Linear(h, h)).cuda()
comb = torch.nn.Linear(8 + h, h).cuda()
sd = {f'zone_net_.{k}': v for k, v in net.state_dict().items()}
sd.update({f'combine_net_.{k}': v for k, v in comb.state_dict().items()})
enc = crl.ZoneEncoder(sd, num_zones=N)
obs, zobs = torch.randn(B, 8, device='cuda'), torch.randn(B, N, Z, device='cuda')
for _ in range(3):
    enc.pooled(obs, zobs)
torch.cuda.synchronize()
lib = _lib.load()
tl = np.zeros((148, 2, 64, 16), dtype=np.int64)
assert lib.crl_debug_enc_timeline(tl.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(tl.nbytes)) == 0
names = ['L1 issued', 'H1 seen', 'L2 issued', 'L1 done', 'E1 over', 'X staged', 'L2b0 done', 'E2 over']
for cta in (0, 77):
    t0 = tl[cta, :, 0, 0].min()
    print(f'CTA {cta}: cycles since the first L1 issue; per tile k of slot g')
    print('  g  k ' + ' '.join(f'{n:>10}' for n in names))
    for k in list(range(0, 4)) + list(range(20, 24)):
        for g in (0, 1):
            print(f'  {g} {k:2d} ' + ' '.join(f'{int(v - t0):10d}' for v in tl[cta, g, k, :8]))
    per = (tl[cta, 0, 25, 7] - tl[cta, 0, 5, 7]) / 20.0
    print(f'  steady state: {per:.0f} cycles per tile pair (slot 0, tiles 5..25)')
    d = tl[cta, :, 5:25, :]
    print('  mean over tiles 5..24, both slots:  L1 done - L1 issued %.0f | E1 %.0f | E1 over -> H1 seen %.0f | L2 issue span %.0f | L2b0 done - H1 seen %.0f | E2 (block 0 warp) %.0f' % (
        (d[..., 3] - d[..., 0]).mean(), (d[..., 4] - d[..., 3]).mean(), (d[..., 1] - d[..., 4]).mean(), (d[..., 2] - d[..., 1]).mean(),
        (d[..., 6] - d[..., 1]).mean(), (d[..., 7] - d[..., 6]).mean()))
    print('  inside epilogue 2 (warp 0): first pair of TMEM loads %.0f | pool + store %.0f | second pair %.0f | pool + store %.0f | arrive %.0f' % (
        (d[..., 8] - d[..., 6]).mean(), (d[..., 9] - d[..., 8]).mean(), (d[..., 12] - d[..., 9]).mean(), (d[..., 13] - d[..., 12]).mean(),
        (d[..., 7] - d[..., 13]).mean()))
