"""ZoneEncoder: the reference's ``ZoneEnvModel`` (main/src/env_model.py:48-79) for the rollout-time
forward, with the wide part of its per-zone network and the mean-pool fused into one tensor-core kernel
(include/crl_b200.h: crl_zone_encode; csrc/crl_encode.cu).  The last Linear of ``zone_net_`` is affine, so
the mean over zones is taken before it: the kernel returns ``pooled = mean_z relu(L2(relu(L1(.))))`` and
``zone_emb = L3(pooled)`` is a (B, h) library GEMM.

    model = ZoneEnvModel(obs_space, h_dim)                  # the reference's module, trained as usual
    enc = ZoneEncoder(model.state_dict(), num_zones=15)     # packs zone_net_ once (repack after an update)
    emb = enc(env.obs, env.zone_obs)                        # == model(DictList(obs=..., zone_obs=...))

The first two layers run with bf16 operands and fp32 accumulation (the reference: fp32); the third
layer and ``combine_net_`` are plain ``torch.nn.functional.linear`` calls in fp32.  Inference only -- no autograd graph is built.  There is
no CPU path: the constructor raises without CUDA.
"""
import ctypes

import torch

from . import _lib


class ZoneEncoder:
    def __init__(self, state_dict, num_zones, device='cuda:0'):
        if not torch.cuda.is_available():
            raise RuntimeError('ZoneEncoder needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.num_zones = int(num_zones)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.load_state_dict(state_dict)

    def load_state_dict(self, state_dict):
        """(Re)pack the weights of ``zone_net_`` / keep ``combine_net_`` (after every optimiser update)."""
        f = lambda k: state_dict[k].detach().to(self.device, torch.float32).contiguous()
        w1, b1 = f('zone_net_.0.weight'), f('zone_net_.0.bias')
        w2, b2 = f('zone_net_.2.weight'), f('zone_net_.2.bias')
        self.w3, self.b3 = f('zone_net_.4.weight'), f('zone_net_.4.bias')
        self.combine_w, self.combine_b = f('combine_net_.weight'), f('combine_net_.bias')
        self.hidden = h = w1.shape[0]
        self.obs_dim = self.combine_w.shape[1] - h
        self.zone_dim = w1.shape[1] - self.obs_dim
        assert w2.shape == self.w3.shape == (h, h) and self.zone_dim > 0
        # forward() = combine_net_([obs, W3 pooled + b3]) is ONE affine map of [obs, pooled]: fold W3 into it
        wc = self.combine_w.double()
        self.fold_w = torch.cat([wc[:, :self.obs_dim], wc[:, self.obs_dim:] @ self.w3.double()], dim=1).float().contiguous()
        self.fold_b = (wc[:, self.obs_dim:] @ self.b3.double() + self.combine_b.double()).float().contiguous()
        self.shape = _lib.CrlEncoderShape(obs_dim=self.obs_dim, zone_dim=self.zone_dim, hidden=h,
                                          num_zones=self.num_zones)
        n = ctypes.c_int64()
        _lib.check(self.lib.crl_encoder_packed_bytes(self.shape, ctypes.byref(n)))
        self.packed = torch.zeros(n.value, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.crl_encoder_pack(self.shape, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                                 self.packed.data_ptr(), self._stream()))
        self._keep = (w1, b1, w2, b2)               # alive until the pack kernel has run

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pooled(self, obs, zone_obs, out=None):
        """mean_z relu(L2(relu(L1([obs, zone_obs[:, z]])))): (B, obs_dim) f32, (B, N, Z) f32 -> (B, h) f32
        (the fused kernel)."""
        B = obs.shape[0]
        assert obs.shape == (B, self.obs_dim) and zone_obs.shape == (B, self.num_zones, self.zone_dim)
        assert obs.dtype == zone_obs.dtype == torch.float32 and obs.is_contiguous() and zone_obs.is_contiguous()
        if out is None:
            out = torch.empty(B, self.hidden, dtype=torch.float32, device=self.device)
        assert out.shape == (B, self.hidden) and out.dtype == torch.float32 and out.is_contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.crl_zone_encode(self.shape, B, obs.data_ptr(), zone_obs.data_ptr(),
                                                self.packed.data_ptr(), out.data_ptr(), self._status.data_ptr(),
                                                self._stream()))
        return out

    def zone_embedding(self, obs, zone_obs):
        """mean_z zone_net_([obs, zone_obs[:, z]]) (env_model.py:73) = L3(pooled)."""
        return torch.nn.functional.linear(self.pooled(obs, zone_obs), self.w3, self.b3)

    def healthy(self):
        """False if a tensor-core completion wait ever expired (synchronises)."""
        return int(self._status.item()) == 0

    def __call__(self, obs, zone_obs=None):
        """ZoneEnvModel.forward: accepts the env's obs dict or the two tensors."""
        if zone_obs is None:
            obs, zone_obs = obs['obs'], obs['zone_obs']
        return torch.nn.functional.linear(torch.cat([obs, self.pooled(obs, zone_obs)], dim=-1), self.fold_w, self.fold_b)
