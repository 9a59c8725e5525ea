"""TEST INFRASTRUCTURE (CPU oracle) -- restatement of the stage-2 TRAINING objective (BASELINE config 5):

  forward / p_losses / q_sample   egoego/model/transformer_cond_diffusion_model.py:557-625
      x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps;  x_cond = x0 (1 - m) + m eps';  model_out = denoise_fn([x_t | x_cond], t, pm)
      loss = mean_b( mean_{t,d}( |model_out - target| * pm[b, t+1] ) * p2_loss_weight[t_b] ),  target = x0 (pred_x0) or eps
  (L1 via F.l1_loss(reduction='none'); `reduce(loss, 'b ... -> b (...)', 'mean')` averages over ALL T*D entries, padded or not.)

The denoiser is oracle.egoego_oracle.denoiser_forward (torch ops, differentiable), so ``loss_and_grads`` gets the gradients of
every parameter from torch autograd -- the same engine the reference trains with.

Dropout (nn.Dropout(0.1) on the attention probabilities, the fc output and the FFN output of every DecoderLayer while the module
is in train() mode, egoego/model/transformer_module.py:53,59,84,92,105,113): torch's own dropout stream cannot be reproduced by
another implementation, so the product derives its masks from a counter-based Philox4x32-10 function of (seed, layer, site,
element index) -- include/egoego_b200.h: egoego_train_set_dropout -- which ``dropout_keep`` / ``DropoutMasks`` restate here in
numpy, bit for bit.  tests/golden/training.npz holds the reference's loss / gradient fingerprints in eval() mode (dropout =
identity) AND in train() mode with these masks injected into the reference's own nn.Dropout modules
(oracle/gen_golden_training.py), so both modes are pinned to the unmodified reference.

The CUDA training step (egoego_train_step, csrc/train.cuh) is held to the same goldens: tests/test_training_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import egoego_oracle as O


_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint64 arrays holding 32-bit values (Salmon et al. 2011; csrc/common.cuh: philox4x32_10)."""
    m32 = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(v, np.uint64) & m32 for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & m32, np.uint64(k1) & m32
    for _ in range(10):
        p0 = np.uint64(_PHILOX_M0) * c0
        p1 = np.uint64(_PHILOX_M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m32, lo1, (hi0 ^ c3 ^ k1) & m32, lo0
        k0, k1 = (k0 + np.uint64(_PHILOX_W0)) & m32, (k1 + np.uint64(_PHILOX_W1)) & m32
    return c0, c1, c2, c3


def dropout_keep(seed: int, stream: int, idx, p: float = 0.1):
    """Boolean keep mask of the elements with flat indices ``idx`` (any-shape integer array) of dropout stream ``stream``."""
    idx = np.asarray(idx, np.uint64)
    quad = idx >> np.uint64(2)
    w = philox4x32_10(quad & np.uint64(0xFFFFFFFF), quad >> np.uint64(32), np.uint64(stream), np.uint64(0x44524F50),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (idx & np.uint64(3)).astype(np.int64)
    word = np.choose(sel, w)
    return word < np.uint64(int((1.0 - p) * 4294967296.0))


class DropoutMasks:
    """Dropout hook for oracle.egoego_oracle.denoiser_forward (``apply(layer, site, x)``): factor tensors (0 or 1/(1-p)) in the
    product's index convention -- site 0: ((b H + h) 128 + query) 128 + key, sites 1 / 2: (b 128 + token) 512 + channel."""

    def __init__(self, seed: int, p: float = 0.1):
        self.seed, self.p = int(seed), float(p)
        self.scale = np.float32(1.0 / (1.0 - p))

    def __call__(self, layer, site, shape):
        if site == 0:
            B, H, L, _ = shape
            b, h, q, k = np.meshgrid(np.arange(B), np.arange(H), np.arange(L), np.arange(L), indexing="ij")
            idx = ((b * H + h) * 128 + q) * 128 + k
        else:
            B, L, d = shape
            b, l, c = np.meshgrid(np.arange(B), np.arange(L), np.arange(d), indexing="ij")
            idx = (b * 128 + l) * 512 + c
        keep = dropout_keep(self.seed, 4 * layer + site, idx, self.p)
        return torch.from_numpy(np.where(keep, self.scale, np.float32(0.0)).astype(np.float32))

    def apply(self, layer, site, x):
        return x * self(layer, site, tuple(x.shape)).to(x.device)


class TorchDropout:
    """torch's own nn.Dropout(p) at the same three sites (what the reference runs in train() mode): the stock-PyTorch training
    baseline of bench.py; its masks come from torch's generator and match nothing else."""

    def __init__(self, p: float = 0.1):
        self.p = p

    def apply(self, layer, site, x):
        return F.dropout(x, self.p, training=True)


def q_sample(sched, x_start, t, noise):
    return (sched["sqrt_alphas_cumprod"][t].reshape(-1, 1, 1) * x_start +
            sched["sqrt_one_minus_alphas_cumprod"][t].reshape(-1, 1, 1) * noise)


def p_losses(p: Dict[str, torch.Tensor], sched, x_start, cond_mask, t, noise, cond_noise,
             padding_mask: Optional[torch.Tensor] = None, objective: str = "pred_x0", loss_type: str = "l1", dropout=None,
             l1_sign: Optional[torch.Tensor] = None, resid_out: Optional[list] = None) -> torch.Tensor:
    """:574-605.  t int64 [B]; noise / cond_noise [B,T,D]; padding_mask bool [B,1,T+1] or None; dropout: None (eval mode) or a
    DropoutMasks (train mode).

    Test hooks (not part of the reference): `resid_out` receives the detached residual out - target; `l1_sign` [B,T,D] replaces
    |d| by l1_sign * d wherever it is not NaN.  The L1 gradient is sign(d) / N per output element -- discontinuous at d = 0 -- so two
    fp32-grade implementations whose outputs differ by 1e-5 legitimately disagree on the WHOLE contribution of an element whose
    residual is smaller than that; the GPU gradient test enumerates the signs of those few elements."""
    x = q_sample(sched, x_start, t, noise)
    x_cond = x_start * (1.0 - cond_mask) + cond_mask * cond_noise
    out = O.denoiser_forward(p, torch.cat((x, x_cond), dim=-1), t, padding_mask, dropout=dropout)
    target = noise if objective == "pred_noise" else x_start
    fn = F.l1_loss if loss_type == "l1" else F.mse_loss
    loss = fn(out, target, reduction="none")
    if resid_out is not None:
        resid_out.append((out - target).detach())
    if l1_sign is not None:
        assert loss_type == "l1"
        given = ~torch.isnan(l1_sign)
        loss = torch.where(given, torch.nan_to_num(l1_sign) * (out - target), loss)
    if padding_mask is not None:
        loss = loss * padding_mask[:, 0, 1:][:, :, None]
    loss = loss.reshape(loss.shape[0], -1).mean(dim=1)
    loss = loss * sched["p2_loss_weight"][t]
    return loss.mean()


def loss_and_grads(p, sched, x_start, cond_mask, t, noise, cond_noise, padding_mask=None, objective="pred_x0", dropout=None,
                   loss_type="l1", l1_sign=None, resid_out=None):
    """Loss and d loss / d parameter for every trainable tensor of the denoiser (the positional table is frozen)."""
    q = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "position_vec" not in k else v) for k, v in p.items()}
    loss = p_losses(q, sched, x_start, cond_mask, t, noise, cond_noise, padding_mask, objective, loss_type=loss_type, dropout=dropout,
                    l1_sign=l1_sign, resid_out=resid_out)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in q.items() if v.requires_grad and v.grad is not None}


def grad_summary(grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Compact fingerprint of a gradient set (44 MB of gradients cannot be committed): per tensor L2 norm, sum, first 8 values."""
    out = {}
    for k, g in grads.items():
        f = g.detach().double().reshape(-1)
        out[k] = torch.cat((torch.stack((f.norm(), f.sum())), f[:8] if f.numel() >= 8 else F.pad(f, (0, 8 - f.numel()))))
    return out
