"""TEST INFRASTRUCTURE (CPU oracle) -- restatement of the stage-2 TRAINING objective (BASELINE config 5):

  forward / p_losses / q_sample   egoego/model/transformer_cond_diffusion_model.py:557-625
      x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps;  x_cond = x0 (1 - m) + m eps';  model_out = denoise_fn([x_t | x_cond], t, pm)
      loss = mean_b( mean_{t,d}( |model_out - target| * pm[b, t+1] ) * p2_loss_weight[t_b] ),  target = x0 (pred_x0) or eps
  (L1 via F.l1_loss(reduction='none'); `reduce(loss, 'b ... -> b (...)', 'mean')` averages over ALL T*D entries, padded or not.)

The denoiser is oracle.egoego_oracle.denoiser_forward (torch ops, differentiable), so ``loss_and_grads`` gets the gradients of
every parameter from torch autograd -- the same engine the reference trains with.  Dropout (p = 0.1 in MultiHeadAttention and
PositionwiseFeedForward while training) is NOT part of this restatement: torch's dropout stream cannot be reproduced by another
implementation, so parity for the training row is defined with the modules in eval() mode (dropout = identity), which is how
tests/golden/training.npz is generated from the unmodified reference (oracle/gen_golden_training.py).

The CUDA training step (egoego_train_step, csrc/train.cuh) is held to the same goldens: tests/test_training_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import egoego_oracle as O


def q_sample(sched, x_start, t, noise):
    return (sched["sqrt_alphas_cumprod"][t].reshape(-1, 1, 1) * x_start +
            sched["sqrt_one_minus_alphas_cumprod"][t].reshape(-1, 1, 1) * noise)


def p_losses(p: Dict[str, torch.Tensor], sched, x_start, cond_mask, t, noise, cond_noise,
             padding_mask: Optional[torch.Tensor] = None, objective: str = "pred_x0", loss_type: str = "l1") -> torch.Tensor:
    """:574-605.  t int64 [B]; noise / cond_noise [B,T,D]; padding_mask bool [B,1,T+1] or None."""
    x = q_sample(sched, x_start, t, noise)
    x_cond = x_start * (1.0 - cond_mask) + cond_mask * cond_noise
    out = O.denoiser_forward(p, torch.cat((x, x_cond), dim=-1), t, padding_mask)
    target = noise if objective == "pred_noise" else x_start
    fn = F.l1_loss if loss_type == "l1" else F.mse_loss
    loss = fn(out, target, reduction="none")
    if padding_mask is not None:
        loss = loss * padding_mask[:, 0, 1:][:, :, None]
    loss = loss.reshape(loss.shape[0], -1).mean(dim=1)
    loss = loss * sched["p2_loss_weight"][t]
    return loss.mean()


def loss_and_grads(p, sched, x_start, cond_mask, t, noise, cond_noise, padding_mask=None, objective="pred_x0"):
    """Loss and d loss / d parameter for every trainable tensor of the denoiser (the positional table is frozen)."""
    q = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "position_vec" not in k else v) for k, v in p.items()}
    loss = p_losses(q, sched, x_start, cond_mask, t, noise, cond_noise, padding_mask, objective)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in q.items() if v.requires_grad and v.grad is not None}


def grad_summary(grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Compact fingerprint of a gradient set (44 MB of gradients cannot be committed): per tensor L2 norm, sum, first 8 values."""
    out = {}
    for k, g in grads.items():
        f = g.detach().double().reshape(-1)
        out[k] = torch.cat((torch.stack((f.norm(), f.sum())), f[:8] if f.numel() >= 8 else F.pad(f, (0, 8 - f.numel()))))
    return out
