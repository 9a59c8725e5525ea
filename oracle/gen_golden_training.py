"""Generate tests/golden/training.npz: loss and gradient fingerprints of the UNMODIFIED reference's training objective
(CondGaussianDiffusion.p_losses, transformer_cond_diffusion_model.py:574-605), fixed t, noise tape for both Gaussian draws, with
and without a padding mask, in two modes:
  * eval()  -- dropout off (cases ``pm``, ``nopm``);
  * train() -- the reference's twelve nn.Dropout(0.1) modules (attn_dropout / dropout of every MultiHeadAttention, dropout of every
    PositionwiseFeedForward, transformer_module.py:53,59,105) keep their place in the graph but draw their masks from the
    product's counter-based function (oracle/training.py: DropoutMasks) instead of torch's generator (cases ``pm_drop``,
    ``nopm_drop``): everything else -- the module code around them, autograd -- is the reference's own.
Run HERE only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import egoego_oracle as O  # noqa: E402
from oracle import training as TR  # noqa: E402
from oracle.gen_golden import Tape, build_model, import_reference  # noqa: E402

CASES = [("pm", 3, 120, 81, True), ("nopm", 2, 30, 82, False)]
DROP_CASES = [("pm_drop", 3, 120, 81, True, 1001), ("nopm_drop", 2, 30, 82, False, 1002)]     # (..., dropout seed)


class InjectedDropout(torch.nn.Module):
    """Stands in for one nn.Dropout(0.1) of the reference: multiplies by the product's mask for (layer, site)."""

    def __init__(self, masks, layer, site, n_head):
        super().__init__()
        self.masks, self.layer, self.site, self.n_head = masks, layer, site, n_head

    def forward(self, x):
        if self.site == 0:                       # reference layout [(n_head * bs), n_q, n_k], head-major (transformer_module.py:71-76)
            hb, L, _ = x.shape
            B = hb // self.n_head
            f = self.masks(self.layer, 0, (B, self.n_head, L, L)).permute(1, 0, 2, 3).reshape(hb, L, L)
        else:
            f = self.masks(self.layer, self.site, tuple(x.shape))
        return x * f


def inject_dropout(model, masks):
    for l, layer in enumerate(model.denoise_fn.motion_transformer.layer_stack):
        assert isinstance(layer.self_attn.attn_dropout, torch.nn.Dropout) and layer.self_attn.attn_dropout.p == 0.1
        assert isinstance(layer.self_attn.dropout, torch.nn.Dropout) and isinstance(layer.pos_ffn.dropout, torch.nn.Dropout)
        layer.self_attn.attn_dropout = InjectedDropout(masks, l, 0, layer.self_attn.n_head)
        layer.self_attn.dropout = InjectedDropout(masks, l, 1, layer.self_attn.n_head)
        layer.pos_ffn.dropout = InjectedDropout(masks, l, 2, layer.self_attn.n_head)


def case_inputs(seed, B, T, with_pm):
    tp = Tape(seed)
    x_start = tp.draw((B, T, 198)).clamp(-1, 1)
    noise, cond_noise = tp.draw((B, T, 198)), tp.draw((B, T, 198))
    t = torch.tensor([(seed * 37 + 211 * b) % 1000 for b in range(B)], dtype=torch.long)
    cm = O.prep_head_condition_mask(x_start.shape)
    pm = None
    if with_pm:
        seq_len = torch.tensor([T, T // 2, 31][:B])
        pm = (torch.arange(T + 1)[None, :] < (seq_len + 1)[:, None])[:, None, :]
    return x_start, cm, t, noise, cond_noise, pm


def main():
    M = import_reference()
    params = O.init_params(seed=0)
    m = build_model(M, params, 1000).eval()
    out = {}
    for tag, B, T, seed, with_pm in CASES:
        x_start, cm, t, noise, cond_noise, pm = case_inputs(seed, B, T, with_pm)
        m.zero_grad()
        tape = Tape(0)
        tape.draw = lambda shape, _c=cond_noise: _c            # p_losses draws x_cond's noise with torch.randn_like
        with tape:
            loss = m.p_losses(x_start, cm, t, noise=noise, padding_mask=pm)
        loss.backward()
        grads = {k: v.grad for k, v in m.named_parameters() if v.grad is not None}
        summ = TR.grad_summary(grads)
        out[f"{tag}_loss"] = np.array(float(loss))
        for k, v in summ.items():
            out[f"{tag}|{k}"] = v.numpy()
        l2, g2 = TR.loss_and_grads(params, O.make_schedule(1000), x_start, cm, t, noise, cond_noise, pm)
        s2 = TR.grad_summary(g2)
        worst = max(float((s2[k] - summ[k]).abs().max()) for k in summ)       # absolute: some gradients are exactly ~0 (w_k.bias)
        print(f"{tag}: loss {float(loss):.6f} (restatement {float(l2):.6f}), {len(summ)} gradient tensors, worst absolute fingerprint diff {worst:.2e}")
        assert set(s2) == set(summ)
    for tag, B, T, seed, with_pm, dseed in DROP_CASES:
        x_start, cm, t, noise, cond_noise, pm = case_inputs(seed, B, T, with_pm)
        masks = TR.DropoutMasks(dseed, 0.1)
        mt = build_model(M, params, 1000).train()
        inject_dropout(mt, masks)
        tape = Tape(0)
        tape.draw = lambda shape, _c=cond_noise: _c
        with tape:
            loss = mt.p_losses(x_start, cm, t, noise=noise, padding_mask=pm)
        loss.backward()
        grads = {k: v.grad for k, v in mt.named_parameters() if v.grad is not None}
        summ = TR.grad_summary(grads)
        out[f"{tag}_loss"] = np.array(float(loss))
        for k, v in summ.items():
            out[f"{tag}|{k}"] = v.numpy()
        l2, g2 = TR.loss_and_grads(params, O.make_schedule(1000), x_start, cm, t, noise, cond_noise, pm, dropout=masks)
        s2 = TR.grad_summary(g2)
        worst = max(float((s2[k] - summ[k]).abs().max()) for k in summ)
        print(f"{tag}: loss {float(loss):.6f} (restatement {float(l2):.6f}; eval-mode loss {float(out[tag[:-5] + '_loss']):.6f}), "
              f"{len(summ)} gradient tensors, worst absolute fingerprint diff {worst:.2e}")
    np.savez(os.path.join(ROOT, "tests", "golden", "training.npz"), **out)
    print("wrote tests/golden/training.npz")


if __name__ == "__main__":
    main()
