"""Generate tests/golden/training.npz: loss and gradient fingerprints of the UNMODIFIED reference's training objective
(CondGaussianDiffusion.p_losses, transformer_cond_diffusion_model.py:574-605) with the modules in eval() mode (dropout off:
see oracle/training.py), fixed t, noise tape for both Gaussian draws, with and without a padding mask.  Run HERE only."""
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
    np.savez(os.path.join(ROOT, "tests", "golden", "training.npz"), **out)
    print("wrote tests/golden/training.npz")


if __name__ == "__main__":
    main()
