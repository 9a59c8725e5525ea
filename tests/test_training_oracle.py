"""Training objective: the oracle's p_losses + autograd gradients against loss / gradient fingerprints of
the unmodified reference (eval-mode modules: dropout is identity, see oracle/training.py), and the CUDA training step
(egoego_train_step through the host mirror's p_losses / loss.backward()) against the same goldens."""
import os

import numpy as np
import pytest

from oracle import egoego_oracle as O
from oracle import training as TR
from oracle.gen_golden_training import CASES, case_inputs


@pytest.mark.parametrize("tag,B,T,seed,with_pm", CASES)
def test_training_loss_and_gradients_vs_reference_golden(tag, B, T, seed, with_pm, golden_dir, params0):
    g = dict(np.load(os.path.join(golden_dir, "training.npz")))
    x_start, cm, t, noise, cond_noise, pm = case_inputs(seed, B, T, with_pm)
    loss, grads = TR.loss_and_grads(params0, O.make_schedule(1000), x_start, cm, t, noise, cond_noise, pm)
    assert abs(float(loss) - float(g[f"{tag}_loss"])) < 1e-6
    summ = TR.grad_summary(grads)
    keys = [k[len(tag) + 1:] for k in g if k.startswith(tag + "|")]
    assert set(keys) == set(summ) and len(keys) == 72
    for k in keys:
        ref = g[f"{tag}|{k}"]
        assert np.abs(summ[k].numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), k
    # a padded frame contributes nothing: the loss is linear in the mask
    if with_pm:
        assert float(loss) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,B,T,seed,with_pm", CASES)
def test_gpu_training_step_vs_reference_golden(tag, B, T, seed, with_pm, golden_dir, params0):
    """CondGaussianDiffusion.p_losses on the device (egoego_train_step: fp32 forward + backward kernels) against the loss and
    the gradient fingerprints of the unmodified reference (eval-mode dropout).  Tolerances are fractions of each tensor's
    gradient norm (products run as 3-term bf16 splits on the tensor cores and reduce over B*128 rows in a different order)."""
    import torch
    import egoego_release_b200 as E
    g = dict(np.load(os.path.join(golden_dir, "training.npz")))
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=1000, objective="pred_x0", loss_type="l1", max_batch=4)
    m.load_state_dict(params0, strict=False)
    m = m.cuda()
    x_start, cm, t, noise, cond_noise, pm = case_inputs(seed, B, T, with_pm)
    loss = m.p_losses(x_start.cuda(), cm.cuda(), t.cuda(), noise=noise.cuda(), padding_mask=None if pm is None else pm.cuda(),
                      cond_noise=cond_noise.cuda())
    assert abs(float(loss.detach()) - float(g[f"{tag}_loss"])) < 2e-5, (float(loss.detach()), float(g[f"{tag}_loss"]))
    loss.backward()
    grads = {k: v.grad.detach().cpu() for k, v in m.named_parameters() if v.grad is not None}
    assert len(grads) == 72
    summ = TR.grad_summary(grads)
    worst = (0.0, "")
    for k, v in summ.items():
        ref = g[f"{tag}|{k}"]
        scale = max(abs(ref[0]), abs(ref[1]), 1e-6)
        err = float(np.abs(v.numpy() - ref).max() / scale)
        worst = max(worst, (err, k))
        # the `sum` entry cancels over up to 1.5 M signed elements: 2e-3 of the norm; norm and the eight sampled values: 5e-4
        assert err < 2e-3, (k, err, v.numpy()[:4], ref[:4])
        assert abs(v.numpy()[0] - ref[0]) < 5e-4 * scale and np.abs(v.numpy()[2:] - ref[2:]).max() < 5e-4 * scale, (k, v.numpy()[:4], ref[:4])
    print(f"[{tag}] loss {float(loss.detach()):.6f} vs {float(g[f'{tag}_loss']):.6f}; worst gradient fingerprint error {worst[0]:.2e} of its norm ({worst[1]})")
    # a small gradient-descent step through a stock torch optimizer lowers the loss on the same batch: the gradients point downhill
    # and the engine picks up the updated parameters
    opt = torch.optim.SGD(m.parameters(), lr=1e-3)
    opt.step()
    loss2 = m.p_losses(x_start.cuda(), cm.cuda(), t.cuda(), noise=noise.cuda(), padding_mask=None if pm is None else pm.cuda(),
                       cond_noise=cond_noise.cuda())
    print(f"[{tag}] loss after one SGD step: {float(loss2.detach()):.6f} (before {float(loss.detach()):.6f})")
    assert float(loss2.detach()) < float(loss.detach())
