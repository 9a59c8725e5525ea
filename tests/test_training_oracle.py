"""Training objective: the oracle's p_losses + autograd gradients against loss / gradient fingerprints of
the unmodified reference (eval-mode modules: dropout is identity, see oracle/training.py).  Oracle only -- the CUDA training
step is not built; these goldens are what it will be held to."""
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
