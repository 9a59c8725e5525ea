"""Shared helpers of the GPU parity tests: build the product model from the oracle's seeded weights."""
import numpy as np
import torch

from oracle import egoego_oracle as O

ENGINES = ["simt", "tcgen05"]


def make_model(timesteps, engine, params=None, max_batch=8, device="cuda:0"):
    import egoego_release_b200 as E
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, timesteps=timesteps, objective="pred_x0",
                                loss_type="l1", max_batch=max_batch, engine=engine)
    missing, unexpected = m.load_state_dict(params if params is not None else O.init_params(0), strict=False)
    assert not unexpected
    return m.to(device)


def joints(x):
    """[B,T,198] raw sample -> FK joint positions in metres via the ORACLE's post-processing."""
    return O.joints_from_model_output(O.MotionDataStub(), x.detach().cpu().float())


def maxabs(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max())
