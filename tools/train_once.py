#!/usr/bin/env python
"""Three training steps (B = 32) through the host mirror, for an ncu launch list of the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import egoego_oracle as O

dev = torch.device("cuda:0")
B, T = 32, 120
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121, out_dim=198,
                            timesteps=1000, objective="pred_x0", loss_type="l1", max_batch=B)
m.load_state_dict(O.init_params(0), strict=False)
m = m.to(dev)
x0 = torch.rand(B, T, 198, device=dev) * 2 - 1
cm = O.prep_head_condition_mask(x0.shape).to(dev)
pm = (torch.arange(T + 1, device=dev)[None, :] < torch.randint(31, T + 2, (B, 1), device=dev))[:, None, :]
tt = torch.randint(0, 1000, (B,), device=dev)
opt = torch.optim.SGD(m.parameters(), lr=1e-6)
for i in range(3):
    opt.zero_grad(set_to_none=True)
    loss = m.p_losses(x0, cm, tt, padding_mask=pm)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
