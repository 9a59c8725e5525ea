#!/usr/bin/env python
"""BASELINE config 5 plumbing check (run under torchrun, one rank per GPU): the training step behind torch DDP and with the engine's
own gradient averaging (CondGaussianDiffusion.set_grad_sync: one all-reduce of the flat gradient buffer, no wrapper).
Every rank trains on its own shard; both must leave the MEAN of the per-rank CUDA-backward gradients in .grad."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

import egoego_release_b200 as E
from oracle import egoego_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, T = 8, 120
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121, out_dim=198,
                            timesteps=1000, objective="pred_x0", loss_type="l1", max_batch=B)
m.load_state_dict(O.init_params(0), strict=False)
m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
x0 = torch.rand(B, T, 198, device=dev, generator=g) * 2 - 1
cm = O.prep_head_condition_mask(x0.shape).to(dev)
tt = torch.randint(0, 1000, (B,), device=dev, generator=g)
noise, cnoise = torch.randn(B, T, 198, device=dev, generator=g), torch.randn(B, T, 198, device=dev, generator=g)
DSEED = 4242 + rank          # the module is in train() mode: every call of this check must draw the SAME dropout masks
# local gradients without DDP
m.zero_grad(set_to_none=True)
m.p_losses(x0, cm, tt, noise=noise, cond_noise=cnoise, dropout_seed=DSEED).backward()
local_g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).clone()
mean_g = local_g.clone()
dist.all_reduce(mean_g)
mean_g /= world


class Step(torch.nn.Module):          # DDP calls forward(); route it to p_losses with the fixed draws of this check
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x0, cm, tt, noise, cnoise):
        return self.model.p_losses(x0, cm, tt, noise=noise, cond_noise=cnoise, dropout_seed=DSEED)


ddp = DDP(Step(m), device_ids=[local])
m.zero_grad(set_to_none=True)
loss = ddp(x0, cm, tt, noise, cnoise)
loss.backward()
ddp_g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
err = float((ddp_g - mean_g).abs().max() / mean_g.abs().max())
print(f"rank {rank}/{world}: loss {float(loss.detach()):.6f}, DDP gradient vs mean of per-rank gradients: max rel err {err:.2e}", flush=True)
assert err < 1e-5
del ddp
# the engine's own averaging: same result without the wrapper; no_grad_sync() keeps the gradients local
m.set_grad_sync()
m.zero_grad(set_to_none=True)
m.p_losses(x0, cm, tt, noise=noise, cond_noise=cnoise, dropout_seed=DSEED).backward()
flat_g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
err_f = float((flat_g - mean_g).abs().max() / mean_g.abs().max())
m.zero_grad(set_to_none=True)
with m.no_grad_sync():
    m.p_losses(x0, cm, tt, noise=noise, cond_noise=cnoise, dropout_seed=DSEED).backward()
loc2 = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
err_l = float((loc2 - local_g).abs().max() / local_g.abs().max())
print(f"rank {rank}/{world}: set_grad_sync gradient vs mean: max rel err {err_f:.2e}; no_grad_sync vs local: {err_l:.2e}", flush=True)
assert err_f < 1e-5 and err_l < 1e-5
dist.barrier()
dist.destroy_process_group()
