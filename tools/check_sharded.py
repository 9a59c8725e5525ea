#!/usr/bin/env python
"""torchrun check of the N>1 path on real GPUs: sharded sampling + one NCCL all-gather == single-GPU run, bit for bit.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/check_sharded.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
import egoego_release_b200 as E  # noqa: E402
from egoego_release_b200.parallel import model_sample_fn, sample_sharded  # noqa: E402
from oracle import egoego_oracle as O  # noqa: E402
from oracle.gen_golden import synth_x_start  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, B = 16, 8 * world
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121, out_dim=198,
                            timesteps=N, objective="pred_x0", max_batch=B)
m.load_state_dict(O.init_params(0), strict=False)
m = m.to(dev)
xs = synth_x_start(3, B, 120).to(dev)
cm = O.prep_head_condition_mask(xs.shape).to(dev)
torch.manual_seed(7)                      # same seed on every rank -> same Philox key
y = sample_sharded(model_sample_fn(m), xs, cm)
torch.manual_seed(7)
y1 = m.sample(xs, cm)                     # the whole batch on this GPU
ok = torch.equal(y, y1)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded({world} GPUs) == single-GPU: {bool(t.item())}  max|diff| {float((y - y1).abs().max()):.3e}")
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
