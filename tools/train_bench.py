#!/usr/bin/env python
"""BASELINE configs[4]: training step of the conditional diffusion (forward + L1 + backward + optimizer) on synthetic AMASS-shaped
batches, batch 32 per GPU (scripts/train_full_body_cond_diffusion.sh:4), 1 GPU or N GPUs under torch DDP:

    python tools/train_bench.py [--steps 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/train_bench.py

Two arms, both the loop body of the reference Trainer.train() (trainer_amass_cond_motion_diffusion.py:124-179: autocast(fp16) +
GradScaler + Adam(lr 1e-4), train() mode, i.e. with dropout):
  ours   egoego_release_b200.CondGaussianDiffusion.forward -> egoego_train_step (CUDA forward + backward kernels, Philox dropout)
  torch  the reference's op sequence (oracle port) as stock PyTorch under autocast(fp16) with nn.Dropout semantics -- the
         reference's real training path on this GPU (the reference itself cannot travel to the GPU box)
Time = CUDA events around `steps` optimizer steps after warm-up, max over ranks; prints one JSON line on rank 0.  With N > 1 the
same measurement WITHOUT gradient synchronisation (DDP no_sync) gives the exposed all-reduce share (10.97 M fp32 gradients = 44 MB)."""
import argparse
import contextlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

import egoego_release_b200 as E
from oracle import egoego_oracle as O
from oracle import training as TR


def run_arms(dev, B=32, T=120, steps=20, warmup=5, rank=0, world=1, local=0, arms=("ours", "torch"), sync_mode="flat"):
    """Time the two arms on this rank's GPU; returns {arm: {ms_per_step, samples_per_s, ...}} (max over ranks when world > 1)."""
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x0 = torch.rand(B, T, 198, device=dev, generator=g) * 2 - 1                       # motion ~ U(-1, 1)
    cm = O.prep_head_condition_mask(x0.shape).to(dev)
    seq_len = torch.randint(30, T + 1, (B, 1), device=dev, generator=g)
    pm = (torch.arange(T + 1, device=dev)[None, :] < seq_len + 1)[:, None, :]

    class OursStep(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                                 out_dim=198, timesteps=1000, objective="pred_x0", loss_type="l1", max_batch=B)
            self.model.load_state_dict(O.init_params(0), strict=False)

        def forward(self, x0, cm, pm):
            return self.model(x0, cm, padding_mask=pm)

    class TorchStep(torch.nn.Module):
        """Reference op sequence (oracle port) with autograd: parameters in a ParameterList, schedule buffers beside it."""

        def __init__(self):
            super().__init__()
            p = O.init_params(0)
            self.keys = list(p)
            self.params = torch.nn.ParameterList([torch.nn.Parameter(v.clone(), requires_grad=v.is_floating_point() and "position_vec" not in k)
                                                  for k, v in p.items()])
            self.sched = {k: v.to(dev) for k, v in O.make_schedule(1000).items()}
            self.drop = TR.TorchDropout(0.1)

        def forward(self, x0, cm, pm):
            p = dict(zip(self.keys, self.params))
            t = torch.randint(0, 1000, (x0.shape[0],), device=x0.device)
            return TR.p_losses(p, self.sched, x0, cm, t, torch.randn_like(x0), torch.randn_like(x0), pm, dropout=self.drop)

    def run(arm):
        mod = (OursStep() if arm == "ours" else TorchStep()).to(dev).train()
        flat = arm == "ours" and world > 1 and sync_mode == "flat"
        if flat:                                   # the engine's own gradient averaging: one all-reduce of the flat buffer, no DDP wrapper
            mod.model.set_grad_sync()
            net = mod
        else:
            net = DDP(mod, device_ids=[local]) if world > 1 else mod
        opt = torch.optim.Adam([p for p in mod.parameters() if p.requires_grad], lr=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=True)

        def step(sync=True):
            opt.zero_grad(set_to_none=True)
            ctx = (mod.model.no_grad_sync() if flat else net.no_sync()) if (world > 1 and not sync) else contextlib.nullcontext()
            with ctx:
                with torch.autocast("cuda", dtype=torch.float16, enabled=True):
                    loss = net(x0, cm, pm)
                scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            return loss

        def timed(sync):
            for _ in range(warmup):
                step(sync)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step(sync)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), float(loss.detach())

        ms, loss = timed(True)
        out = {"ms_per_step": ms, "samples_per_s": world * B / (ms * 1e-3), "last_loss": loss}
        if world > 1:
            out["grad_sync"] = "set_grad_sync (one flat all-reduce)" if flat else "torch DDP"
        if world > 1:
            ms_ns, _ = timed(False)
            out["ms_per_step_no_grad_sync"] = ms_ns
            out["exposed_allreduce_share"] = max(0.0, 1.0 - ms_ns / ms)
        del net, mod, opt
        torch.cuda.empty_cache()
        return out

    res = {"workload": f"configs[4]: training step, batch {B}/GPU x {world} GPU(s), T={T}, autocast(fp16)+GradScaler+Adam, train() mode (dropout 0.1)",
           "n_gpus": world, "steps": steps}
    names = {"ours": "ours", "torch": "torch_autocast_fp16"}
    for arm in arms:
        res[names[arm]] = run(arm)
    if "ours" in res and "torch_autocast_fp16" in res:
        res["speedup_vs_torch_autocast"] = res["torch_autocast_fp16"]["ms_per_step"] / res["ours"]["ms_per_step"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--arms", default="ours,torch", help="comma-separated: ours, torch")
    ap.add_argument("--sync", default="flat", choices=("flat", "ddp"),
                    help="N > 1, our arm: flat = CondGaussianDiffusion.set_grad_sync (one all-reduce), ddp = torch DDP wrapper; the torch arm always uses DDP")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_arms(dev, a.batch, 120, a.steps, a.warmup, rank, world, local, arms=tuple(a.arms.split(",")), sync_mode=a.sync)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
