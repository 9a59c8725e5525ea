#!/usr/bin/env python
"""In-loop time of the sampling step (CUDA events around whole sample() calls), per precision mode:
    python tools/loop_time.py [B] [N]        (environment switches such as EGOEGO_ZIGZAG=0 apply)
Prints us per diffusion step for the all-fp16 loop, the all-split loop and the default policy, plus windows/s of the latter."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import egoego_oracle as O
from oracle.gen_golden import synth_x_start

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
xs = synth_x_start(1, B, 120).cuda()
cm = O.prep_head_condition_mask(xs.shape).cuda()
params = O.init_params(0)
res = {}
os.environ.setdefault("EGOEGO_SPLIT_STEPS", "")
SPLIT_ENV = os.environ["EGOEGO_SPLIT_STEPS"]
ONLY = os.environ.get("LOOP_ONLY", "")          # e.g. LOOP_ONLY=fp16: time just that mode
for tag, K, n in (("fp16", E.PRECISE_ALL_FP16, N), ("split", 10 ** 6, max(N // 4, 50)), ("pair", 10 ** 6, max(N // 4, 50)), ("default", 0, N)):
    if ONLY and tag not in ONLY.split(","):
        continue
    # "split": every step in the 3-term format; "pair": every step with fp16 activations x fp16-pair weights
    os.environ["EGOEGO_SPLIT_STEPS"] = {"split": "1000000", "pair": "0"}.get(tag, SPLIT_ENV)
    if not os.environ["EGOEGO_SPLIT_STEPS"]:
        del os.environ["EGOEGO_SPLIT_STEPS"]
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=n, objective="pred_x0", max_batch=B, precise_last_steps=K)
    m.load_state_dict(params, strict=False)
    m = m.cuda()
    torch.manual_seed(0)
    m.sample(xs, cm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        y = m.sample(xs, cm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[tag] = ms * 1e3 / n
    extra = f"  {B * 1e3 / ms:.1f} windows/s" if tag == "default" else ""
    print(f"{tag:8s} N={n:5d} B={B}: {ms:9.2f} ms per sample call = {ms * 1e3 / n:8.1f} us per diffusion step{extra}  (finite={bool(torch.isfinite(y).all())})", flush=True)
    del m
    torch.cuda.empty_cache()
