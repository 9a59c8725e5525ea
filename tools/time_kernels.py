#!/usr/bin/env python
"""Per-kernel device times of one sampling step (egoego_time_kernel), both operand formats:
    python tools/time_kernels.py [B] [iters]        (environment switches such as EGOEGO_FUSE_LN=0 apply)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import egoego_oracle as O
from oracle.gen_golden import synth_x_start

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                            out_dim=198, timesteps=80, objective="pred_x0", max_batch=B)
m.load_state_dict(O.init_params(0), strict=False)
m = m.cuda()
xs = synth_x_start(1, B, 120).cuda()
cm = O.prep_head_condition_mask(xs.shape).cuda()
m.sample(xs, cm)
torch.cuda.synchronize()
if os.environ.get("PROF_ONLY"):
    sys.exit(0)
cnt = {n: m.launches_per_step(n) for n in m.KERNELS}
for half in (True, False):
    tot = 0.0
    row = []
    for name in m.KERNELS:
        if cnt[name] == 0:
            continue
        ms = min(m.time_kernel(name, B, 120, half, iters=iters) for _ in range(3))
        tot += cnt[name] * ms
        row.append(f"{name} {ms * 1e3:.1f}")
    print(("fp16 " if half else "split") + f" B={B}: " + "  ".join(row) + f"  | step sum {tot * 1e3:.0f} us")
