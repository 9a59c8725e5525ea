#!/usr/bin/env python
"""Joint-position error of the 1000-step golden sample vs the precision policy K (steps t < K use the 3-term split)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import egoego_oracle as O  # noqa: E402
from oracle.gen_golden import Tape, synth_x_start  # noqa: E402
from helpers import joints, make_model, maxabs  # noqa: E402

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "sample.npz")))
N, B, seed = 1000, 1, 22
xs = synth_x_start(100 + N, B, 120)
cm = O.prep_head_condition_mask(xs.shape)
tp = Tape(seed)
tape = torch.stack([tp.draw(xs.shape) for _ in range(N + 2)]).cuda()
ref = torch.from_numpy(g[f"n{N}_b{B}_seed{seed}"])
jr = joints(ref)
params = O.init_params(0)
for K in [int(v) for v in sys.argv[1:]] or [1000, 250, 125, 100, 50, 20, 0]:
    import egoego_release_b200 as E
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=N, objective="pred_x0", max_batch=2, engine="tcgen05",
                                precise_last_steps=K)
    m.load_state_dict(params, strict=False)
    m = m.cuda()
    m.set_noise_tape(tape)
    y = m.sample(xs.cuda(), cm.cuda())
    print(f"precise_last_steps={K:5d}: raw max-abs {maxabs(y, ref):.3e}  joint max-abs {maxabs(joints(y), jr) * 1e3:.4f} mm", flush=True)
    del m
