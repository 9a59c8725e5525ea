#!/usr/bin/env python
"""Precision policy margin per weight set: joint error (mm) of the tensor-core engine against the 1000-step goldens of the
UNMODIFIED reference (tests/golden/sample_ws_*.npz, sample_extra.npz) for several K = precise_last_steps.
    python tools/precise_scan_ws.py [K ...]        (default 32 48 63 125)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import egoego_release_b200 as E
from oracle import egoego_oracle as O
from oracle.gen_golden import Tape, synth_x_start
from oracle.gen_golden_weightsets import WEIGHT_SETS, N, B, T

Ks = [int(a) for a in sys.argv[1:]] or [32, 48, 63, 125]
ds = O.MotionDataStub()
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
cases = [(name, kw, cseed, tseed, B, f"sample_ws_{name}.npz", f"n{N}_b{B}") for name, (kw, cseed, tseed) in WEIGHT_SETS.items()]
cases.append(("seed0_extra", dict(seed=0), 2100, 23, 4, "sample_extra.npz", "n1000_b4_seed23"))
for name, kw, cseed, tseed, nb, fn, key in cases:
    ref = torch.from_numpy(np.load(os.path.join(G, fn))[key])
    jr = O.joints_from_model_output(ds, ref)
    params = O.init_params(**kw)
    xs = synth_x_start(cseed, nb, T)
    cm = O.prep_head_condition_mask(xs.shape)
    tp = Tape(tseed)
    tape = torch.stack([tp.draw(xs.shape) for _ in range(N + 2)]).cuda()
    row = []
    for K in Ks:
        m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                    out_dim=198, timesteps=N, objective="pred_x0", max_batch=nb, precise_last_steps=K)
        m.load_state_dict(params, strict=False)
        m = m.cuda()
        m.set_noise_tape(tape)
        y = m.sample(xs.cuda(), cm.cuda()).cpu()
        jy = O.joints_from_model_output(ds, y)
        row.append(f"K={K}: max {float((jy - jr).abs().max()) * 1e3:.4f} mm, mean {float((jy - jr).norm(dim=-1).mean()) * 1e3:.4f} mm")
        del m
    print(f"{name:14s} " + " | ".join(row), flush=True)
