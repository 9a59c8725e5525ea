#!/usr/bin/env python
"""Workload for `ncu -k regex:<kernel>` captures of the HBM-bound kernels of the path at the bench size (B = 256, T = 120):
a short sampling call (init_sample, stage_rows, ddpm_update, layernorm512 and the split-format start / out GEMMs all run in it),
then post-processing + FK, canonicalisation and the next-window conditioning on 256 windows.  No timing here."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import egoego_oracle as O
from oracle.gen_golden import synth_head_pose, synth_x_start

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                            out_dim=198, timesteps=int(os.environ.get("PROF_STEPS", 12)), objective="pred_x0", max_batch=B,
                            precise_last_steps=int(os.environ.get("PROF_SPLIT_STEPS", 6)))
m.load_state_dict(O.init_params(0), strict=False)
m = m.cuda()
xs = synth_x_start(1, B, 120).cuda()
cm = O.prep_head_condition_mask(xs.shape).cuda()
y = m.sample(xs, cm)
ds = E.MotionDataStub().bind(m)
for _ in range(3):
    aa, root, head, jpos, gq = m.postprocess(ds, y, None, with_fk=True)
    gq2, gj2 = m.fk_smpl(ds, root.reshape(-1, 3), aa.reshape(-1, 22, 3))
    hp = synth_head_pose(3, B, 120).cuda()
    xs2, rq = m.canonicalize_head(ds, hp[:, :, :3].contiguous(), hp[:, :, 3:].contiguous())
    ip = m._tail_condition(ds, gq[:, -10:].contiguous(), jpos[:, -10:].contiguous())
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float(jpos.abs().max()))
