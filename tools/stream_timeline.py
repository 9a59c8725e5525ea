#!/usr/bin/env python
"""Per-CTA timeline of the streamed QKV projection / attention pair (egoego_debug_timeline): runs a short all-fp16 loop at B windows
and prints, for the LAST launch of each kernel, when its CTAs started and ended relative to the first projection CTA.
    python tools/stream_timeline.py [B] [N]      (EGOEGO_STREAM_ATT_PAIRS / _RATIO / EGOEGO_STREAM_ATT apply)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import egoego_release_b200 as E
from egoego_release_b200 import _capi
from oracle import egoego_oracle as O
from oracle.gen_golden import synth_x_start

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
xs = synth_x_start(1, B, 120).cuda()
cm = O.prep_head_condition_mask(xs.shape).cuda()
m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                            out_dim=198, timesteps=N, objective="pred_x0", max_batch=B, precise_last_steps=E.PRECISE_ALL_FP16)
m.load_state_dict(O.init_params(0), strict=False)
m = m.cuda()
torch.manual_seed(0)
m.sample(xs, cm)
torch.cuda.synchronize()
lib = _capi.lib()
assert lib.egoego_debug_timeline(0, 1, None, 0) == 0
m.sample(xs, cm)
buf = np.zeros((2, 160, 2), dtype=np.uint64)
assert lib.egoego_debug_timeline(0, 0, buf.ctypes.data_as(C.c_void_p), buf.size) == 0
print(m.engine_info())
t0 = min(int(v) for v in buf[:, :, 0].reshape(-1) if v)
for k, name in ((0, "qkv projection"), (1, "attention")):
    st = np.array([int(v) - t0 for v in buf[k, :, 0] if v], dtype=np.float64) / 1e3
    en = np.array([int(v) - t0 for v in buf[k, :, 1] if v], dtype=np.float64) / 1e3
    if len(st) == 0:
        print(f"{name}: no CTA recorded (serial mode records the attention kernel only)")
        continue
    print(f"{name}: {len(st)} CTAs started, {len(en)} ended; start min/median/max {st.min():.1f}/{np.median(st):.1f}/{st.max():.1f} us, "
          f"end min/median/max {en.min():.1f}/{np.median(en):.1f}/{en.max():.1f} us")
    if k == 1:
        ctas = [i for i in range(160) if buf[1, i, 0]]
        print("  attention CTAs in blockIdx order (cta: start -> end us): " +
              ", ".join(f"{c}: {(int(buf[1, c, 0]) - t0) / 1e3:.0f}->{(int(buf[1, c, 1]) - t0) / 1e3:.0f}" for c in ctas[::4]))
