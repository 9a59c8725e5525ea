#!/usr/bin/env python
"""Which rounding of the single-pass fp16 steps reaches the final sample?  (emulation in PyTorch fp32, GPU if present)

The reference op sequence (oracle port) runs N = 1000 steps on B windows with one shared noise tape.  For steps t >= K the
operands of every product (projections, Q K^T, P V) are rounded to fp16 before an fp32 product -- what FMT_HALF does on the
tensor cores (oracle/rounding.py) -- in several variants: weights only, activations only, both, and both with the engine's R
dithered weight copies cycled over the steps.  Each variant is compared with the unrounded fp32 run: joint positions
(oracle FK), max / mean / per-window percentiles.  profiles/r1bd_rounding_study.txt is an earlier version of this study
(random instead of stratified dither).
usage: python tools/rounding_study.py [B] [K] [R ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import egoego_oracle as O  # noqa: E402
from oracle.gen_golden import synth_x_start  # noqa: E402
from oracle.rounding import emulate_fp16_steps  # noqa: E402
from helpers import joints  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 63
RS = [int(v) for v in sys.argv[3:]] or [8, 16]
N, T = 1000, 120
dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
if dev.type == "cpu":
    N = int(os.environ.get("STUDY_STEPS", 100))
g = torch.Generator(device=dev)
g.manual_seed(777)
tape = torch.randn(N + 2, B, T, 198, device=dev, generator=g)
xs = synth_x_start(33, B, T).to(dev)
cm = O.prep_head_condition_mask(xs.shape).to(dev)
params = {k: v.to(dev) for k, v in O.init_params(0).items()}
sched = {k: v.to(dev) for k, v in O.make_schedule(N).items()}


def run(**kw):
    with torch.no_grad():
        if not kw:
            return O.p_sample_loop(params, sched, xs, cm, lambda k: tape[k]).float().cpu()
        with emulate_fp16_steps(K, **kw):
            return O.p_sample_loop(params, sched, xs, cm, lambda k: tape[k]).float().cpu()


ref = run()
jr = joints(ref)
print(f"B={B} N={N} K={K}: fp32 reference run done", flush=True)
variants = [("weights fp16 (t>=K)", dict(weights=True, activations=False)), ("activations fp16 (t>=K)", dict(weights=False, activations=True)),
            ("both fp16 (t>=K)", dict(weights=True, activations=True))]
variants += [(f"both, {R} dithered copies", dict(weights=True, activations=True, sets=R)) for R in RS]
for tag, kw in variants:
    y = run(**kw)
    d = (joints(y) - jr).abs()
    pw = d.reshape(B, -1).max(1).values.numpy() * 1e3
    print(f"{tag:28s}: raw {float((y - ref).abs().max()):.2e}  joint max {pw.max():.4f} mean {float(d.mean()) * 1e3:.5f} "
          f"p50/p90 {np.percentile(pw, 50):.4f}/{np.percentile(pw, 90):.4f} mm, >1mm: {int((pw > 1.0).sum())}/{B}", flush=True)
