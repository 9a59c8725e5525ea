#!/usr/bin/env python
"""Which rounding of the single-pass fp16 steps reaches the final sample?  (emulation in PyTorch fp32 on the GPU)

The reference op sequence (oracle port) runs 1000 steps on B windows with one shared noise tape.  For steps t >= K the
operands of every product (projections, Q K^T, P V) are rounded to fp16 before an fp32 product -- what FMT_HALF does on
the tensor cores -- in four variants: weights only, activations only, both, both with per-step dithered weight rounding
(RN(W + u * ulp16(W)), u ~ U(-1/2, 1/2) per step, so the weight error is zero-mean over steps instead of a fixed bias).
Each variant is compared with the unrounded fp32 run: joint positions (oracle FK), max / mean / per-window percentiles.
usage: python tools/rounding_study.py [B] [K]"""
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import egoego_oracle as O  # noqa: E402
from oracle.gen_golden import synth_x_start  # noqa: E402
from helpers import joints  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 63
N, T = 1000, 120
dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
if dev.type == "cpu":
    N = int(os.environ.get("STUDY_STEPS", 8))
g = torch.Generator(device=dev)
g.manual_seed(777)
tape = torch.randn(N + 2, B, T, 198, device=dev, generator=g)
xs = synth_x_start(33, B, T).to(dev)
cm = O.prep_head_condition_mask(xs.shape).to(dev)
params = {k: v.to(dev) for k, v in O.init_params(0).items()}
sched = {k: v.to(dev) for k, v in O.make_schedule(N).items()}

MODE = {"w": False, "x": False, "dither": False, "on": False, "u": 0.0}
rng = random.Random(5)


def r16(x):
    return x.half().float()


def rw(w):
    if not MODE["dither"]:
        return r16(w)
    _, e = torch.frexp(w)                                   # |w| = m 2^e, m in [0.5, 1): fp16 ulp = 2^(e-11), subnormal 2^-24
    ulp = torch.ldexp(torch.ones_like(w), (e - 11).clamp(min=-24))
    return r16(w + MODE["u"] * ulp)


class FShim:
    def __getattr__(self, name):
        return getattr(TF, name)

    @staticmethod
    def linear(x, w, b=None):
        if MODE["on"]:
            if MODE["x"]:
                x = r16(x)
            if MODE["w"]:
                w = rw(w)
        return TF.linear(x, w, b)


class TorchShim:
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def matmul(a, b):
        if MODE["on"] and MODE["x"]:
            a, b = r16(a), r16(b)
        return torch.matmul(a, b)


O.F, O.torch = FShim(), TorchShim()
_time_embed, _p_sample = O.time_embed, O.p_sample


def time_embed(p, t):                                       # the engine's timestep-embedding table is fp32
    on, MODE["on"] = MODE["on"], False
    try:
        return _time_embed(p, t)
    finally:
        MODE["on"] = on


def p_sample(p, s, x, t, *a, **kw):
    MODE["on"] = int(t) >= K and (MODE["w"] or MODE["x"])
    MODE["u"] = rng.random() - 0.5
    return _p_sample(p, s, x, t, *a, **kw)


O.time_embed, O.p_sample = time_embed, p_sample


def run(w, x, dither=False):
    MODE.update(w=w, x=x, dither=dither)
    with torch.no_grad():
        return O.p_sample_loop(params, sched, xs, cm, lambda k: tape[k]).float()


ref = run(False, False)
jr = joints(ref)
print(f"B={B} N={N} K={K}: fp32 reference run done", flush=True)
for tag, a in (("weights fp16 (t>=K)", (True, False)), ("activations fp16 (t>=K)", (False, True)), ("both fp16 (t>=K)", (True, True)),
               ("both, dithered weights", (True, True, True)), ("weights only, dithered", (True, False, True))):
    y = run(*a)
    d = (joints(y) - jr).abs()
    pw = d.reshape(B, -1).max(1).values.numpy() * 1e3
    print(f"{tag:26s}: raw {float((y - ref).abs().max()):.2e}  joint max {pw.max():.4f} mean {float(d.mean()) * 1e3:.5f} "
          f"p50/p90 {np.percentile(pw, 50):.4f}/{np.percentile(pw, 90):.4f} mm, >1mm: {int((pw > 1.0).sum())}/{B}", flush=True)
