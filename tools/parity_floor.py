#!/usr/bin/env python
"""How far apart are fp32 implementations of the SAME 1000-step sampler, and where does the precision policy sit?

B windows (default 64), T = 120, N = 1000, one Gaussian noise tape shared by every run (identical conditioning and noise).
Runs: (1) the reference op sequence (oracle port) as stock PyTorch eager fp32 on this GPU, TF32 off, whole batch;
(2) the same in two half-batch chunks (other cuBLAS tile shapes => other summation order: the reference against itself);
(3) the same in fp64 (the exact arithmetic, when the port runs in double); (4) the fp32 CUDA-core engine (`simt`);
(5) the tensor-core engine for several policies K.  Every run is compared with (1) and with (3): raw max-abs of the
normalised sample, joint positions (oracle FK, metres -> mm): max, mean, per-window-max percentiles, windows over 1 mm.
usage: python tools/parity_floor.py [B] [K[:R[:S]] ...]  (R = number of dithered fp16 weight sets, EGOEGO_WEIGHT_SETS, empty = default;
                                                          S = how many of the last K steps run the 3-term split, EGOEGO_SPLIT_STEPS: the
                                                          others run fp16 activations x fp16-pair weights)
env PARITY_FLOOR_QUICK=1 skips the fp64 and chunked torch runs; PARITY_FLOOR_WEIGHTS=seed1|seed2|trained_like selects one of the
weight sets of oracle/gen_golden_weightsets.py instead of oracle.init_params(0)."""
import math
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
from helpers import joints  # noqa: E402
import egoego_release_b200 as E  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
def _spec(v):
    f = v.split(":") + ["", ""]
    return int(f[0]), int(f[1]) if f[1] else None, int(f[2]) if f[2] else None


KS = [_spec(v) for v in sys.argv[2:]] or [(k, None, None) for k in (1000, 500, 250, 125, 63, 32)]
QUICK = bool(os.environ.get("PARITY_FLOOR_QUICK"))
N, T = 1000, 120
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(777)
tape = torch.randn(N + 2, B, T, 198, device=dev, generator=g)
xs = synth_x_start(33, B, T).to(dev)
cm = O.prep_head_condition_mask(xs.shape).to(dev)
WS = os.environ.get("PARITY_FLOOR_WEIGHTS", "")
if WS:
    from oracle.gen_golden_weightsets import WEIGHT_SETS  # noqa: E402
    params = O.init_params(**WEIGHT_SETS[WS][0])
    print(f"weights: {WS} {WEIGHT_SETS[WS][0]}", flush=True)
else:
    params = O.init_params(0)


_time_embed32 = O.time_embed


def _time_embed_any(p, t):
    """fp64 runs: the sinusoidal timestep features in the weights' dtype (the port builds them in fp32)."""
    w = p["denoise_fn.time_mlp.1.weight"]
    if w.dtype == torch.float32:
        return _time_embed32(p, t)
    F = torch.nn.functional
    freqs = torch.exp(torch.arange(32, device=t.device, dtype=w.dtype) * -(math.log(10000) / 31))
    ang = t[:, None].to(w.dtype) * freqs[None, :]
    h = F.gelu(F.linear(torch.cat((ang.sin(), ang.cos()), dim=-1), w, p["denoise_fn.time_mlp.1.bias"]))
    return F.linear(h, p["denoise_fn.time_mlp.3.weight"], p["denoise_fn.time_mlp.3.bias"])


O.time_embed = _time_embed_any


def torch_run(dtype, chunks=1):
    p = {k: v.to(dev, dtype) if v.is_floating_point() else v.to(dev) for k, v in params.items()}
    s = {k: v.to(dev, dtype) if v.is_floating_point() else v.to(dev) for k, v in O.make_schedule(N).items()}
    outs = []
    with torch.no_grad():
        for c in torch.arange(B).chunk(chunks):
            c = c.to(dev)
            outs.append(O.p_sample_loop(p, s, xs[c].to(dtype), cm[c].to(dtype), lambda k: tape[k][c].to(dtype)))
    return torch.cat(outs).float()


def engine_run(engine, K, R=None, S=None):
    for name, v in (("EGOEGO_WEIGHT_SETS", R), ("EGOEGO_SPLIT_STEPS", S)):
        if v is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = str(v)
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=N, objective="pred_x0", max_batch=B, engine=engine,
                                precise_last_steps=K)
    m.load_state_dict(params, strict=False)
    m = m.cuda()
    m.set_noise_tape(tape)
    y = m.sample(xs, cm)
    torch.cuda.synchronize()
    del m
    return y


def report(tag, y, refs):
    jy = joints(y)
    msg = f"{tag:28s}"
    for rname, (r, jr) in refs.items():
        if r is None or r is y:
            continue
        d = (jy - jr).abs()
        pw = d.reshape(B, -1).max(1).values.numpy() * 1e3
        msg += (f" | vs {rname}: raw {float((y.cpu() - r.cpu()).abs().max()):.2e} joint max {pw.max():.4f} mean {float(d.mean()) * 1e3:.5f} "
                f"p50/p90 {np.percentile(pw, 50):.4f}/{np.percentile(pw, 90):.4f} mm, >1mm: {int((pw > 1.0).sum())}/{B}")
    print(msg, flush=True)


t32 = torch_run(torch.float32)
print("torch fp32 done", flush=True)
t64 = None
if not QUICK:
    try:
        t64 = torch_run(torch.float64)
    except Exception as ex:  # the port may build fp32 tables internally
        print("fp64 run unavailable:", repr(ex)[:200], flush=True)
refs = {"torch32": (t32, joints(t32)), "torch64": (t64, joints(t64) if t64 is not None else None)}
if not QUICK:
    report("torch fp32 (whole batch)", t32, {"torch64": refs["torch64"]})
    report("torch fp32 (2 chunks)", torch_run(torch.float32, 2), refs)
    report("simt fp32 engine", engine_run("simt", N), refs)
for K, R, S in KS:
    report(f"tcgen05 K={K} sets={R if R is not None else 'default'} split={S if S is not None else 'default'}", engine_run("tcgen05", K, R, S), refs)
