#!/usr/bin/env python
"""Precision policy at FULL size: B = 256 windows, T = 120, N = 1000 steps, Philox noise (the bench workload).

Compares, on identical conditioning and identical noise streams (same seed, streams keyed by the global window id):
  * the tensor-core engine under several policies K (steps t < K in the 3-term split, earlier steps single-pass fp16),
  * against the same engine with every step in the split format (K = N), all 256 windows,
  * and against the independent fp32 CUDA-core engine (`simt`) on a 32-window shard (shard invariance makes the shard
    comparable with the same windows of the full batch).
Prints the joint-position max-abs / mean error in mm per K (bar: 1 mm).  The oracle is only used for FK (checker).
usage: python tools/policy_fullsize.py [K ...]      (-1 = the default policy)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import egoego_oracle as O  # noqa: E402
from oracle.gen_golden import synth_x_start  # noqa: E402
from helpers import joints  # noqa: E402
import egoego_release_b200 as E  # noqa: E402

N, B, T, SHARD, SEED = 1000, 256, 120, 32, 4242
params = O.init_params(0)
xs = synth_x_start(31, B, T).cuda()
cm = O.prep_head_condition_mask(xs.shape).cuda()


def run(engine, K, n_windows=B):
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=N, objective="pred_x0", max_batch=n_windows, engine=engine,
                                precise_last_steps=K)
    m.load_state_dict(params, strict=False)
    m = m.cuda()
    torch.manual_seed(SEED)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = m.sample(xs[:n_windows], cm[:n_windows])
    e1.record()
    torch.cuda.synchronize()
    k_eff = m.precise_last_steps() if engine == "tcgen05" else N
    del m
    return y, k_eff, e0.elapsed_time(e1)


def err_mm(a, b):
    d = (joints(a) - joints(b)).abs()
    return float(d.max()) * 1e3, float(d.mean()) * 1e3


y_split, _, ms = run("tcgen05", N)
print(f"tcgen05 all-split (K={N}): {ms:.0f} ms", flush=True)
y_simt, _, ms = run("simt", N, SHARD)
print(f"simt fp32 engine, windows 0..{SHARD - 1}: {ms:.0f} ms", flush=True)
mx, mean = err_mm(y_split[:SHARD], y_simt)
print(f"all-split vs simt ({SHARD} windows): joint max-abs {mx:.4f} mm, mean {mean:.5f} mm", flush=True)
for K in [int(v) for v in sys.argv[1:]] or [-1, 32, 16, 0]:
    y, k_eff, ms = run("tcgen05", K)
    a = err_mm(y, y_split)
    b = err_mm(y[:SHARD], y_simt)
    print(f"K={k_eff:4d}: {ms:7.0f} ms ({B / ms * 1e3:6.1f} windows/s)  vs all-split ({B} windows): max {a[0]:.4f} mm mean {a[1]:.5f} mm"
          f"   vs simt fp32 ({SHARD} windows): max {b[0]:.4f} mm mean {b[1]:.5f} mm", flush=True)
