#!/usr/bin/env python
"""Device time of the ResNet-18 flow encoder (csrc/resnet.cu) on 139 frames: tensor-core engine (default), fp32 CUDA-core engine
(EGOEGO_RESNET=simt) and the reference's ResNet as stock PyTorch / cuDNN (TF32 on = torch's default for convolutions, and off)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import stage1 as S

dev = torch.device("cuda:0")
opt = argparse.Namespace(window=60, n_dec_layers=2, n_head=4, d_k=256, d_v=256, d_model=256, input_of_feats=False, freeze_of_cnn=True, dist_scale=10.0)
flow = torch.randn(1, 139, 224, 224, 2, device=dev)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


outs = {}
engines = ["tcgen05"] if os.environ.get("PROF_ONLY_TC") else ["tcgen05", "simt"]
for eng in engines:
    if eng == "simt":
        os.environ["EGOEGO_RESNET"] = "simt"
    else:
        os.environ.pop("EGOEGO_RESNET", None)
    m = E.HeadFormer(opt, dev)
    m.load_state_dict({**S.init_params(7, S.CFG_HEAD), **{"cnn." + k: v for k, v in S.init_resnet_params(9).items()}})
    m = m.to(dev)
    ms = timed(lambda: m._input_features({"of": flow}))
    outs[eng] = m._input_features({"of": flow}).float().cpu()
    print(f"ResNet-18 encoder [{eng}], 139 frames: {ms:.2f} ms, {139 / ms * 1e3:.0f} frames/s, {139 * 1.814e9 / ms / 1e9:.2f} TFLOP/s", flush=True)
    del m
if len(outs) == 2:
    d = (outs["tcgen05"] - outs["simt"]).abs().max() / outs["simt"].abs().max()
    print(f"tensor-core vs fp32 CUDA-core features: max-abs difference {float(d):.2e} of the feature range")
pr = {k: v.to(dev) for k, v in S.init_resnet_params(9).items()}
xin = S.flow_to_cnn_input(flow.cpu()).to(dev)
for tag, tf32 in (("cuDNN TF32 (torch default)", True), ("cuDNN fp32", False)):
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        ms = timed(lambda: S.resnet18_forward(pr, xin))
        y = S.resnet18_forward(pr, xin).float().cpu()
    extra = ""
    if "simt" in outs:
        extra = f", vs fp32 CUDA-core features {float((y - outs['simt'][0]).abs().max() / outs['simt'].abs().max()):.2e} of range"
    print(f"torch eager {tag}: {ms:.2f} ms, {139 / ms * 1e3:.0f} frames/s{extra}", flush=True)
