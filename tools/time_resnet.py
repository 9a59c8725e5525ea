#!/usr/bin/env python
"""Device time of the ResNet-18 flow encoder (csrc/resnet.cu) on 139 frames."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from oracle import stage1 as S

dev = torch.device("cuda:0")
opt = argparse.Namespace(window=60, n_dec_layers=2, n_head=4, d_k=256, d_v=256, d_model=256, input_of_feats=False, freeze_of_cnn=True, dist_scale=10.0)
m = E.HeadFormer(opt, dev)
m.load_state_dict({**S.init_params(7, S.CFG_HEAD), **{"cnn." + k: v for k, v in S.init_resnet_params(9).items()}})
m = m.to(dev)
flow = torch.randn(1, 139, 224, 224, 2, device=dev)
for _ in range(2):
    m._input_features({"of": flow})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m._input_features({"of": flow})
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"ResNet-18 encoder, 139 frames: {ms:.2f} ms, {139 * 1.814e9 / ms / 1e9:.2f} TFLOP/s")
