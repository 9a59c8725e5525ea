#!/usr/bin/env python
"""Device time of eval_metrics_kernel at two batch sizes (the slope separates kernel time from per-call host overhead)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egoego_release_b200 as E
from egoego_release_b200 import _capi

dev = torch.device("cuda:0")
T = 120
res = {}
for nseq in (2048, 8192):
    gq = torch.randn(nseq, T, 22, 4, device=dev); gj = torch.randn(nseq, T, 22, 3, device=dev)
    pq = torch.randn(nseq, T, 22, 4, device=dev); pj = gj + 0.01 * torch.randn(nseq, T, 22, 3, device=dev)
    fl = torch.zeros(nseq, device=dev)
    out = torch.empty(nseq, 35, device=dev)
    L = _capi.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    call = lambda: L.egoego_eval_metrics(0, gq.data_ptr(), gj.data_ptr(), fl.data_ptr(), pq.data_ptr(), pj.data_ptr(), fl.data_ptr(), nseq, T, out.data_ptr(), st)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = nseq * T * (22 * 3 + 2 * 4) * 4 * 2
    res[nseq] = ms
    print(f"nseq {nseq}: {ms * 1e3:.1f} us per launch through the C ABI, {by / ms / 1e6:.0f} GB/s")
    del gq, gj, pq, pj
print(f"slope: {(res[8192] - res[2048]) / 6144 * 2048 * 1e3:.1f} us per 2048 sequences")
