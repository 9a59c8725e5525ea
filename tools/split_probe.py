#!/usr/bin/env python
"""Error of the 3-term split GEMM primitive against the fp32 CUDA-core GEMM, single vs dual accumulator (EGOEGO_SPLIT_DUAL),
for the shapes of the denoiser: python tools/split_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egoego_release_b200 import _capi  # noqa: E402

L = _capi.lib()
for (M, N, K) in ((1024, 3072, 512), (4096, 512, 1024), (4096, 512, 512), (4096, 256, 256), (32768, 3072, 512)):
    row = []
    for dual in ("0", "1"):
        os.environ["EGOEGO_SPLIT_DUAL"] = dual
        err, ref, ms = C.c_float(), C.c_float(), C.c_float()
        _capi.check(L.egoego_selftest_gemm(0, M, N, K, 42, 1, 0, C.byref(err), C.byref(ref), C.byref(ms)))
        row.append(f"{'dual' if dual == '1' else 'single'}: max|err| {err.value:.3e} ({err.value / ref.value:.2e} of max|ref|), {ms.value * 1e3:.1f} us")
    print(f"split GEMM {M}x{N}x{K}: " + " | ".join(row), flush=True)
os.environ.pop("EGOEGO_SPLIT_DUAL", None)
