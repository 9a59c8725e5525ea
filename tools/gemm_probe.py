#!/usr/bin/env python
"""Run the tensor-core GEMM self test for one shape (for ncu captures): gemm_probe.py M N K two_cta half_fmt"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egoego_release_b200 import _capi  # noqa: E402

M, N, K, two, half = (int(v) for v in sys.argv[1:6])
L = _capi.lib()
err, ref, ms = C.c_float(), C.c_float(), C.c_float()
_capi.check(L.egoego_selftest_gemm(0, M, N, K, 42, two, half, C.byref(err), C.byref(ref), C.byref(ms)))
tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
print(f"M={M} N={N} K={K} two_cta={two} half={half}: err {err.value:.3e} ref {ref.value:.3f} {ms.value*1e3:.1f} us {tf:.1f} alg TFLOP/s")
