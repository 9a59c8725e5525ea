"""TEST INFRASTRUCTURE (never imported by the product): emulation of the tensor-core engine's single-pass fp16 steps inside
the CPU/PyTorch restatement (oracle/egoego_oracle.py), used to study WHICH rounding reaches the final sample.

For steps t >= K the operands of every product of the denoiser (projections, Q K^T, P V) are rounded to fp16 before an fp32
product -- what FMT_HALF does (egoego_release_b200/csrc/gemm_tcgen05.cuh) -- selectively for weights and/or activations, with
the weight rounding either plain round-to-nearest or the engine's dithered copies (engine_tc.cu: dither_offset / dither_round;
step i of the loop uses copy i mod R).  The timestep-embedding MLP stays fp32 (the engine precomputes that table in fp32).

    with emulate_fp16_steps(K, weights=True, activations=True, sets=8):
        y = O.p_sample_loop(params, sched, xs, cm, tape)
"""
from contextlib import contextmanager

import torch
import torch.nn.functional as TF

from oracle import egoego_oracle as O


def dither_offset(r: int, R: int) -> float:
    """engine_tc.cu dither_offset: u_r = (bitrev(r) + 1/2) / R - 1/2 for power-of-two R (0 for R = 1)."""
    if R <= 1:
        return 0.0
    rev = r
    if R & (R - 1) == 0:
        rev, b, x = 0, 1, r
        while b < R:
            rev = (rev << 1) | (x & 1)
            b <<= 1
            x >>= 1
    return (rev + 0.5) / R - 0.5


def round_f16(x: torch.Tensor) -> torch.Tensor:
    return x.half().to(x.dtype)


def dither_round(w: torch.Tensor, u: float) -> torch.Tensor:
    """engine_tc.cu dither_round: RN_fp16(w + u ulp16(w)), ulp16 = 2^(e-11) for |w| = m 2^e (m in [0.5, 1)), 2^-24 once subnormal."""
    if u == 0.0:
        return round_f16(w)
    _, e = torch.frexp(w)
    ulp = torch.ldexp(torch.ones_like(w), (e - 11).clamp(min=-24))
    return round_f16(w + u * ulp)


@contextmanager
def emulate_fp16_steps(K: int, weights: bool = True, activations: bool = True, sets: int = 1):
    """Patch the oracle module so that p_sample calls with t >= K round their product operands to fp16."""
    state = {"on": False, "u": 0.0, "step": 0}
    cache = {}                                            # (weight storage, offset) -> rounded copy (the engine keeps R copies too)

    class FShim:
        def __getattr__(self, name):
            return getattr(TF, name)

        @staticmethod
        def linear(x, w, b=None):
            if state["on"]:
                if activations:
                    x = round_f16(x)
                if weights:
                    key = (w.data_ptr(), tuple(w.shape), state["u"])
                    if key not in cache:
                        cache[key] = dither_round(w, state["u"])
                    w = cache[key]
            return TF.linear(x, w, b)

    class TorchShim:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def matmul(a, b):
            if state["on"] and activations:
                a, b = round_f16(a), round_f16(b)
            return torch.matmul(a, b)

    saved = (O.F, O.torch, O.time_embed, O.p_sample)
    time_embed0, p_sample0 = O.time_embed, O.p_sample

    def time_embed(p, t):
        on, state["on"] = state["on"], False
        try:
            return time_embed0(p, t)
        finally:
            state["on"] = on

    def p_sample(p, s, x, t, *a, **kw):
        state["on"] = int(t) >= K and (weights or activations)
        state["u"] = dither_offset(state["step"] % max(sets, 1), sets)
        state["step"] += 1
        try:
            return p_sample0(p, s, x, t, *a, **kw)
        finally:
            state["on"] = False

    O.F, O.torch, O.time_embed, O.p_sample = FShim(), TorchShim(), time_embed, p_sample
    try:
        yield state
    finally:
        O.F, O.torch, O.time_embed, O.p_sample = saved
