"""Why the single-pass fp16 steps read DITHERED weight copies, checked on the CPU with the oracle:
emulating those steps inside the restatement (oracle/rounding.py) shows that rounding the activations is harmless, rounding
the weights is what reaches the final sample, and cycling through the engine's 8 dithered copies shrinks it -- the full-size
evidence (64 windows x 1000 steps on a B200) is profiles/r1bd_rounding_study.txt / r1be_parity_floor.txt."""
import ctypes as C

import numpy as np
import torch

from oracle import egoego_oracle as O
from oracle.gen_golden import Tape, synth_x_start
from oracle.rounding import dither_offset, dither_round, emulate_fp16_steps


def test_emulated_rounding_matches_the_engines_weight_commit():
    """oracle.rounding.dither_round == the host rounding of egoego_commit_weights (egoego_dither_weights_f16), bit for bit."""
    import __graft_entry__ as g
    g.build()
    from egoego_release_b200 import _capi
    lib = _capi.lib()
    w = O.init_params(0)["denoise_fn.motion_transformer.layer_stack.0.self_attn.w_q.weight"].reshape(-1)[:40000].contiguous()
    wn = w.numpy()
    for R in (1, 8, 16):
        for r in range(R):
            out = np.zeros(wn.size, np.uint16)
            assert lib.egoego_dither_weights_f16(wn.ctypes.data_as(C.c_void_p), wn.size, r, R, out.ctypes.data_as(C.c_void_p)) == 0
            mine = dither_round(w, float(np.float32(dither_offset(r, R)))).half().numpy().view(np.uint16)
            assert np.array_equal(out, mine), (R, r)


def test_weight_rounding_dominates_and_dither_shrinks_it(params0):
    N, B, K = 100, 2, 8
    sched = O.make_schedule(N)
    xs = synth_x_start(5, B, 120)
    cm = O.prep_head_condition_mask(xs.shape)
    tp = Tape(9)
    tape = [tp.draw(xs.shape) for _ in range(N + 2)]
    ds = O.MotionDataStub()

    def run(**kw):
        with torch.no_grad():
            if not kw:
                return O.p_sample_loop(params0, sched, xs, cm, lambda k: tape[k])
            with emulate_fp16_steps(K, **kw):
                return O.p_sample_loop(params0, sched, xs, cm, lambda k: tape[k])

    jr = O.joints_from_model_output(ds, run())
    err = {}
    for tag, kw in (("weights", dict(weights=True, activations=False)), ("activations", dict(weights=False, activations=True)),
                    ("both", dict(weights=True, activations=True)), ("both_8_copies", dict(weights=True, activations=True, sets=8))):
        err[tag] = float((O.joints_from_model_output(ds, run(**kw)) - jr).abs().mean()) * 1e3          # mm
    print("mean joint error, mm:", err)
    assert err["weights"] > 5 * err["activations"]            # measured 23x here, 11x at 64 windows x 1000 steps
    assert err["both"] > 0.8 * err["weights"]                 # ... so the weights explain the fp16 steps
    assert err["both_8_copies"] < 0.75 * err["both"]          # measured 0.56x at N = 100 (0.32x at N = 1000, where more steps average)
    assert O.F is torch.nn.functional and O.torch is torch    # the patch is undone
