"""CPU-side checks of the boundary: the library loads, exports every symbol include/egoego_b200.h declares,
refuses to run without a B200, and the host mirror keeps the reference's state_dict / signatures."""
import ctypes as C
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from egoego_release_b200 import _capi
    return _capi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "egoego_b200.h")).read()
    names = set(re.findall(r"\b(egoego_[a-z_0-9]+)\s*\(", hdr))
    names -= {"egoego_ctx"}
    assert len(names) >= 16
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/egoego_b200.h but not exported"
    from egoego_release_b200 import _capi
    assert set(_capi.EXPORTS) == names


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    from egoego_release_b200 import _capi
    h = C.c_void_p()
    cfg = _capi.Cfg(d_feats=198, d_model=512, n_head=4, n_dec_layers=4, d_k=256, d_v=256, max_timesteps=121,
                    timesteps=50, objective=1, max_batch=2, device=0, engine=0)
    assert lib.egoego_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in lib.egoego_last_error()
    import egoego_release_b200 as E
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, timesteps=10, objective="pred_x0")
    x = torch.zeros(1, 120, 198)
    with pytest.raises(E.EgoEgoError):
        m.sample(x, torch.ones_like(x))


def test_state_dict_and_signatures_match_reference():
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, timesteps=1000, objective="pred_x0")
    sd = m.state_dict()
    p = O.init_params(0)
    assert set(p) <= set(sd)
    for k, v in p.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    sched = O.make_schedule(1000)
    for k, v in sched.items():
        assert torch.equal(sd[k], v), k                      # schedule buffers bit-exact vs the pinned oracle
    assert len(sd) == len(p) + 13
    # reference signatures (transformer_cond_diffusion_model.py:248,258,527,547)
    assert list(inspect.signature(m.sample).parameters)[:3] == ["x_start", "cond_mask", "padding_mask"]
    assert list(inspect.signature(m.p_sample).parameters)[:5] == ["x", "t", "x_cond", "clip_denoised", "padding_mask"]
    assert list(inspect.signature(m.sample_sliding_window_w_canonical).parameters)[:5] == \
        ["ds", "global_head_jpos", "global_head_jquat", "x_start", "cond_mask"]
    with pytest.raises(ValueError):
        E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, beta_schedule="quadratic")


def test_condition_mask_glue():
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    d = torch.zeros(2, 7, 198)
    assert torch.equal(E.prep_head_condition_mask(d), O.prep_head_condition_mask(d.shape))
    pm = E.prep_padding_mask(d, torch.tensor([120, 60]))
    assert pm.shape == (2, 1, 121) and pm[1, 0].sum() == 61


def test_mirror_survives_deepcopy_and_pickle():
    """ema_pytorch.EMA deep-copies the model (trainer_amass_cond_motion_diffusion.py:58): engine handles must not travel."""
    import copy
    import pickle
    import egoego_release_b200 as E
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, timesteps=10, objective="pred_x0")
    for c in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert c.denoise_fn._owner() is c and c._h is None
        assert set(c.state_dict()) == set(m.state_dict())


def test_dithered_weight_sets_host_rounding(lib):
    """The host rounding behind the dithered fp16 weight copies (engine_tc.cu: dither_round / dither_offset):
    one set = plain round-to-nearest; every copy stays within one fp16 ulp of the weight; the MEAN over the R copies is
    ~8x (R = 8) closer to the fp32 weight than plain rounding -- the property the sampler relies on to average the weight
    rounding out over steps."""
    import numpy as np
    rng = np.random.default_rng(0)
    w = ((rng.random(50000, dtype=np.float32) * 2 - 1) * np.float32(0.0442)).astype(np.float32)
    w[:4] = [0.0, 0.75, -0.375, 6.0e-6]                                  # fp16-exact values (not at a binade edge) and a subnormal-range weight

    def copy(r, R):
        out = np.zeros(w.size, np.uint16)
        rc = lib.egoego_dither_weights_f16(w.ctypes.data_as(C.c_void_p), w.size, r, R, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out.view(np.float16).astype(np.float32)

    plain = w.astype(np.float16).astype(np.float32)
    assert np.array_equal(copy(0, 1), plain)
    _, e = np.frexp(w)
    ulp = np.ldexp(np.float32(1), np.maximum(e - 11, -24)).astype(np.float32)
    R = 8
    sets = [copy(r, R) for r in range(R)]
    assert len({s.tobytes() for s in sets}) == R                          # the copies differ
    for s in sets:
        assert np.all(np.abs(s - w) <= ulp * 1.0001)
        assert np.array_equal(s[:3], w[:3])                               # fp16-exact weights inside a binade are never moved
    rms = lambda x: float(np.sqrt(np.mean(((x - w) / ulp) ** 2)))
    assert rms(plain) > 0.25 and rms(np.mean(sets, 0)) < 0.05             # 0.29 ulp -> 0.036 ulp
    assert lib.egoego_dither_weights_f16(w.ctypes.data_as(C.c_void_p), w.size, 8, 8, None) != 0


def test_load_state_dict_accepts_reference_checkpoint_layouts():
    """the 'ema' entry of the reference's checkpoints is ``ema_pytorch.EMA.state_dict()`` -- keys
    ``ema_model.denoise_fn.*`` next to ``online_model.*``, ``initted``, ``step`` (trainer_amass_cond_motion_diffusion.py:100-122).
    The mirror must pick the EMA weights, not silently keep its random initialisation."""
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    mk = lambda: E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                         max_timesteps=121, out_dim=198, timesteps=10, objective="pred_x0")
    p, q = O.init_params(0), O.init_params(1)
    ema = {"ema_model." + k: v for k, v in p.items()}
    ema.update({"online_model." + k: v for k, v in q.items()})
    ema.update(initted=torch.tensor(True), step=torch.tensor(7))
    m = mk()
    res = m.load_state_dict(ema, strict=False)
    assert not res.unexpected_keys
    for k, v in p.items():
        assert torch.equal(m.state_dict()[k], v), k            # the EMA weights, not the online ones
    m2 = mk()
    m2.load_state_dict({"module." + k: v for k, v in p.items()}, strict=False)      # DataParallel / DDP prefix
    assert torch.equal(m2.state_dict()["denoise_fn.linear_out.weight"], p["denoise_fn.linear_out.weight"])
    m3 = mk()
    full = dict(m.state_dict())
    m3.load_state_dict(full)                                    # strict round trip of the mirror's own state_dict
    with pytest.raises(KeyError):
        mk().load_state_dict({"encoder.weight": torch.zeros(3)}, strict=False)


def test_precise_all_fp16_is_not_the_zero_value():
    """ADVICE r1 / VERDICT weak #2: a zero-initialised egoego_cfg must select the default precision policy; the all-fp16 mode
    has its own explicit negative value in the header and in the mirror."""
    import egoego_release_b200 as E
    hdr = open(os.path.join(ROOT, "include", "egoego_b200.h")).read()
    assert re.search(r"#define\s+EGOEGO_PRECISE_ALL_FP16\s+\(-2\)", hdr)
    assert E.PRECISE_ALL_FP16 == -2
    src = open(os.path.join(ROOT, "egoego_release_b200", "csrc", "egoego_b200.cu")).read()
    assert "else if (pl <= 0)" in src                           # 0 and -1 resolve to the default policy
