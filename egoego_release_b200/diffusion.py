"""Host-side mirror of the reference's stage-2 sampler boundary.

``CondGaussianDiffusion`` keeps the constructor arguments, method names, argument meaning, return
values, ``state_dict`` keys and error behaviour of the reference class
(egoego/model/transformer_cond_diffusion_model.py:144-625) so the reference's callers
(``Trainer.cond_sample_res`` / ``full_body_gen_cond_head_pose_sliding_window``,
trainer_amass_cond_motion_diffusion.py:233-277; run_egoego.py:149-151; eval_stage2.py:159-161) run
unchanged -- but every numeric step executes in libegoego_b200.so (hand-written sm_100a CUDA) through
the C ABI of include/egoego_b200.h.  PyTorch is used only for device memory, streams and parameter
containers.  There is no PyTorch / CPU fallback: inputs must be CUDA tensors and the library must load.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _capi
from ._capi import Cfg, EgoEgoError, Rng, check

HEAD_IDX = 15
DEFAULT_ENGINE = "tcgen05"   # "simt" = fp32 CUDA-core validation engine


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor, device=None) -> torch.Tensor:
    if device is not None and t.device != device:
        t = t.to(device)
    return t.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# Schedule (transformer_cond_diffusion_model.py:41-57): host-side construction, fp64 like the reference
# ------------------------------------------------------------------------------------------------
def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def _sinusoid_table(n_position, d_hid):
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab = ang.copy()
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    tab[0] = 0.0
    return torch.FloatTensor(tab)


# ------------------------------------------------------------------------------------------------
# Parameter containers with the reference's module tree (=> identical state_dict keys / shapes).
# They own weights only; they are never called (the math lives in the CUDA library).
# ------------------------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v):
        super().__init__()
        self.w_q = nn.Linear(d_model, n_head * d_k)
        self.w_k = nn.Linear(d_model, n_head * d_k)
        self.w_v = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_q.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_k.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_v.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.layer_norm = nn.LayerNorm(d_model)


class _FFNParams(nn.Module):
    def __init__(self, d_in, d_hid):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)


class _LayerParams(nn.Module):
    def __init__(self, d_model, n_head, d_k, d_v):
        super().__init__()
        self.self_attn = _MHAParams(n_head, d_model, d_k, d_v)
        self.pos_ffn = _FFNParams(d_model, d_model)


class _DecoderParams(nn.Module):
    def __init__(self, d_feats, d_model, n_layers, n_head, d_k, d_v, max_timesteps):
        super().__init__()
        self.start_conv = nn.Conv1d(d_feats, d_model, 1)
        self.position_vec = nn.Embedding.from_pretrained(_sinusoid_table(max_timesteps + 1, d_model), freeze=True)
        self.layer_stack = nn.ModuleList([_LayerParams(d_model, n_head, d_k, d_v) for _ in range(n_layers)])


class TransformerDiffusionModel(nn.Module):
    """Mirror of the denoiser (transformer_cond_diffusion_model.py:75-141); ``forward`` runs on the library."""

    def __init__(self, d_feats, d_model, n_dec_layers, n_head, d_k, d_v, max_timesteps):
        super().__init__()
        self.d_feats, self.d_model, self.n_head = d_feats, d_model, n_head
        self.n_dec_layers, self.d_k, self.d_v, self.max_timesteps = n_dec_layers, d_k, d_v, max_timesteps
        self.motion_transformer = _DecoderParams(d_feats * 2, d_model, n_dec_layers, n_head, d_k, d_v, max_timesteps)
        self.linear_out = nn.Linear(d_model, d_feats)
        dim = 64
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, d_model))
        self._owner = None  # set by CondGaussianDiffusion (plain attribute, not a submodule)

    def __getstate__(self):                       # the back-reference is re-bound by the owner's __setstate__
        st = dict(self.__dict__)
        st["_owner"] = None
        return st

    def forward(self, src, noise_t, padding_mask=None):
        owner = object.__getattribute__(self, "_owner")
        if owner is None:
            raise EgoEgoError("TransformerDiffusionModel must be owned by a CondGaussianDiffusion")
        return owner()._denoise(src, noise_t, padding_mask)


class CondGaussianDiffusion(nn.Module):
    def __init__(self, d_feats, d_model, n_head, n_dec_layers, d_k, d_v, max_timesteps, out_dim,
                 timesteps=1000, loss_type='l1', objective='pred_noise', beta_schedule='cosine',
                 p2_loss_weight_gamma=0., p2_loss_weight_k=1, batch_size=None,
                 max_batch: int = 256, engine: Optional[str] = None, precise_last_steps: int = -1):
        super().__init__()
        import weakref
        self.denoise_fn = TransformerDiffusionModel(d_feats=d_feats, d_model=d_model, n_head=n_head, d_k=d_k, d_v=d_v,
                                                    n_dec_layers=n_dec_layers, max_timesteps=max_timesteps)
        object.__setattr__(self.denoise_fn, "_owner", weakref.ref(self))
        self.objective = objective
        self.seq_len = max_timesteps - 1
        self.out_dim = out_dim
        if beta_schedule == 'linear':
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == 'cosine':
            betas = cosine_beta_schedule(timesteps)
        else:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        if objective not in ('pred_noise', 'pred_x0'):
            # the reference raises on first use (:240); raise at the same place (see p_sample)
            pass
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        rb = lambda name, val: self.register_buffer(name, val.to(torch.float32))
        rb('betas', betas)
        rb('alphas_cumprod', alphas_cumprod)
        rb('alphas_cumprod_prev', alphas_cumprod_prev)
        rb('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        rb('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        rb('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        rb('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        rb('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        rb('posterior_variance', posterior_variance)
        rb('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        rb('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        rb('posterior_mean_coef2', (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        rb('p2_loss_weight', (p2_loss_weight_k + alphas_cumprod / (1 - alphas_cumprod)) ** -p2_loss_weight_gamma)

        # ---- engine state (not part of the state_dict) ----
        self._cfg = dict(d_feats=d_feats, d_model=d_model, n_head=n_head, n_dec_layers=n_dec_layers, d_k=d_k, d_v=d_v,
                         max_timesteps=max_timesteps)
        self._max_batch = int(max_batch)
        # 0 / -1: default policy, the last max(ceil(N/16), 48) steps with exact hi/lo weights (the last 16 of them in the 3-term split);
        # K > 0: K such steps (>= N: the 3-term split everywhere); PRECISE_ALL_FP16: none
        self._precise_last_steps = int(precise_last_steps)
        eng = engine or os.environ.get("EGOEGO_ENGINE", DEFAULT_ENGINE)
        if eng not in ("tcgen05", "simt"):
            raise ValueError(f"unknown engine {eng}")
        self._engine = _capi.ENGINE_TCGEN05 if eng == "tcgen05" else _capi.ENGINE_SIMT
        self._h = None
        self._h_device = None
        self._weights_sig = None
        self._skeleton_sig = None
        self._noise_tape = None   # parity mode: explicit [n_draws, B, T, D] tape consumed by the next sample()
        self._tape_keepalive = None

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if self._h is not None:
                _capi.lib().egoego_destroy(self._h)
            if getattr(self, "_ht", None) is not None:
                _capi.lib().egoego_destroy(self._ht)
        except Exception:
            pass

    # engine handles are process-local: drop them when the module is pickled / deep-copied (ema_pytorch.EMA
    # deep-copies the model, trainer_amass_cond_motion_diffusion.py:58) and re-create lazily in the copy
    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_h", "_h_device", "_weights_sig", "_skeleton_sig", "_noise_tape", "_tape_keepalive", "_ht", "_ht_sig", "_ht_device"):
            st[k] = None
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        import weakref
        object.__setattr__(self.denoise_fn, "_owner", weakref.ref(self))

    def _device(self) -> torch.device:
        dev = self.betas.device
        if dev.type != "cuda":
            raise EgoEgoError("egoego_release_b200 runs on a CUDA device only (no CPU fallback): "
                              "move the model with .cuda() / .to('cuda:N')")
        return dev

    def _signature(self):
        """(data_ptr, _version) of every tensor PLUS a content checksum computed on the device: writes through ``.data``
        (``p.data.copy_`` / ``lerp_``: what ema_pytorch's update does) leave ``_version`` untouched, so the version
        counters alone would let the engine keep sampling with stale packed weights."""
        sd = self.state_dict(keep_vars=True)
        dev = self.betas.device
        chk = _capi.content_checksum([v.detach() for v in sd.values()], dev) if dev.type == "cuda" else 0
        return tuple((k, v.data_ptr(), v._version) for k, v in sd.items()) + (("__content__", chk, 0),)

    # ema_pytorch.EMA.state_dict() (what the reference saves under 'ema', trainer_amass_cond_motion_diffusion.py:100-106)
    # carries 'ema_model.<key>', 'online_model.<key>', 'initted', 'step'; DataParallel / DDP add 'module.'.
    _WRAPPER_PREFIXES = ("ema_model.", "module.", "model.")

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the reference's checkpoints as they are on disk: plain keys (``denoise_fn.*``, the 'model' entry) and
        wrapped ones (the 'ema' entry: ``ema_model.denoise_fn.*`` next to ``online_model.*`` / ``initted`` / ``step``).
        A dict without a single ``denoise_fn.*`` tensor is an error even with ``strict=False`` -- silently keeping the
        random initialisation is never what the caller meant."""
        sd = dict(state_dict)
        for pre in self._WRAPPER_PREFIXES:
            if any(k.startswith(pre + "denoise_fn.") for k in sd):
                sd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
                break
        if not any(k.startswith("denoise_fn.") for k in sd):
            raise KeyError("load_state_dict: no 'denoise_fn.*' tensor found (keys look like "
                           f"{sorted(state_dict)[:3]}); expected the reference's 'model' or 'ema' checkpoint entry")
        res = super().load_state_dict(sd, strict=strict, **kw)
        self._weights_sig = None
        return res

    def _handle(self):
        dev = self._device()
        L = _capi.lib()
        if self._h is None or self._h_device != dev:
            if self._h is not None:
                L.egoego_destroy(self._h)
                self._h = None
            if self.objective not in ('pred_noise', 'pred_x0'):
                raise ValueError(f'unknown objective {self.objective}')
            cfg = Cfg(timesteps=self.num_timesteps, objective=1 if self.objective == 'pred_x0' else 0,
                      max_batch=self._max_batch, device=dev.index if dev.index is not None else torch.cuda.current_device(),
                      engine=self._engine, precise_last_steps=self._precise_last_steps, **self._cfg)
            h = C.c_void_p()
            check(L.egoego_create(C.byref(cfg), C.byref(h)))
            self._h, self._h_device, self._weights_sig, self._skeleton_sig = h, dev, None, None
        sig = self._signature()
        if sig != self._weights_sig:
            with torch.cuda.device(dev):
                for k, v in self.state_dict().items():
                    if v.dtype != torch.float32 and k != "denoise_fn.motion_transformer.position_vec.weight":
                        v = v.float()
                    t = _f32c(v.detach(), dev)
                    check(L.egoego_set_tensor(self._h, k.encode(), _ptr(t), t.numel(), 1))
                check(L.egoego_commit_weights(self._h, _stream(dev)))
            self._weights_sig = sig
        return self._h

    def load_weights(self):
        """Explicitly (re)pack the weights into the engine (done lazily on first use otherwise)."""
        self._weights_sig = None
        self._handle()

    def precise_last_steps(self) -> int:
        """Resolved precision policy of the engine: diffusion steps t < K use the 3-term split."""
        return int(_capi.lib().egoego_precise_last_steps(self._handle()))

    def engine_info(self) -> str:
        """Resolved kernel choices of the engine ("engine=tcgen05 sms=148 ln4_clusters=33 c8_clusters=0 zigzag=1 ...")."""
        buf = C.create_string_buffer(512)
        check(_capi.lib().egoego_engine_info(self._handle(), buf, 512))
        return buf.value.decode()

    def weight_sets(self) -> int:
        """Number of dithered fp16 weight sets the single-pass steps cycle through (EGOEGO_WEIGHT_SETS, default 8)."""
        return int(_capi.lib().egoego_weight_sets(self._handle()))

    def time_dominant_kernel(self, B: int, half_fmt: bool, iters: int = 20) -> float:
        """ms per launch of the fused QKV projection kernel (CUDA events on the current stream)."""
        h = self._handle()
        dev = self._device()
        ms = C.c_float()
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_time_dominant_kernel(h, B, 1 if half_fmt else 0, iters, C.byref(ms), _stream(dev)))
        return float(ms.value)

    KERNELS = ("start", "qkv", "attention", "fc_ln", "w1", "w2_ln", "out", "ddpm_update")

    def time_kernel(self, which, B: int, T: int, half_fmt: bool, iters: int = 20) -> float:
        """ms per launch of one kernel of the sampling step (``which``: name in KERNELS or EGOEGO_KERNEL_* id),
        timed in isolation with CUDA events on the current stream (egoego_time_kernel)."""
        if isinstance(which, str):
            which = self.KERNELS.index(which)
        h = self._handle()
        dev = self._device()
        ms = C.c_float()
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_time_kernel(h, B, T, int(which), 1 if half_fmt else 0, iters, C.byref(ms), _stream(dev)))
        return float(ms.value)

    def launches_per_step(self, which) -> int:
        """How often kernel ``which`` runs in one step of the sampling loop (``ddpm_update``: 0 when the DDPM update is
        fused into linear_out's epilogue, EGOEGO_FUSE_DDPM=1)."""
        if isinstance(which, str):
            which = self.KERNELS.index(which)
        return int(_capi.lib().egoego_launches_per_step(self._handle(), int(which)))

    def launch_count(self) -> int:
        return int(_capi.lib().egoego_launch_count(self._h)) if self._h is not None else 0

    def set_noise_tape(self, tape: Optional[torch.Tensor]):
        """Parity mode: the next sampling call reads its Gaussian draws from ``tape``
        ([n_draws, B, T, D] fp32, the reference's draw order) instead of the Philox generator."""
        self._noise_tape = tape

    def _rng(self, dev, B, T) -> Rng:
        tape = self._noise_tape
        if tape is not None:
            self._noise_tape = None
            tape = _f32c(tape, dev)
            if tape.dim() != 4 or tuple(tape.shape[1:]) != (B, T, self._cfg["d_feats"]) or tape.shape[0] < self.num_timesteps + 2:
                raise ValueError(f"noise tape must be [>= {self.num_timesteps + 2}, {B}, {T}, D], got {tuple(tape.shape)}")
            self._tape_keepalive = tape
            return Rng(tape=tape.data_ptr(), seed=0, window_offset=0)
        # throughput mode: one draw from torch's global generator seeds the counter-based Philox streams
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        return Rng(tape=None, seed=seed, window_offset=int(getattr(self, "window_offset", 0)))

    # ------------------------------------------------------------------------------------------
    # denoiser / p_sample / loops (same signatures as the reference)
    # ------------------------------------------------------------------------------------------
    def _denoise(self, src, noise_t, padding_mask=None):
        h = self._handle()
        dev = self._device()
        B, T, D2 = src.shape
        src = _f32c(src, dev)
        t = noise_t.to(device=dev, dtype=torch.int64).contiguous()
        pm = None if padding_mask is None else _f32c(padding_mask.reshape(B, T + 1), dev)
        out = torch.empty(B, T, D2 // 2, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_denoiser_forward(h, _ptr(src), _ptr(t), _ptr(pm), B, T, _ptr(out), _stream(dev)))
        return out

    @torch.no_grad()
    def p_sample(self, x, t, x_cond, clip_denoised=True, padding_mask=None, noise=None, inpaint=None):
        """x_{t-1} from x_t (reference :248-256).  ``noise`` (optional) replaces the internal draw;
        ``inpaint`` [B, n, D] (optional) overwrites the first n frames of the result (:395-397)."""
        h = self._handle()
        dev = self._device()
        if self.objective not in ('pred_noise', 'pred_x0'):
            raise ValueError(f'unknown objective {self.objective}')
        B, T, D = x.shape
        x = _f32c(x, dev)
        x_cond = _f32c(x_cond, dev)
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        pm = None if padding_mask is None else _f32c(padding_mask.reshape(B, T + 1), dev)
        nz = None if noise is None else _f32c(noise, dev)
        ip = None if inpaint is None else _f32c(inpaint, dev)
        out = torch.empty_like(x)
        rng = Rng(tape=None, seed=int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()) if nz is None else 0,
                  window_offset=0)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_p_sample_step(h, _ptr(x), _ptr(t), _ptr(x_cond), _ptr(nz), C.byref(rng), 0,
                                                  _ptr(pm), 1 if clip_denoised else 0, _ptr(ip),
                                                  0 if ip is None else ip.shape[1], B, T, _ptr(out), _stream(dev)))
        return out

    @torch.no_grad()
    def p_sample_loop(self, shape, x_start, cond_mask, padding_mask=None, x_init=None, inpaint=None):
        """Reference :258-270.  ``padding_mask`` is accepted and ignored exactly like the reference's
        ``sample()`` (which never forwards it, :531-532)."""
        h = self._handle()
        dev = self._device()
        B, T, D = shape
        x_start = _f32c(x_start, dev)
        cond_mask = _f32c(cond_mask, dev)
        xi = None if x_init is None else _f32c(x_init, dev)
        ip = None if inpaint is None else _f32c(inpaint, dev)
        out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        rng = self._rng(dev, B, T)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_sample(h, _ptr(x_start), _ptr(cond_mask), B, T, C.byref(rng), _ptr(xi), _ptr(ip),
                                            0 if ip is None else ip.shape[1], _ptr(out), _stream(dev)))
        return out

    @torch.no_grad()
    def sample(self, x_start, cond_mask, padding_mask=None):
        """Drop-in for CondGaussianDiffusion.sample (reference :527-535)."""
        self.denoise_fn.eval()
        res = self.p_sample_loop(x_start.shape, x_start, cond_mask)
        self.denoise_fn.train()
        return res

    @torch.no_grad()
    def sample_host(self, x_start: torch.Tensor, cond_mask: torch.Tensor, out: Optional[torch.Tensor] = None):
        """End-to-end entry with HOST tensors (pinned or pageable): H2D, loop and D2H inside one call."""
        h = self._handle()
        dev = self._device()
        assert x_start.device.type == "cpu" and cond_mask.device.type == "cpu"
        x_start = x_start.to(torch.float32).contiguous()
        cond_mask = cond_mask.to(torch.float32).contiguous()
        B, T, D = x_start.shape
        if out is None:
            out = torch.empty(B, T, D, dtype=torch.float32).pin_memory()
        rng = self._rng(torch.device("cpu"), B, T) if self._noise_tape is not None else self._rng(dev, B, T)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_sample_host(h, _ptr(x_start), _ptr(cond_mask), B, T, C.byref(rng), _ptr(out), _stream(dev)))
        return out

    # ------------------------------------------------------------------------------------------
    # post-processing + sliding window (reference :329-525)
    # ------------------------------------------------------------------------------------------
    def _set_skeleton(self, ds):
        h = self._handle()
        # The reference's AMASSDataset has no parents attribute (get_smpl_parents is a module-level function that reads the
        # licensed model.npz, egoego/data/amass_diffusion_dataset.py:83-90): fall back to the packaged 22-joint SMPL kintree.
        if hasattr(ds, "parents"):
            par = ds.parents
        elif hasattr(ds, "get_smpl_parents"):
            par = ds.get_smpl_parents()
        else:
            from .motion_data import smpl22_parents
            par = smpl22_parents()
        parents = np.ascontiguousarray(np.asarray(par, dtype=np.int32)[:22])
        off = np.ascontiguousarray(ds.rest_human_offsets.detach().cpu().numpy().reshape(-1).astype(np.float32))
        jmin = np.ascontiguousarray(ds.global_jpos_min.detach().cpu().numpy().reshape(-1).astype(np.float32))
        jmax = np.ascontiguousarray(ds.global_jpos_max.detach().cpu().numpy().reshape(-1).astype(np.float32))
        sig = (parents.tobytes(), off.tobytes(), jmin.tobytes(), jmax.tobytes())
        if sig != self._skeleton_sig:
            check(_capi.lib().egoego_set_skeleton(h, parents.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                                 jmin.ctypes.data_as(C.c_void_p), jmax.ctypes.data_as(C.c_void_p)))
            self._skeleton_sig = sig
        return h

    @torch.no_grad()
    def postprocess(self, ds, all_res_list, recover_rot_quat=None, with_fk=False):
        """convert_model_res_to_data (+ fk_smpl) on the device.
        Returns (aa[B,T,22,3], root[B,T,3], head[B,T,3]) and, if with_fk, also (jpos[B,T,22,3], gquat[B,T,22,4])."""
        h = self._set_skeleton(ds)
        dev = self._device()
        x = _f32c(all_res_list, dev)
        B, T, _ = x.shape
        rq = None
        if recover_rot_quat is not None:
            rq = torch.as_tensor(recover_rot_quat).to(device=dev, dtype=torch.float32).reshape(B, 4).contiguous()
        aa = torch.empty(B, T, 22, 3, device=dev)
        root = torch.empty(B, T, 3, device=dev)
        head = torch.empty(B, T, 3, device=dev)
        jpos = torch.empty(B, T, 22, 3, device=dev) if with_fk else None
        gq = torch.empty(B, T, 22, 4, device=dev) if with_fk else None
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_postprocess(h, _ptr(x), _ptr(rq), B, T, _ptr(aa), _ptr(root), _ptr(head), _ptr(jpos),
                                                 _ptr(gq), _stream(dev)))
        return (aa, root, head, jpos, gq) if with_fk else (aa, root, head)

    def convert_model_res_to_data(self, ds, all_res_list, recover_rot_quat, curr_global_head_jpos=None):
        """Same signature / returns as the reference (:469-525); recover_rot_quat is BS x 1 x 1 x 4 (numpy or tensor)."""
        return self.postprocess(ds, all_res_list, recover_rot_quat)

    @torch.no_grad()
    def fk_smpl(self, ds, root_trans, lrot_aa):
        h = self._set_skeleton(ds)
        dev = self._device()
        root = _f32c(root_trans, dev)
        aa = _f32c(lrot_aa, dev)
        n = root.shape[0]
        gq = torch.empty(n, 22, 4, device=dev)
        gj = torch.empty(n, 22, 3, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_fk_smpl(h, _ptr(root), _ptr(aa), n, _ptr(gq), _ptr(gj), _stream(dev)))
        return gq, gj

    @torch.no_grad()
    def canonicalize_head(self, ds, head_jpos, head_jquat):
        """Device version of rotate_at_frame_smplh + x_start construction for one window
        (reference :358-386).  Returns (x_start[B,T,198], recover_quat[B,4])."""
        h = self._set_skeleton(ds)
        dev = self._device()
        hp = _f32c(head_jpos, dev)
        hq = _f32c(head_jquat, dev)
        B, T, _ = hp.shape
        xs = torch.empty(B, T, 198, device=dev)
        rq = torch.empty(B, 4, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_canonicalize_head(h, _ptr(hp), _ptr(hq), T, B, T, _ptr(xs), _ptr(rq), _stream(dev)))
        return xs, rq

    @torch.no_grad()
    def p_sample_loop_sliding_window_w_canonical(self, ds, shape, global_head_jpos, global_head_jquat, cond_mask,
                                                 noise_fn=None):
        """Reference :329-467, fully device-resident: canonicalisation, conditioning build, the N-step loop with
        per-step in-painting, post-processing, stitching and FK all run in the CUDA library; no numpy round trips.
        ``noise_fn(shape)`` (optional, parity mode) supplies Gaussian tensors in the reference's draw order."""
        dev = self._device()
        self._handle()
        b = shape[0]
        D = self._cfg["d_feats"]
        ghp = _f32c(global_head_jpos, dev)
        ghq = _f32c(global_head_jquat, dev)
        cond_mask = _f32c(cond_mask, dev)
        N = self.num_timesteps
        if noise_fn is not None:
            x_all = _f32c(noise_fn(tuple(shape)), dev)
        else:
            x_all = torch.randn(shape, device=dev)
        whole_aa = whole_root = whole_head = None
        num_steps = ghp.shape[1]
        overlap = 10
        stride = self.seq_len - overlap
        inpaint = None
        for t_idx in range(0, num_steps, stride):
            curr_x = x_all[:, t_idx:t_idx + self.seq_len]
            Tw = curr_x.shape[1]
            if Tw <= self.seq_len - stride:
                break
            xs, rq = self.canonicalize_head(ds, ghp[:, t_idx:t_idx + self.seq_len], ghq[:, t_idx:t_idx + self.seq_len])
            cm = cond_mask[:, t_idx:t_idx + self.seq_len].contiguous()
            if noise_fn is not None:
                tape = torch.empty(N + 2, b, Tw, D, device=dev)
                tape[0] = curr_x
                tape[1] = _f32c(noise_fn((b, Tw, D)), dev)
                for k in range(N):
                    tape[2 + k] = _f32c(noise_fn((b, Tw, D)), dev)
                self.set_noise_tape(tape)
            res = self.p_sample_loop((b, Tw, D), xs, cm, x_init=curr_x.contiguous(), inpaint=inpaint)
            aa, root, head, _, _ = self.postprocess(ds, res, rq, with_fk=False) + (None, None)
            if t_idx == 0:
                whole_aa, whole_root, whole_head = aa, root, head
            else:
                move = whole_head[:, -1:, :] - head[:, overlap - 1:overlap, :]
                root = root + move
                head = head + move
                whole_aa = torch.cat((whole_aa, aa[:, overlap:]), dim=1)
                whole_root = torch.cat((whole_root, root[:, overlap:]), dim=1)
                whole_head = torch.cat((whole_head, head[:, overlap:]), dim=1)
            # conditioning for the next window: FK of the tail, re-canonicalised at its first frame
            gq, gj = self.fk_smpl(ds, root.reshape(-1, 3), aa.reshape(-1, 22, 3))
            gq = gq.reshape(b, -1, 22, 4)[:, -overlap:].contiguous()
            gj = gj.reshape(b, -1, 22, 3)[:, -overlap:].contiguous()
            inpaint = self._tail_condition(ds, gq, gj)
        return whole_aa, whole_root

    def _tail_condition(self, ds, gq, gj):
        """Reference :423-464 in the CUDA library: [b,n,22,4] / [b,n,22,3] tail FK -> in-paint tensor [b,n,198]."""
        h = self._set_skeleton(ds)
        dev = self._device()
        b, n = gq.shape[0], gq.shape[1]
        out = torch.empty(b, n, 198, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_tail_condition(h, _ptr(gq), _ptr(gj), b, n, _ptr(out), _stream(dev)))
        return out

    @torch.no_grad()
    def sample_sliding_window_w_canonical(self, ds, global_head_jpos, global_head_jquat, x_start, cond_mask, noise_fn=None):
        """Drop-in for the reference method (:547-555) -> (aa[B,T',22,3], root[B,T',3])."""
        self.denoise_fn.eval()
        res = self.p_sample_loop_sliding_window_w_canonical(ds, x_start.shape, global_head_jpos, global_head_jquat,
                                                            cond_mask, noise_fn=noise_fn)
        self.denoise_fn.train()
        return res

    # ------------------------------------------------------------------------------------------
    # training-side methods of the reference class: loss AND gradients come from the CUDA library
    # (egoego_train_step: forward with saved activations + backward kernels); torch.autograd only carries the parameter
    # gradients to .grad so the reference's optimizer / EMA / GradScaler code runs unchanged.  Dropout (train() mode) uses
    # counter-based Philox masks re-derived in the backward pass (include/egoego_b200.h: egoego_train_set_dropout).
    # ------------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        """:557-563 (device tensors; the same arithmetic runs inside egoego_train_step)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        sa = self.sqrt_alphas_cumprod.gather(-1, t).reshape(-1, 1, 1)
        sb = self.sqrt_one_minus_alphas_cumprod.gather(-1, t).reshape(-1, 1, 1)
        return sa * x_start + sb * noise

    def _train_handle(self, B):
        """A second, fp32 (SIMT) engine handle for training; weights are re-committed whenever a parameter changed."""
        dev = self._device()
        L = _capi.lib()
        if getattr(self, "_ht", None) is None or self._ht_batch < B or self._ht_device != dev:
            if getattr(self, "_ht", None) is not None:
                L.egoego_destroy(self._ht)
            cfg = Cfg(timesteps=self.num_timesteps, objective=1 if self.objective == 'pred_x0' else 0, max_batch=int(B),
                      device=dev.index if dev.index is not None else torch.cuda.current_device(), engine=_capi.ENGINE_SIMT,
                      precise_last_steps=-1, **self._cfg)
            h = C.c_void_p()
            check(L.egoego_create(C.byref(cfg), C.byref(h)))
            self._ht, self._ht_batch, self._ht_device, self._ht_sig = h, int(B), dev, None
        with torch.cuda.device(dev):
            if self._ht_sig is None:                     # first use: full commit (host staging, workspace allocation)
                for k, v in self.state_dict().items():
                    t = _f32c(v.detach().float(), dev)
                    check(L.egoego_set_tensor(self._ht, k.encode(), _ptr(t), t.numel(), 1))
                check(L.egoego_commit_weights(self._ht, _stream(dev)))
                self._ht_sig = True
            else:
                # every later call: device-to-device refresh of EVERY parameter (72 small async copies, 44 MB, no host sync).  An
                # optimizer step changes all of them anyway, and an unconditional refresh also picks up edits that bypass torch's
                # version counters (p.data.copy_ / lerp_) without the host round trip a content checksum would need.
                ch = [(k, _f32c(v.detach(), dev)) for k, v in self.named_parameters()]
                n = len(ch)
                names = (C.c_char_p * n)(*[k.encode() for k, _ in ch])
                ptrs = (C.c_void_p * n)(*[t.data_ptr() for _, t in ch])
                nums = (C.c_int64 * n)(*[t.numel() for _, t in ch])
                check(L.egoego_update_tensors_device(self._ht, n, names, ptrs, nums, _stream(dev)))
        return self._ht

    DROPOUT_P = 0.1     # nn.Dropout(0.1) at the three sites of every DecoderLayer (transformer_module.py:53,59,105)

    def set_grad_sync(self, group=None, enabled: bool = True):
        """Data-parallel training WITHOUT a DistributedDataParallel wrapper: ``loss.backward()`` averages the gradients over the ranks
        of ``group`` (default: the world group) itself, with ONE all-reduce of the flat 44 MB buffer the CUDA backward fills.  The
        backward pass of this engine produces every gradient at once, so torch DDP's bucket-by-bucket overlap has nothing to
        overlap with and its bucket copies are pure overhead (measured on 8 x B200: 0.92 ms exposed per step under DDP).  The
        caller keeps the ranks' parameters identical at start (same seed or a broadcast), as DDP would.  ``enabled=False`` (or
        the ``no_grad_sync()`` context) restores purely local gradients, e.g. for gradient accumulation."""
        self._grad_sync = (group, True) if enabled else None

    def no_grad_sync(self):
        """Context manager: gradients of the backward passes inside stay local (DDP's ``no_sync``)."""
        import contextlib

        @contextlib.contextmanager
        def _ctx():
            old = getattr(self, "_grad_sync", None)
            self._grad_sync = None
            try:
                yield
            finally:
                self._grad_sync = old
        return _ctx()

    def p_losses(self, x_start, cond_mask, t, noise=None, padding_mask=None, cond_noise=None, dropout_seed=None):
        """:574-605.  Returns the scalar loss; ``loss.backward()`` fills ``.grad`` of every trainable parameter with the
        gradients computed by the CUDA backward pass.  ``cond_noise`` (optional) replaces the second Gaussian draw.
        While ``denoise_fn`` is in train() mode the step applies the reference's dropout (p = 0.1, three sites per layer) with
        counter-based Philox masks; their seed is drawn from torch's generator (so ``torch.manual_seed`` makes a step
        reproducible) unless ``dropout_seed`` is given.  In eval() mode dropout is the identity, as in the reference."""
        if self.objective not in ('pred_noise', 'pred_x0'):
            raise ValueError(f'unknown objective {self.objective}')
        if self.loss_type not in ('l1', 'l2'):
            raise ValueError(f'invalid loss type {self.loss_type}')
        dev = self._device()
        x_start = _f32c(x_start, dev)
        cond_mask = _f32c(cond_mask, dev)
        noise = torch.randn_like(x_start) if noise is None else _f32c(noise, dev)
        cond_noise = torch.randn_like(x_start) if cond_noise is None else _f32c(cond_noise, dev)
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        B, T, _ = x_start.shape
        pm = None if padding_mask is None else _f32c(padding_mask.reshape(B, T + 1), dev)
        names, params = zip(*[(k, v) for k, v in self.named_parameters() if v.requires_grad])
        if self.denoise_fn.training:
            seed = int(dropout_seed) if dropout_seed is not None else int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
            self._dropout = (self.DROPOUT_P, seed)
        else:
            self._dropout = (0.0, 0)
        return _TrainStepFn.apply(self, names, x_start, cond_mask, pm, t, noise, cond_noise, *params)

    def forward(self, x_start, cond_mask, padding_mask=None):
        """:617-625: t ~ U{0..N-1} per sample, then p_losses."""
        bs = x_start.shape[0]
        t = torch.randint(0, self.num_timesteps, (bs,), device=x_start.device).long()
        return self.p_losses(x_start, cond_mask, t, padding_mask=padding_mask)


class _TrainStepFn(torch.autograd.Function):
    """One call of egoego_train_step; backward hands the gradients stored in the engine handle to autograd."""

    @staticmethod
    def forward(ctx, model, names, x_start, cond_mask, pm, t, noise, cond_noise, *params):
        dev = x_start.device
        B, T, _ = x_start.shape
        h = model._train_handle(B)
        sa = model.sqrt_alphas_cumprod.gather(-1, t).contiguous()
        sb = model.sqrt_one_minus_alphas_cumprod.gather(-1, t).contiguous()
        wt = model.p2_loss_weight.gather(-1, t).contiguous()
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        dp, dseed = getattr(model, "_dropout", (0.0, 0))
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_train_set_dropout(h, float(dp), int(dseed)))
            check(_capi.lib().egoego_train_step(h, _ptr(x_start), _ptr(cond_mask), _ptr(pm), _ptr(t), _ptr(noise), _ptr(cond_noise), _ptr(sa),
                                                _ptr(sb), _ptr(wt), 1 if model.loss_type == 'l2' else 0, B, T, _ptr(loss), _stream(dev)))
        ctx.model, ctx.names, ctx.shapes, ctx.handle = model, names, [tuple(p.shape) for p in params], h
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        model, dev = ctx.model, gout.device
        if model._ht is not ctx.handle:
            raise EgoEgoError("the training handle changed between forward and backward")
        numels = [int(torch.Size(sh).numel()) for sh in ctx.shapes]
        flat = torch.empty(sum(numels), device=dev, dtype=torch.float32)      # one buffer, one FFI call; gradients are views into it
        n = len(numels)
        offs = [0]
        for k in numels:
            offs.append(offs[-1] + k)
        names = (C.c_char_p * n)(*[nm.encode() for nm in ctx.names])
        ptrs = (C.c_void_p * n)(*[flat.data_ptr() + 4 * o for o in offs[:-1]])
        nums = (C.c_int64 * n)(*numels)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_train_get_grads(ctx.handle, n, names, ptrs, nums, _stream(dev)))
        flat.mul_(gout)
        sync = getattr(model, "_grad_sync", None)
        if sync is not None:                      # set_grad_sync(): one all-reduce of the whole gradient set, averaged like DDP
            from .parallel import average_flat_gradients
            average_flat_gradients(flat, sync[0])
        grads = [flat[offs[i]:offs[i + 1]].view(ctx.shapes[i]) for i in range(n)]
        return (None,) * 8 + tuple(grads)
