"""Host mirror of EgoEgo's stage-1 networks in the shipped configuration (``--input_of_feats``).

``HeadFormer`` / ``HeadNormalFormer`` keep the reference's class names, constructor arguments (``opt``, ``device``),
``state_dict`` keys and method names (egoego/model/head_estimation_transformer.py, head_normal_estimation_transformer.py),
so ``load_state_dict(ckpt['transformer_encoder_state_dict'])`` and the calls in run_egoego.py:74-115 work unchanged.
All arithmetic runs in libegoego_b200 (csrc/stage1.cu) through the C ABI (``egoego_seqnet_*``, ``egoego_va2rot`` ...):
there is no PyTorch / CPU fallback, CPU tensors are moved to the module's CUDA device.

``HeadFormer(opt)`` with ``opt.input_of_feats=False`` runs the ResNet-18 optical-flow encoder (egoego/model/resnet.py) in
csrc/resnet.cu (inference, eval-mode BatchNorm folded into fused conv+bias+residual+ReLU kernels).
Not built (and why): training losses, and evo's Umeyama xy-plane fit inside ``HeadNormalFormer.forward_for_eval`` (evo is a
third-party dependency that is absent here) -- that one step is a host callable ``xy_align`` (default: evo when importable,
else the restated published algorithm, "parity unpinned").
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Callable, Optional

import numpy as np
import torch
from torch import nn

from . import _capi
from ._capi import EgoEgoError, SeqNetCfg, check
from .diffusion import _DecoderParams


class _MLPParams(nn.Module):
    """Parameter holder with the state_dict layout of egoego/model/mlp.py:4-20."""

    def __init__(self, input_dim, hidden_dims):
        super().__init__()
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32(t, dev) -> torch.Tensor:
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=dev, dtype=torch.float32).contiguous()


class _SeqNetModule(nn.Module):
    """Decoder + MLP heads living in the CUDA library; subclasses declare ``_heads`` = [(attribute prefix, hidden, out)]."""

    _heads = ()

    def _init_net(self, d_feats, d_model, n_layers, n_head, d_k, d_v, window, max_batch=8):
        self._net_cfg = dict(d_feats=d_feats, d_model=d_model, n_layers=n_layers, n_head=n_head, d_k=d_k, d_v=d_v, window=window,
                             max_batch=max_batch)
        self.action_transformer = _DecoderParams(d_feats, d_model, n_layers, n_head, d_k, d_v, window)
        self._h = None
        self._sig = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _capi.lib().egoego_seqnet_destroy(self._h)
        except Exception:
            pass

    def _cuda_device(self) -> torch.device:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise EgoEgoError("egoego_release_b200 stage-1 nets run on a B200 only (no CPU fallback): move the module with .to('cuda')")
        return dev

    def _handle(self):
        dev = self._cuda_device()
        sd_now = self.state_dict()
        sig = (dev.index or 0, tuple((k, v._version, v.data_ptr()) for k, v in sd_now.items()),
               _capi.content_checksum(list(sd_now.values()), dev))         # .data edits do not bump _version (see _capi.content_checksum)
        if self._h is not None and sig == self._sig:
            return self._h
        L = _capi.lib()
        if self._h is not None:
            L.egoego_seqnet_destroy(self._h)
            self._h = None
        g = self._net_cfg
        cfg = SeqNetCfg(d_feats=g["d_feats"], d_model=g["d_model"], n_head=g["n_head"], n_layers=g["n_layers"], d_k=g["d_k"],
                        d_v=g["d_v"], window=g["window"], max_batch=g["max_batch"], device=dev.index or 0, n_heads=len(self._heads))
        for i, (_, hidden, out) in enumerate(self._heads):
            cfg.head_n_hidden[i] = len(hidden)
            for j, nh in enumerate(hidden):
                cfg.head_hidden[i][j] = nh
            cfg.head_out[i] = out
        h = C.c_void_p()
        check(L.egoego_seqnet_create(C.byref(cfg), C.byref(h)))
        try:
            for k, v in self.state_dict().items():
                name = None
                if k.startswith("action_transformer."):
                    name = k[len("action_transformer."):]
                else:
                    for i, (pre, _, _) in enumerate(self._heads):
                        if k.startswith(pre + "_mlp."):
                            name = f"head{i}." + k[len(pre) + 5:]
                        elif k.startswith(pre + "_fc."):
                            name = f"head{i}.fc." + k[len(pre) + 4:]
                if name is None:
                    continue
                t = v.detach().to("cpu", torch.float32).contiguous()
                check(L.egoego_seqnet_set_tensor(h, name.encode(), t.data_ptr(), t.numel()))
            check(L.egoego_seqnet_commit(h))
        except Exception:
            L.egoego_seqnet_destroy(h)
            raise
        self._h, self._sig = h, sig
        return h

    def launch_count(self) -> int:
        return int(_capi.lib().egoego_seqnet_launch_count(self._h)) if self._h is not None else 0

    def _net_forward(self, feats: torch.Tensor, token0_only: bool, want_dec: bool = False):
        """feats [B,T,d_feats] on the module's device -> (dec [B,window,d] or None, [head outputs])."""
        h = self._handle()
        dev = self._cuda_device()
        feats = _f32(feats, dev)
        B, T, D = feats.shape
        g = self._net_cfg
        if D != g["d_feats"]:
            raise ValueError(f"expected {g['d_feats']} input features, got {D}")
        dec = torch.empty(B, g["window"], g["d_model"], device=dev) if want_dec else None
        outs = [torch.empty((B, out) if token0_only else (B, T, out), device=dev) for _, _, out in self._heads]
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        mb = g["max_batch"]
        with torch.cuda.device(dev):
            for b0 in range(0, B, mb):
                b1 = min(B, b0 + mb)
                sl = [None if o is None else o[b0:b1] for o in ([dec] + outs + [None, None])[:3]]
                check(_capi.lib().egoego_seqnet_forward(h, ptr(feats[b0:b1]), b1 - b0, T, ptr(sl[0]), ptr(sl[1]), ptr(sl[2]),
                                                        1 if token0_only else 0, _stream(dev)))
        return dec, outs


class _ResNetParams(nn.Module):
    """Parameter holder with the state_dict layout of egoego/model/resnet.py's ``ResNet`` (``resnet.*`` = torchvision's
    resnet18 with ``fc`` -> out_dim).  Never run through PyTorch: the forward pass is csrc/resnet.cu."""

    def __init__(self, out_dim):
        super().__init__()
        try:
            from torchvision import models
        except ImportError as ex:                        # the reference needs torchvision for this path too
            raise EgoEgoError("the ResNet-18 flow encoder needs torchvision for its parameter container") from ex
        self.out_dim = out_dim
        self.resnet = models.resnet18(weights=None)
        self.resnet.fc = nn.Linear(self.resnet.fc.in_features, out_dim)


class HeadFormer(_SeqNetModule):
    """HeadNet (egoego/model/head_estimation_transformer.py:48-308): optical-flow features -> head angular velocity
    (integrated to rotations) and per-frame travelled distance (fixes the scale of the SLAM trajectory)."""

    _heads = (("action_va", (1024, 512, 256), 3), ("action_dist", (1024, 512, 256), 1))

    def __init__(self, opt, device=None):
        super().__init__()
        self.opt = opt
        self.device = device
        self.cnn_fdim = 512
        self.transformer_window_size = opt.window
        self.input_of_feats = getattr(opt, "input_of_feats", True)
        self._cnn_h = None
        self._cnn_sig = None
        if not self.input_of_feats:                      # raw optical flow: ResNet-18 encoder (inference only; always frozen here)
            self.cnn = _ResNetParams(self.cnn_fdim)
            self.freeze_of_cnn = getattr(opt, "freeze_of_cnn", True)
        self._init_net(self.cnn_fdim, opt.d_model, opt.n_dec_layers, opt.n_head, opt.d_k, opt.d_v, opt.window)
        self.action_va_mlp = _MLPParams(opt.d_model, self._heads[0][1])
        self.action_va_fc = nn.Linear(256, 3)
        self.action_dist_mlp = _MLPParams(opt.d_model, self._heads[1][1])
        self.action_dist_fc = nn.Linear(256, 1)

    def __del__(self):
        try:
            if getattr(self, "_cnn_h", None) is not None:
                _capi.lib().egoego_resnet18_destroy(self._cnn_h)
        except Exception:
            pass
        super().__del__()

    def _cnn_handle(self):
        dev = self._cuda_device()
        sd = self.cnn.state_dict()
        sig = (dev.index or 0, tuple((k, v._version, v.data_ptr()) for k, v in sd.items()),
               _capi.content_checksum(list(sd.values()), dev))
        if self._cnn_h is not None and sig == self._cnn_sig:
            return self._cnn_h
        L = _capi.lib()
        if self._cnn_h is not None:
            L.egoego_resnet18_destroy(self._cnn_h)
            self._cnn_h = None
        h = C.c_void_p()
        check(L.egoego_resnet18_create(dev.index or 0, self.cnn_fdim, C.byref(h)))
        try:
            for k, v in sd.items():
                if not k.startswith("resnet.") or k.endswith("num_batches_tracked"):
                    continue
                t = v.detach().to("cpu", torch.float32).contiguous()
                check(L.egoego_resnet18_set_tensor(h, k[len("resnet."):].encode(), t.data_ptr(), t.numel()))
            check(L.egoego_resnet18_commit(h))
        except Exception:
            L.egoego_resnet18_destroy(h)
            raise
        self._cnn_h, self._cnn_sig = h, sig
        return h

    @torch.no_grad()
    def _input_features(self, data):
        """:215-224: pre-extracted features [B,T,512], or raw flow [B,T,224,224,2] through the ResNet-18 encoder (eval-mode
        BatchNorm; the reference appends one zero channel, which the kernel's channel padding subsumes)."""
        dev = self._cuda_device()
        of = _f32(data["of"], dev)
        if self.input_of_feats:
            return of
        if of.dim() != 5 or tuple(of.shape[2:]) != (224, 224, 2):
            raise ValueError("optical flow must be [B,T,224,224,2]")
        B, T = of.shape[:2]
        feats = torch.empty(B * T, self.cnn_fdim, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_resnet18_forward(self._cnn_handle(), of.data_ptr(), B * T, feats.data_ptr(), _stream(dev)))
        return feats.reshape(B, T, self.cnn_fdim)

    @torch.no_grad()
    def va2rot(self, curr_rot, pred_head_vels, dt=1 / 30):
        """:102-124.  curr_rot [B,4] (wxyz), pred_head_vels [B,T,3] -> [B,T+1,4]."""
        dev = self._cuda_device()
        q0, va = _f32(curr_rot, dev), _f32(pred_head_vels, dev)
        B, T, _ = va.shape
        out = torch.empty(B, T + 1, 4, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_va2rot(dev.index or 0, q0.data_ptr(), va.data_ptr(), B, T, float(dt), out.data_ptr(), _stream(dev)))
        return out

    @torch.no_grad()
    def cal_scale_for_slam_w_pred_scale(self, slam_trans, dist_scalar, dist_scale: float = 1.0):
        """:184-212.  slam_trans [(T+1),3], dist_scalar [T'] (already divided by opt.dist_scale unless ``dist_scale`` is given)
        -> rescaled_trans [(T+1),3], scale (0-d tensor)."""
        dev = self._cuda_device()
        st, ds = _f32(slam_trans, dev), _f32(dist_scalar, dev).reshape(-1)
        out = torch.empty_like(st)
        scale = torch.empty(1, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_rescale_slam(dev.index or 0, st.data_ptr(), st.shape[0], ds.data_ptr(), ds.numel(), float(dist_scale),
                                                  out.data_ptr(), scale.data_ptr(), _stream(dev)))
        return out, scale[0]

    @torch.no_grad()
    def _va_dist(self, input_features):
        """Network part of forward_for_eval (:232-262): blocks of ``window`` frames, the full blocks batched in one call."""
        W = self.transformer_window_size
        B, T, _ = input_features.shape
        n_full, rem = T // W, T % W
        vas, dists = [], []
        if n_full:
            blocks = input_features[:, :n_full * W].reshape(B * n_full, W, -1)
            _, (va, dist) = self._net_forward(blocks, token0_only=False)
            vas.append(va.reshape(B, n_full * W, 3)); dists.append(dist.reshape(B, n_full * W, 1))
        if rem:
            _, (va, dist) = self._net_forward(input_features[:, n_full * W:], token0_only=False)
            vas.append(va); dists.append(dist)
        return torch.cat(vas, dim=1), torch.cat(dists, dim=1)

    @torch.no_grad()
    def forward(self, data):
        """:126-182 for sequences of at most one window: head_va, head_rot_quat [B,T+1,4], head_dist_scalar."""
        dev = self._cuda_device()
        feats = self._input_features(data)
        if feats.shape[1] > self.transformer_window_size:
            raise ValueError("forward() takes one window; use forward_for_eval for longer sequences")
        va, dist = self._va_dist(feats)
        out = defaultdict(list)
        out["head_va"] = va
        out["head_rot_quat"] = self.va2rot(_f32(data["head_pose"], dev)[:, 0, 3:], va)
        out["head_dist_scalar"] = dist
        return out

    @torch.no_grad()
    def forward_for_eval(self, data):
        """:214-308.  data: 'of' [1,T,512], 'aligned_slam_trans' [1,T+1,3], 'head_pose' [1,T+1,7] (first quaternion used)
        -> {'head_pose': [1,T'',7], 'pred_scale': 0-d tensor}."""
        dev = self._cuda_device()
        feats = self._input_features(data)
        va, dist = self._va_dist(feats)
        # the reference integrates block by block, restarting from the previous block's last rotation: one continuous scan
        quat = self.va2rot(_f32(data["head_pose"], dev)[:, 0, 3:], va)
        slam = _f32(data["aligned_slam_trans"], dev)
        trans, scale = self.cal_scale_for_slam_w_pred_scale(slam[0], dist[0, :, 0], float(self.opt.dist_scale))
        if trans.shape[0] != quat.shape[1]:
            quat = quat[:, :trans.shape[0]]
        out = defaultdict(list)
        out["head_pose"] = torch.cat((trans[None], quat), dim=-1)
        out["pred_scale"] = scale
        return out


def umeyama_alignment(x: np.ndarray, y: np.ndarray, with_scale: bool = True):
    """Least-squares similarity transform y ~ c R x + t (Umeyama 1991), as evo.core.geometry.umeyama_alignment states it;
    x, y [3,n].  Restated from the published algorithm: evo is absent here, so this is "parity unpinned"."""
    n = x.shape[1]
    mx, my = x.mean(axis=1), y.mean(axis=1)
    sx = 1.0 / n * (np.linalg.norm(x - mx[:, None]) ** 2)
    cov = 1.0 / n * (y - my[:, None]) @ (x - mx[:, None]).T
    u, d, vt = np.linalg.svd(cov)
    s = np.eye(3)
    if np.linalg.det(u) * np.linalg.det(vt) < 0.0:
        s[2, 2] = -1
    r = u @ s @ vt
    c = 1 / sx * np.trace(np.diag(d) @ s) if with_scale else 1.0
    return r, my - c * r @ mx, c


def _default_xy_align(traj_est: np.ndarray, traj_ref: np.ndarray) -> np.ndarray:
    """align_xy_plane_traj (:167-212): both trajectories flattened to z = 1, rotation of the similarity fit est -> ref."""
    est, ref = traj_est[:, :3].copy(), traj_ref[:, :3].copy()
    est[:, 2] = 1
    ref[:, 2] = 1
    try:
        from evo.core import geometry  # type: ignore
        r, _, _ = geometry.umeyama_alignment(est.T, ref.T, True)
    except ImportError:
        r, _, _ = umeyama_alignment(est.T, ref.T, True)
    return r


class HeadNormalFormer(_SeqNetModule):
    """GravityNet (egoego/model/head_normal_estimation_transformer.py:64-294): SLAM head trajectory -> floor normal, then the
    rotation / scale that puts the SLAM trajectory into a gravity-aligned metric frame."""

    _heads = (("action_normal", (512, 256), 3),)

    def __init__(self, opt, device=None, eval_whole_pipeline=False):
        super().__init__()
        self.opt = opt
        self.device = device
        pre = "normal_" if eval_whole_pipeline else ""
        get = lambda k: getattr(opt, pre + k)
        self.transformer_window_size = get("window")
        self._init_net(6 + 3 + 6 + 3, get("d_model"), get("n_dec_layers"), get("n_head"), get("d_k"), get("d_v"), get("window"))
        self.action_normal_mlp = _MLPParams(get("d_model"), self._heads[0][1])
        self.action_normal_fc = nn.Linear(256, 3)

    @torch.no_grad()
    def forward(self, data):
        """:118-165.  data: 'head_rot_mat' [B,P,3,3], 'head_trans' [B,P,3] -> {'pred_normal': [B,3]}."""
        dev = self._cuda_device()
        rot, trans = _f32(data["head_rot_mat"], dev), _f32(data["head_trans"], dev)
        B, P = trans.shape[:2]
        n_pose = min(P, self.transformer_window_size + 1)          # sequences longer than the window are cut (:123-126)
        if n_pose < 2:
            raise ValueError("need at least two poses")
        feats = torch.empty(B, n_pose - 1, 18, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_slam_features(dev.index or 0, rot.data_ptr(), trans.data_ptr(), B, P, n_pose, feats.data_ptr(), _stream(dev)))
        _, (normal,) = self._net_forward(feats, token0_only=True)
        out = defaultdict(list)
        out["pred_normal"] = normal
        return out

    @torch.no_grad()
    def apply_normal_and_scale(self, pred_normal, scale, rot_mat, trans):
        """:219-250: -> (trans_after_rot_scale [B,P,3], aligned_slam_rot_mat [B,P,3,3], aligned_slam_quat [B,P,4], R_align [B,3,3])."""
        dev = self._cuda_device()
        nrm, rot, tr = _f32(pred_normal, dev), _f32(rot_mat, dev), _f32(trans, dev)
        B, P = tr.shape[:2]
        sc = _f32(scale, dev).reshape(-1).expand(B).contiguous()
        to, ro, qo, ao = (torch.empty(B, P, 3, device=dev), torch.empty(B, P, 3, 3, device=dev), torch.empty(B, P, 4, device=dev),
                          torch.empty(B, 3, 3, device=dev))
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_apply_floor_normal(dev.index or 0, nrm.data_ptr(), sc.data_ptr(), rot.data_ptr(), tr.data_ptr(), B, P,
                                                        to.data_ptr(), ro.data_ptr(), qo.data_ptr(), ao.data_ptr(), _stream(dev)))
        return to, ro, qo, ao

    @torch.no_grad()
    def forward_for_eval(self, data, pred_scale=None, use_gt_aligned_rot=False,
                         xy_align: Optional[Callable[[np.ndarray, np.ndarray], np.ndarray]] = None):
        """:214-294 (batch size 1).  ``xy_align(traj_est[T,7], traj_ref[T,7]) -> 3x3`` replaces evo's fit (see module doc)."""
        dev = self._cuda_device()
        if use_gt_aligned_rot:
            raise NotImplementedError("use_gt_aligned_rot (an evaluation upper bound) is not part of the B200 path")
        normal = self.forward(data)["pred_normal"]
        if normal.shape[0] != 1:
            raise ValueError("forward_for_eval expects batch size 1, like the reference (:221)")
        rot, trans = _f32(data["head_rot_mat"], dev), _f32(data["head_trans"], dev)
        scale = _f32(pred_scale, dev).reshape(1) if pred_scale is not None else _f32(data["aligned_scale"], dev).reshape(1)
        trans_after, arm, aq, _ = self.apply_normal_and_scale(normal, scale, rot, trans)
        ref = _f32(data["ori_head_pose"], dev)
        est_np = torch.cat((trans_after, aq), dim=-1)[0].cpu().numpy()
        ref_np = ref[0].cpu().numpy()
        if est_np.shape[0] > ref_np.shape[0]:
            est_np = est_np[:ref_np.shape[0]]
        rxy = torch.from_numpy(np.asarray((xy_align or _default_xy_align)(est_np, ref_np), dtype=np.float32)).to(dev).reshape(1, 3, 3).contiguous()
        P = trans_after.shape[1]
        off = ref[:, 0, :3].contiguous()
        t2, r2, q2 = torch.empty(1, P, 3, device=dev), torch.empty(1, P, 3, 3, device=dev), torch.empty(1, P, 4, device=dev)
        with torch.cuda.device(dev):
            check(_capi.lib().egoego_rigid_apply(dev.index or 0, rxy.data_ptr(), off.data_ptr(), arm.data_ptr(), trans_after.data_ptr(), 1, P,
                                                 t2.data_ptr(), r2.data_ptr(), q2.data_ptr(), _stream(dev)))
        out = defaultdict(list)
        out["head_trans"], out["head_rot_mat"] = t2, r2
        out["head_pose"] = torch.cat((t2, q2), dim=-1)
        out["gt_head_trans"] = ref[:, :, :3]
        out["gt_head_pose"] = ref
        return out
