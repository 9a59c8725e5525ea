"""TEST INFRASTRUCTURE (CPU oracle) -- restatement of EgoEgo's stage-1 networks in the shipped configuration
(scripts/test_egoego_pipeline.sh: --input_of_feats, HeadNet window 60, GravityNet window 120, d_model 256, 2 layers,
4 heads x 256).  torch CPU fp32; every function cites the reference lines it follows.

  HeadFormer (HeadNet)         egoego/model/head_estimation_transformer.py
      forward_for_eval          :214-308   blocks of `window` frames -> Decoder -> two MLP heads -> va2rot -> SLAM re-scaling
      va2rot                    :102-124   angular-velocity integration (quaternion_apply / axis_angle_to_quaternion / multiply)
      cal_scale_for_slam_w_pred_scale :184-212
  HeadNormalFormer (GravityNet) egoego/model/head_normal_estimation_transformer.py
      forward                   :118-165   SLAM features (rot6d, trans, frame-to-frame diffs) -> Decoder -> MLP on token 0 -> normal
      rotation_matrix_from_two_vectors / cal_rotation_from_floor_normal :47-62 (numpy float64)
      forward_for_eval          :214-250   rotation + scale applied to the SLAM trajectory (the part BEFORE the xy-plane
                                           alignment, which is evo's Umeyama fit -- third-party, absent, not restated)
  Decoder / DecoderLayer / MLP  egoego/model/transformer_module.py:119-142,172-226; egoego/model/mlp.py:4-27
      (use_full_attention=True: no time mask; padded positions ARE attended to -- their rows are only zeroed after
       every sub-layer by the padding mask)

Pinned by tests/golden/stage1.npz (oracle/gen_golden_stage1.py runs the UNMODIFIED reference modules behind import stubs).
pytorch3d.transforms is third-party and absent: served by oracle/rotations.py ("parity unpinned" for those functions).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import rotations as R
from .egoego_oracle import _uniform, ffn, mha, sinusoid_table

CFG_HEAD = dict(kind="head", d_feats=512, d_model=256, n_head=4, n_dec_layers=2, d_k=256, d_v=256, window=60,
                heads={"action_va": ([1024, 512, 256], 3), "action_dist": ([1024, 512, 256], 1)}, dist_scale=10.0)
CFG_NORMAL = dict(kind="normal", d_feats=18, d_model=256, n_head=4, n_dec_layers=2, d_k=256, d_v=256, window=120,
                  heads={"action_normal": ([512, 256], 3)})


def init_params(seed: int, cfg: Dict) -> Dict[str, torch.Tensor]:
    """Seeded weights with the reference modules' state_dict names / shapes / init scales (bit-reproducible on any host)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d, H, dk, dv, D = cfg["d_model"], cfg["n_head"], cfg["d_k"], cfg["d_v"], cfg["d_feats"]
    p: Dict[str, torch.Tensor] = {}

    def default_linear(name, out_f, in_f, conv=False):
        bound = 1.0 / math.sqrt(in_f)
        w = _uniform(rng, (out_f, in_f), bound / math.sqrt(3.0))
        p[name + ".weight"] = w[:, :, None].contiguous() if conv else w
        p[name + ".bias"] = _uniform(rng, (out_f,), bound / math.sqrt(3.0))

    pre = "action_transformer."
    default_linear(pre + "start_conv", d, D, conv=True)
    p[pre + "position_vec.weight"] = sinusoid_table(cfg["window"] + 1, d)
    for l in range(cfg["n_dec_layers"]):
        a = f"{pre}layer_stack.{l}.self_attn."
        for nm, dd in (("w_q", dk), ("w_k", dk), ("w_v", dv)):
            p[a + nm + ".weight"] = _uniform(rng, (H * dd, d), math.sqrt(2.0 / (d + dd)))
            p[a + nm + ".bias"] = _uniform(rng, (H * dd,), 1.0 / math.sqrt(3.0 * d))
        p[a + "fc.weight"] = _uniform(rng, (d, H * dv), math.sqrt(2.0 / (d + H * dv)))
        p[a + "fc.bias"] = _uniform(rng, (d,), 1.0 / math.sqrt(3.0 * H * dv))
        p[a + "layer_norm.weight"] = 1.0 + _uniform(rng, (d,), 0.1)
        p[a + "layer_norm.bias"] = _uniform(rng, (d,), 0.1)
        f = f"{pre}layer_stack.{l}.pos_ffn."
        default_linear(f + "w_1", d, d, conv=True)
        default_linear(f + "w_2", d, d, conv=True)
        p[f + "layer_norm.weight"] = 1.0 + _uniform(rng, (d,), 0.1)
        p[f + "layer_norm.bias"] = _uniform(rng, (d,), 0.1)
    for head, (hidden, out) in cfg["heads"].items():
        last = d
        for i, nh in enumerate(hidden):
            default_linear(f"{head}_mlp.affine_layers.{i}", nh, last)
            last = nh
        default_linear(f"{head}_fc", out, last)
    return p


# ---- Decoder without a leading token (transformer_module.py:188-226) ---------------------------------------------------
def decoder_forward(p, x: torch.Tensor, padding_mask: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """x [B, window, d_feats] (already zero-padded to the window), padding_mask [B, window] (1 = real frame)."""
    pre = "action_transformer."
    B, L, _ = x.shape
    h = F.linear(x, p[pre + "start_conv.weight"][:, :, 0], p[pre + "start_conv.bias"])
    h = h + p[pre + "position_vec.weight"][1:L + 1][None]            # pos_vec = 1..window for every position, padded or not
    pm = padding_mask.reshape(B, L, 1).float()
    for l in range(cfg["n_dec_layers"]):
        h = mha(p, f"{pre}layer_stack.{l}.self_attn.", h, cfg["n_head"], cfg["d_k"], cfg["d_v"]) * pm
        h = ffn(p, f"{pre}layer_stack.{l}.pos_ffn.", h) * pm
    return h


def mlp_head(p, head: str, x: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """MLP (mlp.py:22-25: activation after EVERY affine layer, relu) followed by the head's Linear."""
    for i in range(len(cfg["heads"][head][0])):
        x = torch.relu(F.linear(x, p[f"{head}_mlp.affine_layers.{i}.weight"], p[f"{head}_mlp.affine_layers.{i}.bias"]))
    return F.linear(x, p[f"{head}_fc.weight"], p[f"{head}_fc.bias"])


def pad_window(x: torch.Tensor, window: int):
    """[B, T<=window, D] -> zero-padded [B, window, D] and the padding mask [B, window]."""
    B, T, D = x.shape
    out = torch.zeros(B, window, D, dtype=x.dtype, device=x.device)
    out[:, :T] = x
    mask = (torch.arange(window, device=x.device)[None, :] < T).expand(B, window)
    return out, mask


# ---- HeadFormer ------------------------------------------------------------------------------------------------------
def va2rot(curr_rot: torch.Tensor, vels: torch.Tensor, dt: float = 1 / 30) -> torch.Tensor:
    """head_estimation_transformer.py:102-124: q_{t+1} = normalise(quat(aa = (q_t * w_t) dt) * q_t);  [B,4],[B,T,3] -> [B,T+1,4]."""
    seq = [curr_rot]
    for t in range(vels.shape[1]):
        angv = R.quaternion_apply(curr_rot.float(), vels[:, t, :].float())
        new_rot = R.quaternion_multiply(R.axis_angle_to_quaternion(angv * dt), curr_rot)
        curr_rot = new_rot / torch.norm(new_rot, dim=1).reshape(-1, 1)
        seq.append(curr_rot)
    return torch.stack(seq, dim=1)


def cal_scale_for_slam_w_pred_scale(slam_trans: torch.Tensor, dist_scalar: torch.Tensor):
    """:184-212.  slam_trans [(T+1),3], dist_scalar [T'] -> rescaled [(T+1),3], scale."""
    n = slam_trans.shape[0] - 1
    lens = torch.linalg.norm(slam_trans[1:] - slam_trans[:-1], dim=-1)        # per-step SLAM displacement
    if n < dist_scalar.shape[0]:
        dist_scalar = dist_scalar[:n]
    elif n > dist_scalar.shape[0]:
        lens = lens[:dist_scalar.shape[0]]
    scale = dist_scalar.mean() / lens.mean()
    out = [slam_trans[0]]
    for t in range(n):
        out.append(out[-1] + scale * (slam_trans[t + 1] - slam_trans[t]))
    return torch.stack(out), scale


def headformer_blocks(p, feats: torch.Tensor, cfg: Dict = CFG_HEAD):
    """The network part of forward_for_eval (:232-262): per block of `window` frames -> (va [B,T,3], dist [B,T,1])."""
    W = cfg["window"]
    T = feats.shape[1]
    vas, dists = [], []
    for b in range(T // W + 1):
        cur = feats[:, b * W:(b + 1) * W]
        if cur.shape[1] == 0:
            continue
        n = cur.shape[1]
        x, mask = pad_window(cur.float(), W)
        h = decoder_forward(p, x, mask, cfg)[:, :n]
        vas.append(mlp_head(p, "action_va", h, cfg))
        dists.append(mlp_head(p, "action_dist", h, cfg))
    return vas, dists


def headformer_forward_for_eval(p, feats: torch.Tensor, aligned_slam_trans: torch.Tensor, head_q0: torch.Tensor,
                                cfg: Dict = CFG_HEAD):
    """:214-308.  feats [1,T,512], aligned_slam_trans [1,T+1,3], head_q0 [1,4] (wxyz) -> head_pose [1,T'',7], pred_scale."""
    vas, dists = headformer_blocks(p, feats, cfg)
    quats, prev = [], None
    for i, va in enumerate(vas):
        cur = va2rot(head_q0 if i == 0 else prev, va)
        quats.append(cur if i == 0 else cur[:, 1:])
        prev = cur[:, -1]
    quat = torch.cat(quats, dim=1)
    dist = torch.cat(dists, dim=1) / cfg["dist_scale"]
    trans, scale = cal_scale_for_slam_w_pred_scale(aligned_slam_trans[0], dist[0].squeeze(-1))
    if trans.shape[0] != quat.shape[1]:
        quat = quat[:, :trans.shape[0]]
    return torch.cat((trans[None], quat), dim=-1), scale


# ---- HeadNormalFormer ------------------------------------------------------------------------------------------------
def slam_features(rot_mat: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    """head_normal_estimation_transformer.py:128-137: [B,T+1,3,3],[B,T+1,3] -> [B,T,18] = rot6d | trans | rot6d(R_{t+1} R_t^T) | dtrans."""
    r6 = R.matrix_to_rotation_6d(rot_mat)
    diff = torch.matmul(rot_mat[:, 1:], rot_mat[:, :-1].transpose(2, 3))
    return torch.cat((r6[:, :-1], trans[:, :-1], R.matrix_to_rotation_6d(diff), trans[:, 1:] - trans[:, :-1]), dim=-1)


def headnormal_forward(p, rot_mat: torch.Tensor, trans: torch.Tensor, cfg: Dict = CFG_NORMAL) -> torch.Tensor:
    """:118-165 (eval, batch of independent sequences): sequences longer than the window are cut to window+1 poses."""
    W = cfg["window"]
    if trans.shape[1] > W:
        rot_mat, trans = rot_mat[:, :W + 1], trans[:, :W + 1]
    x, mask = pad_window(slam_features(rot_mat.float(), trans.float()), W)
    h = decoder_forward(p, x, mask, cfg)
    return mlp_head(p, "action_normal", h[:, 0, :], cfg)


def rotation_matrix_from_two_vectors(vec1, vec2):
    """:47-57 (numpy float64, Rodrigues form)."""
    a, b = (vec1 / np.linalg.norm(vec1)).reshape(3), (vec2 / np.linalg.norm(vec2)).reshape(3)
    v = np.cross(a, b)
    c = np.dot(a, b)
    s = np.linalg.norm(v)
    k = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + k + k.dot(k) * ((1 - c) / (s ** 2))


def cal_rotation_from_floor_normal(normal):
    return rotation_matrix_from_two_vectors(normal, np.asarray([0, 0, 1]))


def apply_normal_and_scale(pred_normal: torch.Tensor, scale: torch.Tensor, rot_mat: torch.Tensor, trans: torch.Tensor):
    """forward_for_eval :219-250 up to (not including) the evo xy-plane alignment: rotation that takes the predicted floor
    normal to +z, applied with the predicted scale to the frame-to-frame SLAM translations (re-integrated from frame 0) and
    to the SLAM rotations.  pred_normal [1,3], scale scalar, rot_mat [1,T,3,3], trans [1,T,3]
    -> trans_after [1,T,3], aligned_rot_mat [1,T,3,3], aligned_quat [1,T,4]."""
    Ra = torch.from_numpy(cal_rotation_from_floor_normal(pred_normal[0].double().numpy())).float()
    rot_mat, trans = rot_mat.float(), trans.float()
    d = trans[:, 1:] - trans[:, :-1]
    d = torch.matmul(Ra[None, None].expand(d.shape[0], d.shape[1], 3, 3), d[:, :, :, None]).squeeze(-1) * scale.reshape(1, 1, 1)
    out = [trans[:, 0]]
    for t in range(d.shape[1]):
        out.append(out[-1] + d[:, t])
    out = torch.stack(out, dim=1)
    arm = torch.matmul(Ra[None, None].expand(rot_mat.shape[0], rot_mat.shape[1], 3, 3), rot_mat)
    return out, arm, R.matrix_to_quaternion(arm)


def synth_stage1_inputs(seed: int, T: int):
    """Seeded stage-1 inputs of ONE sequence: optical-flow features [1,T,512], SLAM/GT head trajectory [1,T+1,7]
    (random-walk position, yaw-dominant random-walk orientation), aligned SLAM translation (scaled copy)."""
    rng = np.random.default_rng(seed)
    feats = rng.normal(0, 1, (1, T, 512)).astype(np.float32)
    pos = np.cumsum(rng.normal(0, 0.01, (T + 1, 3)), axis=0).astype(np.float32) + np.float32([0, 0, 1.6])
    yaw = np.cumsum(rng.normal(0, 0.03, T + 1))
    tilt = rng.normal(0, 0.05, (T + 1, 2))
    aa = np.stack([tilt[:, 0], tilt[:, 1], yaw], -1).astype(np.float32)
    quat = R.axis_angle_to_quaternion(torch.from_numpy(aa))
    head_pose = torch.cat((torch.from_numpy(pos), quat), -1)[None]
    slam_trans = (torch.from_numpy(pos) - torch.from_numpy(pos[:1])) * 0.37          # SLAM lives at an unknown scale
    tiltq = R.axis_angle_to_quaternion(torch.tensor([[0.2, -0.1, 0.4]]))
    slam_rot = R.quaternion_to_matrix(R.quaternion_multiply(tiltq.expand(T + 1, 4), quat))
    slam_trans = torch.matmul(R.quaternion_to_matrix(tiltq)[0], slam_trans[:, :, None])[:, :, 0]
    return torch.from_numpy(feats), head_pose, slam_trans[None].contiguous(), slam_rot[None].contiguous()


# ---- ResNet-18 optical-flow encoder (HeadFormer with input_of_feats=False) ---------------------------------------------------
RESNET_PLANES = (64, 128, 256, 512)


def init_resnet_params(seed: int, out_dim: int = 512) -> Dict[str, torch.Tensor]:
    """Seeded weights with the state_dict layout of egoego/model/resnet.py's ``ResNet`` (keys ``resnet.*``: torchvision's
    resnet18 with ``fc`` -> out_dim), incl. non-trivial BatchNorm running statistics (a trained checkpoint has them)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        p[name + ".weight"] = _uniform(rng, (cout, cin, k, k), math.sqrt(2.0 / (cin * k * k)))

    def bn(name, c):
        p[name + ".weight"] = 1.0 + _uniform(rng, (c,), 0.1)
        p[name + ".bias"] = _uniform(rng, (c,), 0.1)
        p[name + ".running_mean"] = _uniform(rng, (c,), 0.2)
        p[name + ".running_var"] = 1.0 + _uniform(rng, (c,), 0.2).abs()
        p[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("resnet.conv1", 64, 3, 7); bn("resnet.bn1", 64)
    inpl = 64
    for L, planes in enumerate(RESNET_PLANES):
        for blk in range(2):
            pre = f"resnet.layer{L + 1}.{blk}."
            stride = 2 if (L > 0 and blk == 0) else 1
            conv(pre + "conv1", planes, inpl, 3); bn(pre + "bn1", planes)
            conv(pre + "conv2", planes, planes, 3); bn(pre + "bn2", planes)
            if stride != 1 or inpl != planes:
                conv(pre + "downsample.0", planes, inpl, 1); bn(pre + "downsample.1", planes)
            inpl = planes
    bound = 1.0 / math.sqrt(512)
    p["resnet.fc.weight"] = _uniform(rng, (out_dim, 512), bound / math.sqrt(3.0))
    p["resnet.fc.bias"] = _uniform(rng, (out_dim,), bound / math.sqrt(3.0))
    return p


def resnet18_forward(p, x: torch.Tensor) -> torch.Tensor:
    """torchvision resnet18 in eval mode (BatchNorm with running statistics), as egoego/model/resnet.py:16-17 runs it.
    x [N,3,224,224] -> [N,out_dim]."""
    def bn(name, h):
        return F.batch_norm(h, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"], p[name + ".bias"], False, 0.0, 1e-5)

    h = F.relu(bn("resnet.bn1", F.conv2d(x, p["resnet.conv1.weight"], None, 2, 3)))
    h = F.max_pool2d(h, 3, 2, 1)
    inpl = 64
    for L, planes in enumerate(RESNET_PLANES):
        for blk in range(2):
            pre = f"resnet.layer{L + 1}.{blk}."
            stride = 2 if (L > 0 and blk == 0) else 1
            idt = h
            o = F.relu(bn(pre + "bn1", F.conv2d(h, p[pre + "conv1.weight"], None, stride, 1)))
            o = bn(pre + "bn2", F.conv2d(o, p[pre + "conv2.weight"], None, 1, 1))
            if stride != 1 or inpl != planes:
                idt = bn(pre + "downsample.1", F.conv2d(h, p[pre + "downsample.0.weight"], None, stride, 0))
            h = F.relu(o + idt)
            inpl = planes
    h = F.adaptive_avg_pool2d(h, 1).flatten(1)
    return F.linear(h, p["resnet.fc.weight"], p["resnet.fc.bias"])


def flow_to_cnn_input(of: torch.Tensor) -> torch.Tensor:
    """head_estimation_transformer.py:218-223: [B,T,224,224,2] -> zero third channel -> [B*T,3,224,224]."""
    of = torch.cat((of, torch.zeros(of.shape[:-1] + (1,))), dim=-1)
    return of.reshape(-1, 224, 224, 3).permute(0, 3, 1, 2)


def synth_flow(seed: int, T: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.normal(0, 1, (1, T, 224, 224, 2)).astype(np.float32))
