"""Device-side mirror of the reference's post-sampling evaluation.

``compute_metrics_for_smpl`` keeps the name, argument order and result keys of
kinpoly/scripts/eval_metrics_imu_rec.py:264-342 (what eval_stage2.py:192 calls per generated sequence) but runs as
ONE CUDA kernel (csrc/metrics.cuh, through the C ABI ``egoego_eval_metrics``) on tensors that are already on the GPU --
the reference moves every sequence to numpy and loops over frames in Python.  ``compute_metrics_batch`` evaluates
B sequences in one launch.  No CPU fallback: CPU tensors raise ``EgoEgoError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from ._capi import EgoEgoError, check, lib

KEYS = ("root_trans_dist", "accel_pred", "accel_gt", "accel_err", "pred_fs", "gt_fs", "head_trans_dist", "root_dist",
        "root_rot_dist", "mpjpe", "mpjpe_wo_hand", "head_dist", "head_rot_dist")
N_OUT = 35


def _dev32(t, dev, what):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if dev.type != "cuda":
        raise EgoEgoError(f"{what}: egoego_release_b200 evaluates on the GPU only (no CPU fallback); got a {dev.type} tensor")
    return t.to(device=dev, dtype=torch.float32).contiguous()


def compute_metrics_batch(gt_global_quat, gt_global_jpos, gt_floor_height, pred_global_quat, pred_global_jpos,
                          pred_floor_height) -> torch.Tensor:
    """[B,T,22,4] / [B,T,22,3] CUDA tensors, floor heights [B] -> float32 [B,35] on the same device
    (13 scalars in ``KEYS`` order, then single_jpe[22])."""
    dev = pred_global_jpos.device if torch.is_tensor(pred_global_jpos) else torch.device("cpu")
    pj = _dev32(pred_global_jpos, dev, "pred_global_jpos")
    if pj.dim() != 4 or pj.shape[2:] != (22, 3):
        raise ValueError("joint positions must be [B,T,22,3]")
    B, T = pj.shape[:2]
    gj = _dev32(gt_global_jpos, dev, "gt_global_jpos")
    gq, pq = _dev32(gt_global_quat, dev, "gt_global_quat"), _dev32(pred_global_quat, dev, "pred_global_quat")
    if gj.shape != pj.shape or gq.shape != (B, T, 22, 4) or pq.shape != (B, T, 22, 4):
        raise ValueError("gt / pred shapes differ or quaternions are not [B,T,22,4]")
    if T < 3:
        raise ValueError("need at least 3 frames (accelerations)")
    gf = _dev32(gt_floor_height, dev, "gt_floor_height").reshape(-1)
    pf = _dev32(pred_floor_height, dev, "pred_floor_height").reshape(-1)
    if gf.numel() != B or pf.numel() != B:
        raise ValueError("one floor height per sequence expected")
    out = torch.empty(B, N_OUT, device=dev, dtype=torch.float32)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib().egoego_eval_metrics(dev.index or 0, gq.data_ptr(), gj.data_ptr(), gf.data_ptr(), pq.data_ptr(), pj.data_ptr(),
                                    pf.data_ptr(), B, T, out.data_ptr(), stream))
    return out


def compute_metrics_for_smpl(gt_global_quat, gt_global_jpos, gt_floor_height, pred_global_quat, pred_global_jpos,
                             pred_floor_height) -> Dict[str, float]:
    """Reference signature: T X J X 4 and T X J X 3 tensors of ONE sequence -> dict of python floats with the
    reference's keys (``single_jpe`` is the mean over joints, as the reference's final np.mean leaves it; ``jpe_<i>``
    per joint)."""
    v = compute_metrics_batch(gt_global_quat[None], gt_global_jpos[None], torch.as_tensor([float(gt_floor_height)]),
                              pred_global_quat[None], pred_global_jpos[None], torch.as_tensor([float(pred_floor_height)]))[0]
    v = v.double().cpu()
    res = {k: float(v[i]) for i, k in enumerate(KEYS)}
    single = v[len(KEYS):]
    res["single_jpe"] = float(single.mean())
    for i in range(single.numel()):
        res["jpe_%d" % i] = float(single[i])
    return res


def floor_contacts_batch(body_joint_seq, fps: int = 30):
    """[B,T,22,3] CUDA tensor of global joint positions (z up) -> (offset_floor_height [B] float32, contacts [B,T,22] float32,
    discard [B] bool), all on the device, one kernel launch (csrc/floor.cuh)."""
    dev = body_joint_seq.device if torch.is_tensor(body_joint_seq) else torch.device("cpu")
    j = _dev32(body_joint_seq, dev, "body_joint_seq")
    if j.dim() != 4 or j.shape[2:] != (22, 3):
        raise ValueError("joint positions must be [B,T,22,3]")
    B, T = j.shape[:2]
    if T < 2:
        raise IndexError("need at least 2 frames (the reference indexes the last frame difference)")
    fl = torch.empty(B, device=dev, dtype=torch.float32)
    ct = torch.empty(B, T, 22, device=dev, dtype=torch.float32)
    dc = torch.empty(B, device=dev, dtype=torch.int32)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib().egoego_floor_contacts(dev.index or 0, j.data_ptr(), B, T, int(fps), fl.data_ptr(), ct.data_ptr(), dc.data_ptr(), stream))
    return fl, ct, dc.bool()


def determine_floor_height_and_contacts(body_joint_seq, fps):
    """Reference signature (utils/data_utils/process_amass_dataset.py:160): ONE sequence N x 22 x 3 (CUDA tensor; the reference
    takes the numpy copy of it) -> (offset_floor_height: float, contacts: N x 22 numpy array, discard_seq: bool)."""
    if not torch.is_tensor(body_joint_seq) or body_joint_seq.device.type != "cuda":
        raise EgoEgoError("determine_floor_height_and_contacts: egoego_release_b200 evaluates on the GPU only (no CPU fallback); "
                          "pass the CUDA tensor instead of its .cpu().numpy() copy")
    fl, ct, dc = floor_contacts_batch(body_joint_seq[None], fps)
    return float(fl[0]), ct[0].double().cpu().numpy(), bool(dc[0])
