"""End-to-end EgoEgo inference on one B200 (BASELINE config 4): the body of ``run_egoego.test()`` (run_egoego.py:95-171)
with every numeric step in libegoego_b200:

    HeadFormer.forward_for_eval        -> head rotations + SLAM scale                    (run_egoego.py:102-104)
    HeadNormalFormer.forward_for_eval  -> gravity-aligned, metric head translations      (:106-115)
    head pose assembly / floor offset  -> conditioning of stage 2                        (:119-135)
    full_body_gen_cond_head_pose_sliding_window -> local axis-angle + root translation   (:149-151)
    fk_smpl, head re-centring          -> global joint rotations / positions             (:153-166)

Out of scope here exactly as in DESIGN.md: data loading (ARESDemoDataset), floor-height estimation
(determine_floor_height_and_contacts: sklearn DBSCAN on the host), mesh export / Blender.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .trainer_glue import full_body_gen_cond_head_pose_sliding_window


@torch.no_grad()
def estimate_head_pose(head_net, gravity_net, data: Dict, floor_z_offset: float = 0.13,
                       xy_align: Optional[Callable] = None) -> Dict[str, torch.Tensor]:
    """Stage 1 (run_egoego.py:102-135).  ``data`` holds the demo loader's keys: 'of' [1,T,512], 'aligned_slam_trans' [1,T+1,3],
    'head_pose' [1,T+1,7] (GT, only frame 0 is used), 'ori_slam_trans' [1,T+1,3], 'ori_slam_rot_mat' [1,T+1,3,3].
    Returns {'head_pose': [1,T',7] (xyz + wxyz quaternion, float32), 'pred_scale'}.  ``floor_z_offset`` is the reference's
    sequence-specific constant (:135: "This values is only for this sequence")."""
    dev = next(head_net.parameters()).device
    s1 = head_net.forward_for_eval(data)
    pred_scale = s1["pred_scale"]
    ori_trans = data["ori_slam_trans"].to(dev).float()
    normal_in = {"head_trans": ori_trans - ori_trans[:, 0:1, :], "head_rot_mat": data["ori_slam_rot_mat"].to(dev).float(),
                 "ori_head_pose": data["head_pose"].to(dev).float(),
                 "seq_len": torch.tensor(ori_trans.shape[1]).float()[None]}
    s1n = gravity_net.forward_for_eval(normal_in, pred_scale, xy_align=xy_align)
    n = min(s1n["head_pose"].shape[1], s1["head_pose"].shape[1])
    head_pose = torch.cat((s1n["head_pose"][:, :n, :3], s1["head_pose"][:, :n, 3:]), dim=-1).clone()
    head_pose[0, :, :2] -= head_pose[0, 0:1, :2].clone()
    move_to_floor = data["head_pose"][0, 0:1, :3].to(dev).float() - head_pose[0, 0:1, :3]
    head_pose[0, :, :3] += move_to_floor
    head_pose[0, :, 2] -= floor_z_offset
    return {"head_pose": head_pose, "pred_scale": pred_scale}


@torch.no_grad()
def generate_full_body(diffusion, ds, head_pose: torch.Tensor, sample_bs: int = 1, noise_fn=None) -> Dict[str, torch.Tensor]:
    """Stage 2 + FK (run_egoego.py:143-166): head_pose [1,T,7] -> local axis-angle [BS,T',22,3], root translation, FK joints."""
    rep = head_pose.repeat(sample_bs, 1, 1)
    aa, root = full_body_gen_cond_head_pose_sliding_window(diffusion, ds, rep, noise_fn=noise_fn)
    jrot, jpos = ds.fk_smpl(root.reshape(-1, 3), aa.reshape(-1, 22, 3))
    jrot = jrot.reshape(sample_bs, -1, 22, 4)
    jpos = jpos.reshape(sample_bs, -1, 22, 3)
    move = jpos[:, 0:1, 15:16, :].clone()            # put the first frame's head at x = y = 0 (:160-163)
    move[:, :, :, 2] *= 0
    jpos = jpos - move
    return {"local_aa": aa, "root_trans": jpos[:, :, 0, :].clone(), "global_jrot": jrot, "global_jpos": jpos}


@torch.no_grad()
def run_egoego(head_net, gravity_net, diffusion, ds, data: Dict, sample_bs: int = 1, floor_z_offset: float = 0.13,
               xy_align: Optional[Callable] = None, noise_fn=None) -> Dict[str, torch.Tensor]:
    """The whole pipeline of run_egoego.test() for one sequence; returns stage-1 and stage-2 results in one dict."""
    s1 = estimate_head_pose(head_net, gravity_net, data, floor_z_offset, xy_align)
    out = generate_full_body(diffusion, ds, s1["head_pose"], sample_bs, noise_fn)
    out.update(s1)
    return out
