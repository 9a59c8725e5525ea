"""The three Trainer helpers that sit directly on the sampler boundary
(trainer_amass_cond_motion_diffusion.py:210-231,261-277), as free functions so scripts without the
reference Trainer (wandb / ema_pytorch / AMASS) can drive the same calls."""
from __future__ import annotations

import torch


def prep_head_condition_mask(data: torch.Tensor, joint_idx: int = 15) -> torch.Tensor:
    """1 = missing, 0 = conditioned: head position [45:48] and head rot6d [156:162] (reference :210-221)."""
    mask = torch.ones_like(data)
    mask[:, :, joint_idx * 3:joint_idx * 3 + 3] = 0
    mask[:, :, 22 * 3 + joint_idx * 6:22 * 3 + joint_idx * 6 + 6] = 0
    return mask


def prep_padding_mask(val_data: torch.Tensor, seq_len: torch.Tensor, window: int = 120) -> torch.Tensor:
    """BS x 1 x (window+1) bool, True on the time token and the first seq_len frames (reference :223-231)."""
    actual = seq_len + 1
    m = torch.arange(window + 1, device=val_data.device).expand(val_data.shape[0], window + 1) < \
        actual.to(val_data.device)[:, None].repeat(1, window + 1)
    return m[:, None, :]


@torch.no_grad()
def full_body_gen_cond_head_pose_sliding_window(model, ds, head_pose: torch.Tensor, noise_fn=None):
    """head_pose BS x T x 7 (xyz + wxyz) -> (local axis-angle BS x T' x 22 x 3, root BS x T' x 3) (reference :261-277)."""
    global_head_jpos = head_pose[:, :, :3]
    global_head_quat = head_pose[:, :, 3:]
    data = torch.zeros(head_pose.shape[0], head_pose.shape[1], 22 * 3 + 22 * 6, device=head_pose.device)
    cond_mask = prep_head_condition_mask(data)
    return model.sample_sliding_window_w_canonical(ds, global_head_jpos, global_head_quat, x_start=data,
                                                   cond_mask=cond_mask, noise_fn=noise_fn)
