"""egoego_release_b200 -- B200-native (sm_100a) implementation of EgoEgo's stage-2 conditional
motion-diffusion sampling path behind the reference's own Python boundary.

Public surface (mirrors egoego/model/transformer_cond_diffusion_model.py):
    CondGaussianDiffusion, TransformerDiffusionModel, MotionDataStub, prep_head_condition_mask
Next rows of the path: HeadFormer / HeadNormalFormer (stage-1 networks, egoego/model/head_*_transformer.py) and
compute_metrics_for_smpl (kinpoly/scripts/eval_metrics_imu_rec.py) -- same boundary, same library.
The numeric path lives in lib/libegoego_b200.so (C ABI: include/egoego_b200.h); importing this package
never falls back to PyTorch math -- using it without the built library or without a B200 raises.
"""
from ._capi import EgoEgoError, PRECISE_ALL_FP16  # noqa: F401
from .diffusion import CondGaussianDiffusion, TransformerDiffusionModel  # noqa: F401
from .motion_data import MotionDataStub  # noqa: F401
from .stage1 import HeadFormer, HeadNormalFormer  # noqa: F401
from .eval_metrics import (compute_metrics_for_smpl, compute_metrics_batch, determine_floor_height_and_contacts,  # noqa: F401
                           floor_contacts_batch)
from .trainer_glue import prep_head_condition_mask, prep_padding_mask, full_body_gen_cond_head_pose_sliding_window  # noqa: F401

__all__ = ["PRECISE_ALL_FP16", "CondGaussianDiffusion", "TransformerDiffusionModel", "MotionDataStub", "EgoEgoError", "HeadFormer", "HeadNormalFormer",
           "compute_metrics_for_smpl", "compute_metrics_batch", "determine_floor_height_and_contacts", "floor_contacts_batch",
           "prep_head_condition_mask", "prep_padding_mask", "full_body_gen_cond_head_pose_sliding_window"]
