"""ctypes binding of libegoego_b200.so (declarations mirror include/egoego_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

EXPORTS = [
    "egoego_last_error", "egoego_version", "egoego_create", "egoego_destroy", "egoego_set_tensor",
    "egoego_make_cosine_schedule", "egoego_commit_weights", "egoego_denoiser_forward", "egoego_p_sample_step",
    "egoego_sample", "egoego_sample_host", "egoego_set_skeleton", "egoego_postprocess", "egoego_fk_smpl",
    "egoego_canonicalize_head", "egoego_tail_condition", "egoego_launch_count", "egoego_selftest_gemm", "egoego_time_dominant_kernel", "egoego_time_kernel",
    "egoego_precise_last_steps", "egoego_weight_sets", "egoego_dither_weights_f16", "egoego_eval_metrics", "egoego_launches_per_step",
    "egoego_seqnet_create", "egoego_seqnet_destroy", "egoego_seqnet_set_tensor", "egoego_seqnet_commit", "egoego_seqnet_forward",
    "egoego_seqnet_launch_count", "egoego_va2rot", "egoego_rescale_slam", "egoego_slam_features", "egoego_apply_floor_normal",
    "egoego_rigid_apply", "egoego_resnet18_create", "egoego_resnet18_destroy", "egoego_resnet18_set_tensor", "egoego_resnet18_commit",
    "egoego_resnet18_forward", "egoego_resnet18_launch_count", "egoego_train_step", "egoego_train_get_grad", "egoego_update_tensor_device", "egoego_train_get_grads", "egoego_update_tensors_device",
    "egoego_tensors_checksum", "egoego_floor_contacts", "egoego_train_set_dropout", "egoego_engine_info",
    "egoego_debug_timeline",
]

ENGINE_TCGEN05, ENGINE_SIMT = 0, 1
PRECISE_ALL_FP16 = -2     # egoego_cfg.precise_last_steps: explicit opt-out of the precision policy (EGOEGO_PRECISE_ALL_FP16)


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_feats", "d_model", "n_head", "n_dec_layers", "d_k", "d_v", "max_timesteps", "timesteps",
        "objective", "max_batch", "device", "engine", "precise_last_steps")]


class SeqNetCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("d_feats", "d_model", "n_head", "n_layers", "d_k", "d_v", "window", "max_batch", "device",
                                         "n_heads")] + [("head_n_hidden", C.c_int32 * 2), ("head_hidden", (C.c_int32 * 3) * 2),
                                                        ("head_out", C.c_int32 * 2)]


class Rng(C.Structure):
    _fields_ = [("tape", C.c_void_p), ("seed", C.c_uint64), ("window_offset", C.c_uint64)]


class EgoEgoError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EgoEgoError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  egoego_release_b200 has no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.egoego_last_error.restype = C.c_char_p
    L.egoego_version.restype = i32
    L.egoego_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    L.egoego_destroy.argtypes = [vp]
    L.egoego_set_tensor.argtypes = [vp, C.c_char_p, vp, i64, i32]
    L.egoego_make_cosine_schedule.argtypes = [vp]
    L.egoego_commit_weights.argtypes = [vp, vp]
    L.egoego_denoiser_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    L.egoego_p_sample_step.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Rng), u64, vp, i32, vp, i32, i32, i32, vp, vp]
    L.egoego_sample.argtypes = [vp, vp, vp, i32, i32, C.POINTER(Rng), vp, vp, i32, vp, vp]
    L.egoego_sample_host.argtypes = [vp, vp, vp, i32, i32, C.POINTER(Rng), vp, vp]
    L.egoego_set_skeleton.argtypes = [vp, vp, vp, vp, vp]
    L.egoego_postprocess.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.egoego_fk_smpl.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.egoego_canonicalize_head.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp]
    L.egoego_selftest_gemm.argtypes = [i32, i32, i32, i32, u64, i32, i32, vp, vp, vp]
    L.egoego_time_dominant_kernel.argtypes = [vp, i32, i32, i32, vp, vp]
    L.egoego_time_kernel.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    L.egoego_precise_last_steps.argtypes = [vp]
    L.egoego_engine_info.argtypes = [vp, C.c_char_p, i32]
    L.egoego_debug_timeline.argtypes = [i32, i32, vp, i32]
    L.egoego_weight_sets.argtypes = [vp]
    L.egoego_dither_weights_f16.argtypes = [vp, C.c_int64, i32, i32, vp]
    L.egoego_launches_per_step.argtypes = [vp, i32]
    L.egoego_tail_condition.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.egoego_eval_metrics.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.egoego_launch_count.argtypes = [vp]
    L.egoego_floor_contacts.argtypes = [i32, vp, i32, i32, i32, vp, vp, vp, vp]
    L.egoego_tensors_checksum.argtypes = [i32, i32, vp, vp, vp, vp]
    L.egoego_launch_count.restype = i64
    f32 = C.c_float
    L.egoego_seqnet_create.argtypes = [C.POINTER(SeqNetCfg), C.POINTER(vp)]
    L.egoego_seqnet_destroy.argtypes = [vp]
    L.egoego_seqnet_destroy.restype = None
    L.egoego_seqnet_set_tensor.argtypes = [vp, C.c_char_p, vp, i64]
    L.egoego_seqnet_commit.argtypes = [vp]
    L.egoego_seqnet_forward.argtypes = [vp, vp, i32, i32, vp, vp, vp, i32, vp]
    L.egoego_seqnet_launch_count.argtypes = [vp]
    L.egoego_seqnet_launch_count.restype = i64
    L.egoego_va2rot.argtypes = [i32, vp, vp, i32, i32, f32, vp, vp]
    L.egoego_rescale_slam.argtypes = [i32, vp, i32, vp, i32, f32, vp, vp, vp]
    L.egoego_slam_features.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp]
    L.egoego_apply_floor_normal.argtypes = [i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.egoego_rigid_apply.argtypes = [i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.egoego_train_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    L.egoego_train_get_grad.argtypes = [vp, C.c_char_p, vp, i64, vp]
    L.egoego_train_set_dropout.argtypes = [vp, C.c_double, u64]
    L.egoego_update_tensor_device.argtypes = [vp, C.c_char_p, vp, i64, vp]
    L.egoego_train_get_grads.argtypes = [vp, i32, vp, vp, vp, vp]
    L.egoego_update_tensors_device.argtypes = [vp, i32, vp, vp, vp, vp]
    L.egoego_resnet18_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.egoego_resnet18_destroy.argtypes = [vp]
    L.egoego_resnet18_destroy.restype = None
    L.egoego_resnet18_set_tensor.argtypes = [vp, C.c_char_p, vp, i64]
    L.egoego_resnet18_commit.argtypes = [vp]
    L.egoego_resnet18_forward.argtypes = [vp, vp, i32, vp, vp]
    L.egoego_resnet18_launch_count.argtypes = [vp]
    L.egoego_resnet18_launch_count.restype = i64
    for name in EXPORTS:
        if name not in ("egoego_last_error", "egoego_launch_count", "egoego_seqnet_destroy", "egoego_seqnet_launch_count",
                        "egoego_resnet18_destroy", "egoego_resnet18_launch_count"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise EgoEgoError(lib().egoego_last_error().decode())


def content_checksum(tensors, device) -> int:
    """64-bit content checksum of CUDA fp32 tensors (egoego_tensors_checksum: one kernel + an 8-byte read-back).
    The mirrors add it to their (data_ptr, _version) signatures: in-place edits through ``.data`` (ema_pytorch's
    ``copy_`` / ``lerp_``, trainer_amass_cond_motion_diffusion.py:179-192) do not bump torch's version counters."""
    import torch
    ts = [t for t in tensors if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() > 0]
    if not ts:
        return 0
    n = len(ts)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    nums = (C.c_int64 * n)(*[t.numel() for t in ts])
    out = C.c_uint64(0)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(device):
        check(lib().egoego_tensors_checksum(idx, n, ptrs, nums, C.byref(out), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return int(out.value)
