"""Batch-sharded sampling across the GPUs of one box.

Windows are independent, so the only communication of the path is ONE all-gather of the finished windows; the
sampling loop itself contains no collective.  Each rank draws its noise from Philox streams keyed by the GLOBAL
window id (``window_offset``), so the gathered result equals the single-GPU run of the global batch window for window.
One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) for the plumbing."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, start+count) of ``global_batch`` windows for ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def sample_sharded(sample_fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor], x_start: torch.Tensor,
                   cond_mask: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                   post_fn: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None) -> torch.Tensor:
    """``x_start`` / ``cond_mask`` hold the GLOBAL batch on every rank; each rank samples its shard with
    ``sample_fn(x_shard, mask_shard, window_offset)`` and all ranks return the full [B, T, D] result.

    ``post_fn(local_windows, window_offset) -> Tensor[count, ...]`` (optional) runs on the rank's own windows BEFORE the
    gather, so the single collective moves the post-processed result (e.g. the SMPL parameters of ``smpl_post_fn``: 69 floats
    per frame instead of 198 -- BASELINE configs[2])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x_start.shape[0]
    start, count = shard_range(B, world, rank)
    local = sample_fn(x_start[start:start + count], cond_mask[start:start + count], start)
    if post_fn is not None:
        local = post_fn(local, start)
        if local.shape[0] != count:
            raise ValueError(f"post_fn must keep one row per window: got {local.shape[0]} for {count} windows")
    if world == 1:
        return local
    if B % world == 0:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged shards: pad to the largest shard, gather, then strip
    mx = shard_range(B, world, 0)[1]
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:count] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:shard_range(B, world, r)[1]] for r in range(world)], dim=0)


def model_sample_fn(model) -> Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor]:
    """Adapter: CondGaussianDiffusion.sample with Philox streams keyed by the global window id."""
    def fn(xs, cm, offset):
        model.window_offset = int(offset)
        try:
            return model.sample(xs, cm)
        finally:
            model.window_offset = 0
    return fn


def smpl_post_fn(model, ds, recover_rot_quat=None) -> Callable[[torch.Tensor, int], torch.Tensor]:
    """Adapter for ``sample_sharded(post_fn=...)``: convert_model_res_to_data on the rank's windows (device kernel), packed as
    [count, T, 69] = local joint axis-angles (22 x 3) followed by the root translation (3) -- the "generated SMPL params" that
    BASELINE configs[2] gathers.  ``recover_rot_quat`` is indexed by GLOBAL window id ([B, 1, 1, 4] or [B, 4]); None = identity."""
    def fn(windows, offset):
        rq = None
        if recover_rot_quat is not None:
            rq = torch.as_tensor(recover_rot_quat).reshape(-1, 4)[offset:offset + windows.shape[0]]
        aa, root, _head = model.postprocess(ds, windows, rq)
        return torch.cat((aa.reshape(aa.shape[0], aa.shape[1], 66), root), dim=-1)
    return fn


def unpack_smpl(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B, T, 69] from ``smpl_post_fn`` -> (local_aa [B, T, 22, 3], root_trans [B, T, 3])."""
    return packed[..., :66].reshape(packed.shape[0], packed.shape[1], 22, 3), packed[..., 66:]


def average_flat_gradients(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Data-parallel training without a DDP wrapper (CondGaussianDiffusion.set_grad_sync): ONE all-reduce of the flat buffer that
    holds every gradient of the step, divided by the group size (DDP's mean).  No-op outside an initialised process group or in a
    group of one.  In place; returns ``flat``."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(world)
    return flat

