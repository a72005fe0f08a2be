"""Multi-GPU plumbing for the hot path: one process per GPU, rays sharded, the neural
point cloud / occupancy grid / weights / reference images replicated, one NCCL all-reduce of the
gradients per training step.  Rendering needs no collective at all (frames / ray blocks are
independent units).

The reference has no multi-GPU path (its querier owns one pycuda context); this is new.

Semantics preserved on every replica:
  * the loss of the reference is a MEAN over the ray-masked rays of the whole batch
    (models/base_rendering_model.py:1114-1118), so a rank's gradients are scaled by
    n_local_valid / n_global_valid before the SUM all-reduce (device-side, no host sync);
  * dense Adam (every point row is updated every step, zero gradients still decay the moments):
    point-attribute gradients are all-reduced DENSE, so each replica applies the identical update.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_patches(n_rays: int, rays_per_patch: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) ray range of `rank`; whole patches only (the blur module and
    drop_patch_rays work on whole 8x8 patches).  Earlier ranks get the remainder patches."""
    assert n_rays % rays_per_patch == 0, "ray count must be a whole number of patches"
    n_patch = n_rays // rays_per_patch
    base, rem = divmod(n_patch, world)
    p0 = rank * base + min(rank, rem)
    p1 = p0 + base + (1 if rank < rem else 0)
    return p0 * rays_per_patch, p1 * rays_per_patch


def shard_patch_rows(patch_num: int, patch_size: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) ray range of `rank` for a training frame whose rays are the raster of a (patch_num*patch_size)^2 pixel grid
    (data/scannet_ft_dataset.py:917-945): whole ROWS of patches, i.e. multiples of patch_size * S consecutive rays, so that
    every 8x8 patch stays on one rank (a contiguous 64-ray range of the raster is one pixel row of 8 different patches)."""
    S = patch_num * patch_size
    r0, r1 = shard_patches(patch_num * patch_size * S, patch_size * S, rank, world)
    return r0, r1


def shard_frame(frame: Dict[str, torch.Tensor], rank: int, world: int, rays_per_patch: int = 64) -> Dict[str, torch.Tensor]:
    """slice the per-ray entries of a frame dict; everything else is replicated.
    Frames that carry the dilated-patch layout (`dilation_PatchNum`, `dilation_PatchSize`) are cut along whole patch rows
    (`shard_patch_rows`); other frames (full-frame rendering) in contiguous multiples of `rays_per_patch` rays."""
    R = frame["raydir"].shape[1]
    PN, PS = frame.get("dilation_PatchNum"), frame.get("dilation_PatchSize")
    if PN is not None and PS is not None and (int(PN) * int(PS)) ** 2 == R:
        b, e = shard_patch_rows(int(PN), int(PS), rank, world)
    else:
        b, e = shard_patches(R, rays_per_patch, rank, world)
    out = dict(frame)
    for k in ("raydir", "gt_image"):
        if k in frame and frame[k] is not None and frame[k].dim() >= 2 and frame[k].shape[1] == R:
            out[k] = frame[k][:, b:e].contiguous()
    pix = frame.get("pixel_idx")
    if pix is not None and torch.is_tensor(pix):
        if pix.dim() == 3 and pix.shape[1] == R:                       # (1,R,2)
            out["pixel_idx"] = pix[:, b:e].contiguous()
        elif pix.dim() == 4 and pix.shape[1] * pix.shape[2] == R:      # (1,S_h,S_w,2) as the dataset / FrameProducer hand it over
            out["pixel_idx"] = pix.reshape(pix.shape[0], R, 2)[:, b:e].contiguous()
    return out


def _flatten(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors]) if len(tensors) else torch.empty(0)


def _unflatten_into(flat: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def allreduce_gradients(params: Iterable[torch.nn.Parameter], n_local_valid: torch.Tensor, group=None,
                        bucket_bytes: int = 64 << 20) -> torch.Tensor:
    """Scale every existing .grad by n_local/n_global (the global-mean normalisation) and SUM all-reduce
    them in buckets (small tensors coalesced; tensors >= bucket_bytes reduced in place).  Parameters
    whose grad is None on this rank (e.g. it saw no valid ray) take part with zeros so the collective
    sequence is identical on all ranks.  Returns n_global (tensor).  Asynchronous ops are all waited
    for before returning."""
    params = [p for p in params if p.requires_grad]
    n_local = n_local_valid.detach().to(torch.float32).reshape(1)
    n_global = n_local.clone()
    dist.all_reduce(n_global, op=dist.ReduceOp.SUM, group=group)
    scale = n_local / torch.clamp(n_global, min=1.0)
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        p.grad.mul_(scale.to(p.grad.dtype))
    small, handles = [], []
    for p in params:
        g = p.grad
        if g.numel() * g.element_size() >= bucket_bytes:
            handles.append((dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True), None, None))
        else:
            small.append(g)
    # coalesce the small gradients (MLP + conv weights: 449,381 floats = 1.8 MB) into as few buckets as fit
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if bucket:
            flat = _flatten(bucket)
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bucket))
            bucket, size = [], 0
    for g in small:
        nb = g.numel() * g.element_size()
        if size + nb > bucket_bytes:
            flush()
        bucket.append(g)
        size += nb
    flush()
    for h, flat, tensors in handles:
        h.wait()
        if flat is not None:
            _unflatten_into(flat, tensors)
    return n_global


def train_step(net, frame_shard: Dict[str, torch.Tensor], optimizers: Sequence[torch.optim.Optimizer], group=None,
               zero_one_weight: float = 1e-4, next_frame_shard: Optional[Dict[str, torch.Tensor]] = None):
    """one data-parallel training step on this rank's ray shard: forward (fused hot path), loss, backward,
    gradient all-reduce with global-mean normalisation, optimiser steps.  Returns (loss, n_global_valid).
    `next_frame_shard`: the shard of the NEXT step, if known -- its voxel query is enqueued before this step's backward pass
    so that the next forward does not stall at the query's read-back (NeuralPointsRayMarching.prefetch_query)."""
    from .renderer import training_loss
    for o in optimizers:
        o.zero_grad(set_to_none=True)
    out = net(**frame_shard)
    n_local = (out["ray_mask"] > 0).sum()
    if next_frame_shard is not None and getattr(net, "near_far", None) is not None:
        net.prefetch_query(**next_frame_shard)
    if out["coarse_raycolor"].shape[1] > 0:
        loss = training_loss(out, frame_shard["gt_image"], zero_one_weight)
        loss.backward()
    else:
        loss = torch.zeros((), device=out["ray_mask"].device)
    params = [p for o in optimizers for g in o.param_groups for p in g["params"]]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        n_global = allreduce_gradients(params, n_local, group)
    else:
        n_global = n_local
    for o in optimizers:
        o.step()
    return loss.detach(), n_global
