"""Fused dense Adam for the neural-point tables.

The reference keeps one ``torch.optim.Adam`` over the point parameters (models/mvs_points_volumetric_model.py:94-104) and
updates EVERY row every step (dense semantics: a row with zero gradient still decays its moments and moves).  ``FusedAdam``
keeps those semantics and the same arithmetic but makes one pass over memory per parameter (csrc/adam.cu) instead of
torch's several for-each passes.  Parameters whose ``.grad`` is None are skipped, exactly like torch.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import check, i64_array, lib, ptr, ptr_array, stream


class FusedAdam(torch.optim.Optimizer):
    """`guard_status=True`: the update is skipped ON THE DEVICE when the range guard of the split-fp16 kernels (ops.status_word) fired
    during this step -- the host raises at its next synchronisation point and no parameter has absorbed a saturated gradient."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, guard_status: bool = True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.guard_status = guard_status

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
            guard = ops.status_word(ps[0].device) if self.guard_status else None
            # tensors with the same step count (normally all of them) share one launch, 64 tensors at most per launch
            by_step = {}
            for p in ps:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            for step, plist in by_step.items():
                for c0 in range(0, len(plist), 64):
                    chunk = plist[c0:c0 + 64]
                    gs = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                    with ops._launch(name="adam_step"):
                        check(lib().hnr_adam_multi(len(chunk), ptr_array(chunk), ptr_array(gs), ptr_array([self.state[p]["exp_avg"] for p in chunk]),
                                                   ptr_array([self.state[p]["exp_avg_sq"] for p in chunk]), i64_array([p.numel() for p in chunk]),
                                                   float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                   int(step), ptr(guard), stream()), "adam_multi")
            for p in ps:
                # the kernel writes through the raw pointer: tell autograd / the version-keyed caches (voxel grid, packed weights)
                torch.autograd.graph.increment_version(p)
        return loss
