"""Fused dense Adam for the neural-point tables.

The reference keeps one ``torch.optim.Adam`` over the point parameters (models/mvs_points_volumetric_model.py:94-104) and
updates EVERY row every step (dense semantics: a row with zero gradient still decays its moments and moves).  ``FusedAdam``
keeps those semantics and the same arithmetic but makes one pass over memory per parameter (csrc/adam.cu) instead of
torch's several for-each passes.  Parameters whose ``.grad`` is None are skipped, exactly like torch.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import check, lib, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with ops._launch(name="adam_step"):
                    check(lib().hnr_adam_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]),
                                              float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(st["step"]),
                                              stream()), "adam_step")
                # the kernel writes through the raw pointer: tell autograd / the version-keyed caches (voxel grid, packed weights)
                torch.autograd.graph.increment_version(p)
        return loss
