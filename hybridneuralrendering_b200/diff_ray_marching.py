"""Ray marching -- host-side mirror of models/rendering/diff_ray_marching.py and
models/rendering/diff_render_func.py (only what the hot path uses).

``ray_march`` keeps the reference signature and 7-tuple (:508-557).  ``render_func`` / ``blend_func``
are looked up by name in the reference (``find_render_function`` / ``find_blend_function``); the CUDA
kernel implements ``radiance_render`` + ``alpha_blend`` (the only pair any shipped config selects) and
anything else raises -- there is no fallback path.
"""
from __future__ import annotations

import torch

from . import ops


def alpha_blend(opacity, acc_transmission):
    return opacity * acc_transmission


def radiance_render(ray_feature):
    return ray_feature[..., 1:]


def no_tone_map(color, gamma=2.2, exposure=1):
    return color


def find_render_function(name):
    if name == 'radiance':
        return radiance_render
    raise RuntimeError('Unknown / unsupported render function: ' + name)


def find_blend_function(name):
    if name == 'alpha':
        return alpha_blend
    raise RuntimeError('Unknown / unsupported blend function: ' + name)


def find_tone_map(name):
    if name == 'off':
        return no_tone_map
    raise RuntimeError('Unknown / unsupported tone map: ' + name)


def near_far_linear_ray_generation(campos, raydir, point_count, near=0.1, far=10, jitter=0., **kargs):
    """(:349-392) kept for API parity; the fused query never materialises `raypos`."""
    tvals = torch.linspace(0, 1, point_count + 1, device=campos.device).view(1, -1)
    tvals = near * (1 - tvals) + far * tvals
    segment_length = (tvals[..., 1:] - tvals[..., :-1]) * (1 + jitter * (torch.rand((raydir.shape[0], raydir.shape[1], point_count), device=campos.device) - 0.5))
    end_point_ts = torch.cumsum(segment_length, dim=2)
    end_point_ts = torch.cat([torch.zeros((end_point_ts.shape[0], end_point_ts.shape[1], 1), device=end_point_ts.device), end_point_ts], dim=2)
    end_point_ts = near + end_point_ts
    middle_point_ts = (end_point_ts[:, :, :-1] + end_point_ts[:, :, 1:]) / 2
    raypos = campos[:, None, None, :] + raydir[:, :, None, :] * middle_point_ts[:, :, :, None]
    valid = torch.ones_like(middle_point_ts)
    segment_length = segment_length * torch.linalg.norm(raydir[..., None, :], axis=-1)
    return raypos, segment_length, valid, middle_point_ts


def _check_funcs(render_func, blend_func):
    if render_func is not radiance_render and getattr(render_func, "__name__", "") != "radiance_render":
        raise NotImplementedError("ray_march: only radiance_render is implemented in CUDA")
    if blend_func is not alpha_blend and getattr(blend_func, "__name__", "") != "alpha_blend":
        raise NotImplementedError("ray_march: only alpha_blend is implemented in CUDA")


def ray_march(ray_dist, ray_valid, ray_features, render_func, blend_func, bg_color=None):
    """ray_dist, ray_valid (N,R,SR); ray_features (N,R,SR,4) -> (ray_color (N,R,3), point_color,
    opacity, acc_transmission, blend_weight (N,R,SR,1), background_transmission (N,R,1),
    background_blend_weight)."""
    _check_funcs(render_func, blend_func)
    N, R, SR = ray_valid.shape
    assert ray_features.shape[-1] == 4, "radiance_render with 3 colour channels"
    valid = ray_valid.reshape(N * R, SR).to(torch.uint8).contiguous()
    dist = ray_dist.reshape(N * R, SR).float().contiguous()
    bg = None
    if bg_color is not None:
        bg = bg_color.reshape(-1, 3).float()
        if bg.shape[0] != 1:
            raise NotImplementedError("one background colour per call")
    color, opacity, accT, bw, bgT, _ = ops.CompositeFn.apply(ray_features.reshape(N * R, SR, 4), valid, None, 0, dist, bg, 0.0, 0)
    bgT = bgT.view(N, R, 1)
    return (color.view(N, R, 3), ray_features[..., 1:], opacity.view(N, R, SR), accT.view(N, R, SR), bw.view(N, R, SR, 1), bgT, bgT)


def ray_march_from_depth(sample_loc, ray_valid, ray_features, vsize_z, unit_mode, bg_color=None):
    """fused C1+C2: segment lengths are derived in-kernel from the samples' camera depth
    (neural_points_volumetric_model.py:331-339) -- used by NeuralPointsRayMarching."""
    N, R, SR = ray_valid.shape
    valid = ray_valid.reshape(N * R, SR).to(torch.uint8).contiguous()
    loc = sample_loc.reshape(N * R, SR, 3).float().contiguous()
    bg = bg_color.reshape(-1, 3).float() if bg_color is not None else None
    with ops.tag("composite"):
        color, opacity, accT, bw, bgT, dist = ops.CompositeFn.apply(ray_features.reshape(N * R, SR, 4), valid, loc.view(-1)[2:], 3, None, bg,
                                                                    float(vsize_z), int(unit_mode))
    return color.view(N, R, 3), opacity.view(N, R, SR), accT.view(N, R, SR), bw.view(N, R, SR, 1), bgT.view(N, R, 1), dist.view(N, R, SR)
