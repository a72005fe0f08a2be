"""Deterministic synthetic scenes, cameras and rays for the BASELINE configs.

There is no network and no dataset: every parity test and the bench use inputs made here.  All
generators are seeded numpy and return float32 numpy arrays (callers move them to torch/CUDA).
Ray conventions follow the reference data layer's *output contract* (SURVEY.md Appendix A.1):
``raydir = [(px+0.5-cx)/fx, (py+0.5-cy)/fy, 1] @ c2w[:3,:3].T`` un-normalised
(reference: data/data_utils.py:57-71, dir_norm=0).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

LEGO_BOX = np.array([-0.638, -1.141, -0.346, 0.634, 1.149, 1.141], dtype=np.float32)  # lego_hybrid.sh ranges


# ----------------------------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------------------------
def look_at(eye, center=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """camera-to-world 4x4, camera looks along +z, x right, y down (OpenCV style)."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(center, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = r, d, f, eye
    return c2w.astype(np.float32)


def intrinsic_matrix(fx, fy, cx, cy) -> np.ndarray:
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)


def rays_for_pixels(px: np.ndarray, py: np.ndarray, K: np.ndarray, c2w: np.ndarray) -> np.ndarray:
    x = (px.astype(np.float32) + np.float32(0.5) - K[0, 2]) / K[0, 0]
    y = (py.astype(np.float32) + np.float32(0.5) - K[1, 2]) / K[1, 1]
    d = np.stack([x, y, np.ones_like(x)], axis=-1).astype(np.float32)
    return (d @ c2w[:3, :3].T.astype(np.float32)).astype(np.float32)


def full_frame_pixels(H: int, W: int) -> Tuple[np.ndarray, np.ndarray]:
    py, px = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    return px.reshape(-1), py.reshape(-1)


def dilated_patch_pixels(rng: np.random.Generator, H: int, W: int, patch_num: int, patch_size: int,
                         dil_lo: int = 1, dil_hi: int = 8, margin: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """patch_num^2 patches of patch_size^2 pixels with a random dilation each, laid out on an
    S x S raster (S = patch_num*patch_size), row-major -- the layout blur/drop_patch rely on
    (mimics data/scannet_ft_dataset.py:917-940)."""
    S = patch_num * patch_size
    px = np.zeros((S, S), np.float32)
    py = np.zeros((S, S), np.float32)
    gy, gx = np.meshgrid(np.arange(patch_size, dtype=np.float32), np.arange(patch_size, dtype=np.float32), indexing="ij")
    for i in range(patch_num):
        for j in range(patch_num):
            dil = int(rng.integers(dil_lo, dil_hi + 1))
            x0 = int(rng.integers(margin, W - margin - (patch_size - 1) * dil))
            y0 = int(rng.integers(margin, H - margin - (patch_size - 1) * dil))
            px[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size] = x0 + dil * gx
            py[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size] = y0 + dil * gy
    return px.reshape(-1), py.reshape(-1)


# ----------------------------------------------------------------------------------------------
# point clouds
# ----------------------------------------------------------------------------------------------
def point_attributes(rng: np.random.Generator, N: int, F: int = 32) -> Dict[str, np.ndarray]:
    d = rng.standard_normal((N, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True) + 1e-12
    return dict(
        emb=(rng.random((N, F), dtype=np.float32) - np.float32(0.5)),       # feature_init_method=rand
        conf=rng.random((N, 1), dtype=np.float32),
        color=rng.random((N, 3), dtype=np.float32),
        dir=d.astype(np.float32),
    )


def _grid_origin(xyz: np.ndarray, cell: float, ranges, kernel: int = 3) -> np.ndarray:
    """origin of the query grid for this cloud: bbox (clipped to `ranges`) minus cell*kernel/2, in
    the f32 arithmetic of the reference's get_hyperparameters (query_point_indices_worldcoords.py:46-77)."""
    mn = xyz.min(axis=0).astype(np.float32)
    if ranges is not None:
        mn = np.maximum(mn, np.asarray(ranges[:3], np.float32))
    pad = (np.full(3, cell, np.float32) * np.asarray([kernel] * 3) / 2).astype(np.float32)
    return (mn - pad).astype(np.float32)


def _thin_cells(xyz: np.ndarray, cell: float, max_per_cell: int, ranges=None) -> np.ndarray:
    """drop points so that no voxel of the query grid holds more than max_per_cell (keeps the parity
    configs off the reference's random-replacement path).  The bbox-defining points
    are kept so the grid origin does not move."""
    for _ in range(4):
        origin = _grid_origin(xyz, cell, ranges)
        c = np.floor(((xyz - origin).astype(np.float32) / np.float32(cell)).astype(np.float32)).astype(np.int64)
        c -= c.min(axis=0)
        dims = c.max(axis=0) + 1
        key = (c[:, 0] * dims[1] + c[:, 1]) * dims[2] + c[:, 2]
        prio = np.ones(len(xyz), np.int8)
        prio[np.r_[xyz.argmin(axis=0), xyz.argmax(axis=0)]] = 0        # extreme points first in their voxel
        order = np.lexsort((np.arange(len(xyz)), prio, key))
        ks = key[order]
        start = np.r_[0, np.nonzero(np.diff(ks))[0] + 1]
        rank = np.arange(len(ks)) - np.repeat(start, np.diff(np.r_[start, len(ks)]))
        keep = np.zeros(len(ks), bool)
        keep[order] = rank < max_per_cell
        if keep.all():
            break
        xyz = xyz[keep]
    # bbox-defining points to the front so that truncating the cloud keeps the grid origin
    ext = np.unique(np.r_[xyz.argmin(axis=0), xyz.argmax(axis=0)])
    rest = np.setdiff1d(np.arange(len(xyz)), ext, assume_unique=False)
    return xyz[np.r_[ext, rest]]


def lego_scene(N: int, seed: int = 0, vsize: float = 0.004, vscale: int = 2, P: int = 12) -> np.ndarray:
    """surface-like cloud: union of 32 random ellipsoid shells inside the lego box (config 2)."""
    rng = np.random.default_rng(seed)
    lo, hi = LEGO_BOX[:3], LEGO_BOX[3:]
    n_shell = 32
    centres = lo + (hi - lo) * (0.2 + 0.6 * rng.random((n_shell, 3)))
    radii = rng.uniform(0.05, 0.35, size=(n_shell, 3))
    out = []
    need = int(N * 1.15) + 1024
    per = need // n_shell + 1
    for s in range(n_shell):
        u = rng.standard_normal((per, 3))
        u /= np.linalg.norm(u, axis=-1, keepdims=True)
        p = centres[s] + u * radii[s] + rng.uniform(-0.5, 0.5, size=(per, 1)) * vsize * u
        out.append(p)
    xyz = np.concatenate(out).astype(np.float32)
    inside = np.all((xyz > lo) & (xyz < hi), axis=-1)
    xyz = xyz[inside]
    xyz = xyz[rng.permutation(len(xyz))]
    xyz = _thin_cells(xyz, vsize * vscale, P - 1, LEGO_BOX)
    if len(xyz) < N:
        raise RuntimeError(f"lego_scene produced {len(xyz)} < {N} points; lower N")
    return np.ascontiguousarray(xyz[:N])


def room_scene(N: int, seed: int = 0, size=(6.0, 5.0, 3.0), vsize: float = 0.008, vscale: int = 2, P: int = 26,
               n_boxes: int = 20) -> np.ndarray:
    """inner surfaces of a room + random 1 m cuboids, +-4 mm surface noise (configs 3-5)."""
    rng = np.random.default_rng(seed)
    size = np.asarray(size, np.float64)
    boxes = [(np.zeros(3), size)]
    for _ in range(n_boxes):
        c = rng.uniform(0.6, 1.0, 3) * 0 + rng.uniform([0.6, 0.6, 0.0], size - [0.6, 0.6, 1.0])
        boxes.append((c - [0.5, 0.5, 0.0], c + [0.5, 0.5, 1.0]))
    areas = []
    for lo, hi in boxes:
        e = hi - lo
        areas.append(2 * (e[0] * e[1] + e[1] * e[2] + e[0] * e[2]))
    areas = np.asarray(areas)
    need = int(N * 1.2) + 1024
    counts = np.maximum((need * areas / areas.sum()).astype(int), 1)
    out = []
    for (lo, hi), n in zip(boxes, counts):
        e = hi - lo
        fa = np.array([e[1] * e[2], e[1] * e[2], e[0] * e[2], e[0] * e[2], e[0] * e[1], e[0] * e[1]])
        face = rng.choice(6, size=n, p=fa / fa.sum())
        p = lo + rng.random((n, 3)) * e
        ax = face // 2
        side = face % 2
        p[np.arange(n), ax] = np.where(side == 0, lo[ax], hi[ax]) + rng.uniform(-0.004, 0.004, n)
        out.append(p)
    xyz = np.concatenate(out).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    xyz = _thin_cells(xyz, vsize * vscale, P - 1, None)
    if len(xyz) < N:
        raise RuntimeError(f"room_scene produced {len(xyz)} < {N} points; lower N")
    return np.ascontiguousarray(xyz[:N])


# ----------------------------------------------------------------------------------------------
# frame dicts (Appendix A.1 layout, numpy)
# ----------------------------------------------------------------------------------------------
def lego_frame(H: int = 800, W: int = 800, V: int = 4, seed: int = 0, azimuth: float = 0.6,
               elevation: float = 0.5, radius: float = 4.0) -> Dict[str, np.ndarray]:
    """config 2: camera on a radius-4 sphere looking at the scene centre; V neighbouring views."""
    rng = np.random.default_rng(seed + 17)
    focal = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
    K = intrinsic_matrix(focal, focal, W / 2.0, H / 2.0)
    centre = 0.5 * (LEGO_BOX[:3] + LEGO_BOX[3:])

    def cam(az):
        eye = centre + radius * np.array([math.cos(elevation) * math.cos(az), math.cos(elevation) * math.sin(az), math.sin(elevation)])
        return look_at(eye, centre)

    c2w = cam(azimuth)
    px, py = full_frame_pixels(H, W)
    offs = [(-1) ** i * (i // 2 + 1) * 0.08 for i in range(V)]
    c2w_n = np.stack([cam(azimuth + o) for o in offs]).astype(np.float32)
    return dict(
        campos=c2w[None, :3, 3].copy(), camrotc2w=c2w[None, :3, :3].copy(), c2w=c2w[None],
        raydir=rays_for_pixels(px, py, K, c2w)[None], pixel_idx=np.stack([px, py], -1)[None],
        near=np.full((1, 1, 1), 2.0, np.float32), far=np.full((1, 1, 1), 6.0, np.float32),
        h=np.array([H]), w=np.array([W]), intrinsic=K[None], bg_color=np.ones((1, 3), np.float32),
        images_nearest=rng.random((1, V, H, W, 3), dtype=np.float32),
        c2w_nearest=c2w_n[None], campos_nearest=c2w_n[None, :, :3, 3].copy(), intrinsic_nearest=K[None],
    )


def room_frame(H: int = 480, W: int = 640, V: int = 8, patch_num: int = 8, patch_size: int = 8, seed: int = 0,
               size=(6.0, 5.0, 3.0)) -> Dict[str, np.ndarray]:
    """config 3/4: ScanNet-like intrinsics, camera inside the room, dilated patch rays."""
    rng = np.random.default_rng(seed + 29)
    s = W / 640.0
    K = intrinsic_matrix(577.87 * s, 577.87 * s, W / 2.0, H / 2.0)
    size = np.asarray(size)

    def cam(t):
        eye = np.array([0.35 * size[0] + 0.05 * t, 0.4 * size[1] + 0.03 * t, 1.4])
        tgt = np.array([0.9 * size[0], 0.55 * size[1] + 0.2 * t, 1.0])
        return look_at(eye, tgt)

    c2w = cam(0.0)
    px, py = dilated_patch_pixels(rng, H, W, patch_num, patch_size)
    offs = [(-1) ** i * (i // 2 + 1) for i in range(V)]
    c2w_n = np.stack([cam(float(o)) for o in offs]).astype(np.float32)
    R = px.shape[0]
    return dict(
        campos=c2w[None, :3, 3].copy(), camrotc2w=c2w[None, :3, :3].copy(), c2w=c2w[None],
        raydir=rays_for_pixels(px, py, K, c2w)[None], pixel_idx=np.stack([px, py], -1)[None],
        near=np.full((1, 1, 1), 0.1, np.float32), far=np.full((1, 1, 1), 8.0, np.float32),
        h=np.array([H]), w=np.array([W]), intrinsic=K[None], bg_color=np.ones((1, 3), np.float32),
        gt_image=rng.random((1, R, 3), dtype=np.float32),
        images_nearest=rng.random((1, V, H, W, 3), dtype=np.float32),
        c2w_nearest=c2w_n[None], campos_nearest=c2w_n[None, :, :3, 3].copy(), intrinsic_nearest=K[None],
        dilation_PatchNum=np.array([patch_num]), dilation_PatchSize=np.array([patch_size]),
    )


def frame_scene(seed: int = 31, F: int = 14, H: int = 40, W: int = 56, step: int = 5):
    """inputs of the frame-producer golden: random uint8 frames, room-like poses, frame numbers 0,5,10,...; every 4th frame is a
    test frame (not in train_id_list).  Deterministic, regenerated by the tests (only reference OUTPUTS are stored)."""
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    vids = [step * i for i in range(F)]
    c2w = np.stack([look_at(np.array([2.0 + 0.05 * t, 2.0 + 0.03 * t, 1.4]), np.array([5.4, 2.75 + 0.2 * t, 1.0])) for t in range(F)]).astype(np.float32)
    s = W / 640.0
    K = intrinsic_matrix(577.87 * s, 577.87 * s, W / 2.0 - 0.37, H / 2.0 + 0.21).astype(np.float32)
    train_ids = [v for i, v in enumerate(vids) if i % 4 != 3]
    test_ids = [v for i, v in enumerate(vids) if i % 4 == 3]
    return images, c2w, vids, K, train_ids, test_ids


# ----------------------------------------------------------------------------------------------
# config 1: gathered inputs with precomputed (random) neighbours
# ----------------------------------------------------------------------------------------------
def render_stage_inputs(seed: int = 0, N: int = 200_000, R: int = 1024, SR: int = 80, K: int = 8, V: int = 4,
                        H: int = 120, W: int = 160, empty_frac: float = 0.5, F: int = 32) -> Dict[str, np.ndarray]:
    """config 1 recipe: uniform points in the lego box, random neighbour ids,
    `empty_frac` of the samples fully masked, V random reference images, random projections."""
    rng = np.random.default_rng(seed)
    lo, hi = LEGO_BOX[:3], LEGO_BOX[3:]
    xyz = (lo + (hi - lo) * rng.random((N, 3))).astype(np.float32)
    att = point_attributes(rng, N, F)
    pidx = rng.integers(-1, N, size=(1, R, SR, K)).astype(np.int32)
    empty = rng.random((1, R, SR)) < empty_frac
    pidx[empty] = -1
    loc_w = (lo + (hi - lo) * rng.random((1, R, SR, 3))).astype(np.float32)
    campos = np.array([[0.1, -0.2, -3.0]], np.float32)
    camrot = np.eye(3, dtype=np.float32)[None]

    def pers(p):
        c = p - campos[0]
        return np.stack([c[..., 0] / c[..., 2], c[..., 1] / c[..., 2], c[..., 2]], -1).astype(np.float32)

    raydir = rng.standard_normal((1, R, 1, 3)).astype(np.float32) * 0.3 + np.array([0, 0, 1], np.float32)
    out = dict(
        xyz=xyz, xyz_pers=pers(xyz), **att, sample_pidx=pidx, sample_loc_w=loc_w, sample_loc=pers(loc_w),
        sample_ray_dirs=np.ascontiguousarray(np.broadcast_to(raydir, (1, R, SR, 3))).astype(np.float32),
        campos=campos, camrotc2w=camrot,
        images_nearest=rng.random((1, V, H, W, 3), dtype=np.float32),
        sample_loc_i_n=np.stack([rng.uniform(-5, 1.2 * W, size=(V, R, SR)), rng.uniform(-5, 1.2 * H, size=(V, R, SR))], -1).astype(np.float32),
        delta_viewdir_n=(rng.standard_normal((V, R, SR, 3)) * 0.1).astype(np.float32),
        vsize=np.array([0.004, 0.004, 0.004], np.float32),
    )
    return out


def gather_neighbours(d: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """materialise the (1,R,SR,K,C) tensors NeuralPoints.forward returns (reference:
    models/neural_points/neural_points.py:708-733): masked slots alias point 0."""
    idx = np.maximum(d["sample_pidx"], 0)
    return dict(
        sampled_embedding=d["emb"][idx], sampled_xyz=d["xyz"][idx], sampled_xyz_pers=d["xyz_pers"][idx],
        sampled_color=d["color"][idx], sampled_dir=d["dir"][idx], sampled_conf=d["conf"][idx],
        sample_pnt_mask=d["sample_pidx"] >= 0,
    )


# ----------------------------------------------------------------------------------------------
# seeded aggregator weights (shipped layer shapes)
# ----------------------------------------------------------------------------------------------
LAYER_SHAPES = {
    "block1.0": (256, 284), "block1.2": (256, 256),
    "block3.0": (256, 263), "block3.2": (256, 256),
    "alpha_branch.0": (1, 256),
    "color_feature_branch.0": (128, 280), "color_feature_branch.2": (128, 128), "color_feature_branch.4": (128, 128),
    "aux_merge_weight_block.0": (64, 176), "aux_merge_weight_block.2": (64, 64),
    "aux_merge_weight_block.4": (64, 64), "aux_merge_weight_block.6": (1, 64),
    "color_mixup_block.0": (45, 90), "color_mixup_block.2": (45, 45), "color_mixup_block.4": (45, 45),
    "color_final_block.0": (3, 128),
}
CONV_SHAPES = {
    "aux_block_s1.0": (6, 3, 3, 3), "aux_block_s1.2": (6, 6, 3, 3),
    "aux_block_s2.0": (12, 6, 3, 3), "aux_block_s2.2": (12, 12, 3, 3),
    "aux_block_s3.0": (24, 12, 3, 3), "aux_block_s3.2": (24, 24, 3, 3),
}


def random_aggregator_params(seed: int = 0, dtype=torch.float32, bias_scale: float = 0.05) -> Dict[str, torch.Tensor]:
    """Xavier-like random weights with the shipped layer shapes, drawn from a
    numpy PCG64 stream so the values are stable across torch versions.  Biases are non-zero on
    purpose so parity tests exercise them."""
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp in {**LAYER_SHAPES, **CONV_SHAPES}.items():
        fan_out = shp[0] * (int(np.prod(shp[2:])) if len(shp) > 2 else 1)
        fan_in = int(np.prod(shp[1:]))
        bound = math.sqrt(2.0) * math.sqrt(6.0 / (fan_in + fan_out))
        P[name + ".weight"] = torch.from_numpy(((rng.random(shp) * 2 - 1) * bound).astype(np.float32)).to(dtype)
        P[name + ".bias"] = torch.from_numpy(((rng.random(shp[0]) * 2 - 1) * bias_scale).astype(np.float32)).to(dtype)
    return P
