"""CPU oracle for the voxel-grid neural-point query (numpy; torch CPU only for the ray generator).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Restates models/neural_points/query_point_indices_worldcoords.py of the reference:
  grid_params        <- get_hyperparameters            (:46-77)
  cell_of            <- the voxel-coordinate expression (:259-261, :400-402, :465-467)
  build_grid         <- claim_occ / map_coor2occ / fill_occ2pnts (:237-381, :540-602)
  ray_candidates     <- near_far_linear_ray_generation  (models/rendering/diff_ray_marching.py:349-392)
  select_samples     <- mask_raypos + cumsum + get_shadingloc (:384-433, :623-667)
  layered_knn        <- query_neigh_along_ray_layered   (:436-522)
  query              <- query_grid_point_index + query_points tail (:605-711, :80-103)

The reference's grid build is racy by design (atomic arrival order, time-seeded reservoir
replacement); the deterministic canonical form used here and by the CUDA product:
  * stored points of a cell are in ascending point-index order, at most P kept (lowest indices);
    parity configs keep every cell <= P so this coincides with the reference;
  * the "slot 0" cell (the cell that loses its points because of `voxel_idx > 0`, :366) is the
    cell of the first in-grid point unless `skip_cell` overrides it (tests that run the real
    reference kernels read the winner from the reference's occ_2_coor[0] and pass it in);
  * neighbour sets are compared as sets; this oracle also returns the reference's visit-order
    slots (x-major shell walk, replace-farthest buffer) for information.

Parity pinning: misc.npz pins ray_candidates against the reference's torch generator on CPU; the
kernels themselves can only run on a GPU, so tests/test_gpu_query_vs_reference.py runs the
reference's own compiled kernels (oracle/_ref/ref_query_k8.cubin, built by
oracle/build_ref_query_cubin.py) on the B200 box against this file and against the product.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

f32 = np.float32


@dataclass
class GridParams:
    origin: np.ndarray      # (3,) f32   ranges[:3] after padding
    upper: np.ndarray       # (3,) f32
    cell: np.ndarray        # (3,) f32   vsize*vscale
    dims: np.ndarray        # (3,) i32
    radius2: np.float32
    vsize: np.ndarray       # (3,) as given


def grid_params(xyz: np.ndarray, vsize, vscale, kernel_size, ranges, radius_limit_scale) -> GridParams:
    """xyz (N,3) f32.  Mirrors the mixed f32/f64 arithmetic of get_hyperparameters (:46-77)."""
    xyz = np.asarray(xyz, f32)
    mn, mx = xyz.min(axis=0), xyz.max(axis=0)
    vsize_l = [float(v) for v in vsize]
    vscale_np = np.array(vscale, dtype=np.int32)
    scaled_vsize = (np.asarray(vsize_l) * vscale_np).astype(f32)
    if ranges is not None:
        r = np.asarray(ranges, dtype=f32)
        mn, mx = np.maximum(mn, r[:3]), np.minimum(mx, r[3:])
    pad = (scaled_vsize * np.asarray(list(kernel_size)) / 2).astype(f32)      # f32*int64 -> f64 -> f32
    mn, mx = (mn - pad).astype(f32), (mx + pad).astype(f32)
    vdim = (mx - mn).astype(f32) / np.asarray(vsize_l)                           # f64
    dims = np.ceil(vdim / vscale_np).astype(np.int32)
    rl = f32(radius_limit_scale * max(vsize_l[0], vsize_l[1]))
    return GridParams(mn, mx, scaled_vsize, dims, f32(rl * rl), np.asarray(vsize_l))


def cell_of(p: np.ndarray, gp: GridParams) -> np.ndarray:
    """(…,3) f32 -> (…,3) int64 cell coordinate: (int)floor((p - origin)/cell), IEEE f32."""
    q = (np.asarray(p, f32) - gp.origin).astype(f32) / gp.cell
    return np.floor(q.astype(f32)).astype(np.int64)


def in_grid(c: np.ndarray, gp: GridParams) -> np.ndarray:
    return np.all((c >= 0) & (c < gp.dims.astype(np.int64)), axis=-1)


def lin_index(c: np.ndarray, gp: GridParams) -> np.ndarray:
    d = gp.dims.astype(np.int64)
    return (c[..., 0] * d[1] + c[..., 1]) * d[2] + c[..., 2]


@dataclass
class Grid:
    gp: GridParams
    cell_points: Dict[int, np.ndarray]    # lin cell -> ascending point ids (<=P), skip cell absent
    occupied: np.ndarray                  # sorted unique lin ids of cells holding >=1 in-grid point
    dilated: np.ndarray                   # bool (X,Y,Z)
    skip_cell: int
    max_cell_count: int


def build_grid(xyz: np.ndarray, gp: GridParams, P: int, query_size, max_o: Optional[int] = None,
               skip_cell: Optional[int] = None) -> Grid:
    xyz = np.asarray(xyz, f32)
    c = cell_of(xyz, gp)
    ok = in_grid(c, gp)
    ids = np.nonzero(ok)[0]
    lin = lin_index(c[ok], gp)
    if skip_cell is None:
        skip_cell = int(lin[0]) if len(lin) else -1
    order = np.argsort(lin, kind="stable")          # stable -> ascending point id inside a cell
    ls, ps = lin[order], ids[order]
    uniq, start, counts = np.unique(ls, return_index=True, return_counts=True)
    if max_o is not None and len(uniq) > max_o:
        raise ValueError(f"{len(uniq)} occupied cells > max_o={max_o}: reference would take its random-replacement path")
    cell_points = {}
    for u, s, n in zip(uniq.tolist(), start.tolist(), counts.tolist()):
        if u == skip_cell:
            continue
        cell_points[u] = ps[s:s + min(n, P)]
    dims = gp.dims.astype(np.int64)
    dil = np.zeros(tuple(dims), dtype=bool)
    uc = np.stack(np.unravel_index(uniq, tuple(dims)), axis=-1)
    q = [int(v) for v in query_size]
    for dx in range(-(q[0] // 2), (q[0] + 1) // 2):
        for dy in range(-(q[1] // 2), (q[1] + 1) // 2):
            for dz in range(-(q[2] // 2), (q[2] + 1) // 2):
                n = uc + np.array([dx, dy, dz])
                m = np.all((n >= 0) & (n < dims), axis=-1)
                dil[n[m, 0], n[m, 1], n[m, 2]] = True
    return Grid(gp, cell_points, uniq, dil, skip_cell, int(counts.max()) if len(counts) else 0)


def candidate_ts(D: int, near: float, far: float, noise=None, jitter: float = 0.0, R: int = 1):
    """mid-point parameters of the D depth candidates, (1,R,D) f32, with torch CPU ops in the order
    of diff_ray_marching.py:369-385 (`noise` = the torch.rand((1,R,D)) draw when jittering)."""
    import torch
    t = torch.linspace(0, 1, D + 1).view(1, -1)
    t = near * (1 - t) + far * t
    if noise is None:
        noise = torch.zeros((1, R, D))
    else:
        noise = torch.as_tensor(noise)
    seg = (t[..., 1:] - t[..., :-1]) * (1 + jitter * (noise - 0.5))
    end = torch.cumsum(seg, dim=2)
    end = torch.cat([torch.zeros((end.shape[0], end.shape[1], 1)), end], dim=2)
    end = near + end
    return ((end[:, :, :-1] + end[:, :, 1:]) / 2).numpy()


def ray_candidates(campos, raydir, D, near, far, noise=None, jitter: float = 0.0):
    """-> raypos (1,R,D,3) f32, ts (1,R,D).  pos = campos + raydir*t with separate f32 mul, add."""
    campos, raydir = np.asarray(campos, f32), np.asarray(raydir, f32)
    ts = candidate_ts(D, near, far, noise, jitter, R=raydir.shape[1])
    pos = campos[:, None, None, :] + (raydir[:, :, None, :] * ts[..., None]).astype(f32)
    return pos.astype(f32), ts


def positions_from_ts(campos, raydir, ts):
    campos, raydir, ts = np.asarray(campos, f32), np.asarray(raydir, f32), np.asarray(ts, f32)
    if ts.ndim == 1:
        ts = ts[None, None, :]
    elif ts.ndim == 2:
        ts = ts[None]
    return (campos[:, None, None, :] + (raydir[:, :, None, :] * ts[..., None]).astype(f32)).astype(f32)


def select_samples(raypos: np.ndarray, grid: Grid, SR: int):
    """raypos (1,R,D,3) -> (hit_mask (R,D) bool, ray_mask (R,) bool, sample_loc (R',SR,3) f32,
    sample_mask (R',SR) i32) where R' = rays with >=1 hit; first SR hits in depth order."""
    rp = raypos[0]
    c = cell_of(rp, grid.gp)
    ok = in_grid(c, grid.gp)
    hit = np.zeros(rp.shape[:2], dtype=bool)
    cc = c[ok]
    hit[ok] = grid.dilated[cc[:, 0], cc[:, 1], cc[:, 2]]
    ray_mask = hit.any(axis=1)
    rows = np.nonzero(ray_mask)[0]
    loc = np.zeros((len(rows), SR, 3), f32)
    msk = np.zeros((len(rows), SR), np.int32)
    for i, r in enumerate(rows):
        idx = np.nonzero(hit[r])[0][:SR]
        loc[i, :len(idx)] = rp[r, idx]
        msk[i, :len(idx)] = 1
    return hit, ray_mask, loc, msk


def _d2(p: np.ndarray, c: np.ndarray) -> np.ndarray:
    """squared distance with the rounding of the compiled reference kernel (sm_100a SASS of
    :489-492: FMUL y*y, FFMA x*x+., FFMA z*z+.), emulated through float64."""
    v = (p.astype(f32) - c.astype(f32)).astype(f32).astype(np.float64)
    t = (v[..., 1] * v[..., 1]).astype(f32).astype(np.float64)
    t = (v[..., 0] * v[..., 0] + t).astype(f32).astype(np.float64)
    return (v[..., 2] * v[..., 2] + t).astype(f32)


def layered_knn(xyz: np.ndarray, grid: Grid, loc: np.ndarray, K: int, kernel_size) -> Tuple[np.ndarray, np.ndarray, int]:
    """one sample: returns (visit-order slots (K,) i32 with -1 padding -- the reference's output --,
    canonical ascending (d2, id) order (K,), number of stored points visited)."""
    gp = grid.gp
    f = cell_of(loc, gp)
    dims = gp.dims.astype(np.int64)
    slots = np.full(K, -1, np.int32)
    buf = np.zeros(K, f32)
    kid, far2, far_ind, visited = 0, f32(0), 0, 0
    seen = []
    for layer in range((int(kernel_size[0]) + 1) // 2):
        for x in range(max(-f[0], -layer), min(dims[0] - f[0], layer + 1)):
            for y in range(max(-f[1], -layer), min(dims[1] - f[1], layer + 1)):
                for z in range(max(-f[2], -layer), min(dims[2] - f[2], layer + 1)):
                    if max(abs(x), abs(y), abs(z)) != layer:
                        continue
                    pts = grid.cell_points.get(int(((f[0] + x) * dims[1] + f[1] + y) * dims[2] + f[2] + z))
                    if pts is None:
                        continue
                    visited += len(pts)
                    d2 = _d2(xyz[pts], loc)
                    for pid, dd in zip(pts.tolist(), d2.tolist()):
                        dd = f32(dd)
                        if gp.radius2 == 0 or dd <= gp.radius2:
                            seen.append((dd, pid))
                            if kid < K:
                                slots[kid], buf[kid] = pid, dd
                                if dd > far2:
                                    far2, far_ind = dd, kid
                                kid += 1
                            else:
                                kid += 1
                                if dd < far2:
                                    slots[far_ind], buf[far_ind] = pid, dd
                                    far2 = dd
                                    for i in range(K):
                                        if buf[i] > far2:
                                            far2, far_ind = buf[i], i
        if kid >= K:
            break
    canon = np.full(K, -1, np.int32)
    seen.sort()
    for i, (_, pid) in enumerate(seen[:K]):
        canon[i] = pid
    return slots, canon, visited


def query(xyz, campos, camrot, raydir, ts, *, vsize, vscale, kernel_size, query_size, ranges, radius_limit_scale,
          SR, K, P, max_o=None, skip_cell=None) -> Dict[str, np.ndarray]:
    """Full query for one camera.  xyz (N,3); campos (1,3); camrot (1,3,3) c2w rotation;
    raydir (1,R,3); ts (1,R,D) or (D,).  Returns the tensors lighting_fast_querier.query_points
    returns (:80-93) plus diagnostics."""
    xyz = np.asarray(xyz, f32)
    gp = grid_params(xyz, vsize, vscale, kernel_size, ranges, radius_limit_scale)
    grid = build_grid(xyz, gp, P, query_size, max_o, skip_cell)
    raypos = positions_from_ts(campos, raydir, ts)
    hit, ray_mask, loc, smask = select_samples(raypos, grid, SR)
    Rp = loc.shape[0]
    pidx_visit = np.full((Rp, SR, K), -1, np.int32)
    pidx = np.full((Rp, SR, K), -1, np.int32)
    cand = 0
    for r in range(Rp):
        for s in range(SR):
            if smask[r, s] > 0:
                pidx_visit[r, s], pidx[r, s], v = layered_knn(xyz, grid, loc[r, s], K, kernel_size)
                cand += v
    keep = (pidx >= 0).reshape(Rp, -1).any(axis=1)
    final_mask = np.zeros(ray_mask.shape, np.int8)
    final_mask[np.nonzero(ray_mask)[0][keep]] = 1
    loc_w = loc[keep]
    cp, rot = np.asarray(campos, f32)[0], np.asarray(camrot, f32)[0]
    sh = (loc_w - cp).astype(f32)
    xc = np.stack([(sh * rot[:, j]).astype(f32).sum(axis=-1, dtype=f32) for j in range(3)], axis=-1) if len(loc_w) else np.zeros((0, SR, 3), f32)
    pers = np.stack([xc[..., 0] / xc[..., 2], xc[..., 1] / xc[..., 2], xc[..., 2]], axis=-1).astype(f32) if len(loc_w) else xc
    dirs = np.broadcast_to(np.asarray(raydir, f32)[0][final_mask > 0][:, None, :], loc_w.shape).copy()
    return dict(sample_pidx=pidx[keep][None], sample_pidx_visit_order=pidx_visit[keep][None], sample_loc=pers[None],
                sample_loc_w=loc_w[None], sample_ray_dirs=dirs[None], ray_mask=final_mask[None],
                sample_mask=smask[keep][None], hit_mask=hit, n_candidates=cand, grid=grid, gp=gp)
