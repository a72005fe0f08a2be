"""Drop-in replacement of the reference's world-coordinate voxel querier.

Mirrors ``lighting_fast_querier`` of models/neural_points/query_point_indices_worldcoords.py:
same constructor ``(device, opt)``, same ``query_points`` positional arguments and the same
7-tuple return (``:80-93``), same ``clean_up``.  Differences that are NOT observable through that
API: the occupancy grid is built once per point-set change and cached (the reference rebuilds it
every call, ``:616``), candidate positions are never materialised, and there is exactly one
device->host readback per call (the two output counts).

Determinism: the reference's tables depend on atomic arrival order and a time-seeded RNG
(SURVEY.md §0.2).  Here neighbours come back sorted by (distance, point id); as a *set* they equal
the reference's whenever no voxel holds more than ``opt.P`` points and the occupied voxels fit in
``opt.max_o`` (checked, with a warning, at grid-build time).  The voxel that the reference leaves
empty because it won occupied-slot 0 (``:366``) is emulated by ``skip_cell`` (default: the voxel of
the first in-grid point; override with ``querier.skip_cell_override``).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import GridT, check, lib, ptr, stream


@dataclass
class QueryExtras:
    """by-products of the last query that the fused aggregation path consumes"""
    vlist: torch.Tensor        # (Nv,) int32: valid samples as indices into the (R''*SR) sample axis
    ray_ids: torch.Tensor      # (R'',) int32: original ray index of every kept ray
    n_rays: int
    n_valid: int


class _Grid:
    def __init__(self):
        self.key = None
        self.g: Optional[GridT] = None
        self.cell_start = self.pts_sorted = self.occ_bits = None
        self.info = None
        self.ranges_np = self.vsize_np = None


def _as_list(v):
    return [float(x) for x in (v.tolist() if hasattr(v, "tolist") else v)]


class lighting_fast_querier:
    def __init__(self, device, opt):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("lighting_fast_querier needs a CUDA device (no CPU fallback)")
        self.gpu = self.device.index
        self.opt = opt
        self.inverse = getattr(opt, "inverse", 0)
        if getattr(opt, "NN", 2) <= 0:
            # the reference asks for a kernel that does not exist when NN == 0 (:530)
            raise ValueError("opt.NN must be > 0 (layered nearest-neighbour query)")
        self.count = 0
        self._grid = _Grid()
        self._bufs = {}
        self.skip_cell_override: Optional[int] = None
        self.last: Optional[QueryExtras] = None
        lib()  # fail now if the library is missing

    # ------------------------------------------------------------------ lifecycle
    def clean_up(self):
        self._grid = _Grid()
        self._bufs = {}
        self.last = None

    def invalidate(self):
        """call after the point set changed in place (prune/grow replace the Parameter, which is
        detected automatically through data_ptr/_version)."""
        self._grid = _Grid()

    # ------------------------------------------------------------------ grid
    def get_hyperparameters(self, vsize_np, point_xyz_w_tensor, ranges=None):
        """(:46-77) grid origin/extent/cell/dims and radius limit; same mixed f32/f64 arithmetic."""
        opt = self.opt
        xyz = point_xyz_w_tensor
        mn, mx = torch.min(xyz, dim=-2)[0][0], torch.max(xyz, dim=-2)[0][0]
        vsize_l = _as_list(vsize_np)
        vscale_np = np.array(opt.vscale, dtype=np.int32)
        scaled_vsize_np = (np.asarray(vsize_l) * vscale_np).astype(np.float32)
        if ranges is not None:
            r = torch.as_tensor(_as_list(ranges), dtype=torch.float32, device=xyz.device)
            mn, mx = torch.maximum(mn, r[:3]), torch.minimum(mx, r[3:])
        pad = torch.as_tensor(scaled_vsize_np * np.asarray(list(opt.kernel_size)) / 2, device=xyz.device, dtype=torch.float32)
        mn, mx = mn - pad, mx + pad
        ranges_np = torch.cat([mn, mx], dim=-1).cpu().numpy().astype(np.float32)          # the one sync of a grid build
        vdim_np = (ranges_np[3:] - ranges_np[:3]).astype(np.float32) / np.asarray(vsize_l)
        scaled_vdim_np = np.ceil(vdim_np / vscale_np).astype(np.int32)
        radius_limit = np.float32(opt.radius_limit_scale * max(vsize_l[0], vsize_l[1]))
        depth_limit = np.float32(getattr(opt, "depth_limit_scale", 0) * vsize_l[2])
        return radius_limit, depth_limit, ranges_np, np.asarray(vsize_l), vdim_np, scaled_vsize_np, scaled_vdim_np, vscale_np

    def _ensure_grid(self, xyz: torch.Tensor) -> _Grid:
        opt = self.opt
        key = (xyz.data_ptr(), tuple(xyz.shape), xyz._version, tuple(opt.vsize), tuple(opt.vscale), tuple(opt.kernel_size),
               tuple(opt.query_size), tuple(opt.ranges) if opt.ranges is not None else None, int(opt.P), float(opt.radius_limit_scale),
               self.skip_cell_override)
        G = self._grid
        if G.key == key:
            return G
        xyz2 = xyz.reshape(-1, 3)
        if xyz2.dtype != torch.float32 or not xyz2.is_contiguous():
            xyz2 = xyz2.float().contiguous()
        N = xyz2.shape[0]
        radius_limit, _, ranges_np, vsize_np, _, scaled_vsize_np, scaled_vdim_np, _ = self.get_hyperparameters(opt.vsize, xyz.reshape(1, -1, 3), ranges=opt.ranges)
        g = GridT()
        for i in range(3):
            g.origin[i] = float(ranges_np[i])
            g.cell[i] = float(scaled_vsize_np[i])
            g.dims[i] = int(scaled_vdim_np[i])
            q = int(opt.query_size[i])
            g.qhalf_lo[i] = q // 2
            g.qhalf_hi[i] = (q + 1) // 2 - 1
        g.radius2 = float(np.float32(radius_limit) * np.float32(radius_limit))
        g.n_cells = int(scaled_vdim_np[0]) * int(scaled_vdim_np[1]) * int(scaled_vdim_np[2])
        g.P = int(opt.P)
        g.layers = (int(opt.kernel_size[0]) + 1) // 2
        if g.n_cells <= 0 or g.n_cells >= 2 ** 31 - 64:
            raise ValueError(f"voxel grid {list(scaled_vdim_np)} out of range; check vsize/ranges")
        dev = xyz2.device
        i32 = dict(device=dev, dtype=torch.int32)
        cell_of_pt = torch.empty(N, **i32)
        counts = torch.empty(g.n_cells, **i32)
        cell_start = torch.empty(g.n_cells + 1, **i32)
        tmp_idx = torch.empty(max(N, 1), **i32)
        pts_sorted = torch.empty((max(N, 1), 4), device=dev, dtype=torch.float32)
        occ_bits = torch.empty((g.n_cells + 31) // 32, **i32)
        scratch = torch.empty(int(lib().hnr_scan_scratch_elems(g.n_cells)), **i32)
        info = torch.empty(8, **i32)
        skip = -2 if self.skip_cell_override is None else int(self.skip_cell_override)
        check(lib().hnr_grid_build(ptr(xyz2), N, g, skip, ptr(cell_of_pt), ptr(counts), ptr(cell_start), ptr(tmp_idx), ptr(pts_sorted),
                                   ptr(occ_bits), ptr(scratch), ptr(info), stream()), "grid_build")
        ops._count(8)
        info_h = info.cpu().tolist()
        if info_h[1] > g.P:
            warnings.warn(f"a voxel holds {info_h[1]} points > P={g.P}: the reference would pick a random subset; "
                          "this implementation keeps the P lowest point ids")
        max_o = getattr(opt, "max_o", None)
        if max_o is not None and info_h[0] > max_o:
            warnings.warn(f"{info_h[0]} occupied voxels > max_o={max_o}: the reference would drop voxels at random; "
                          "this implementation keeps all of them")
        G = _Grid()
        G.key, G.g, G.cell_start, G.pts_sorted, G.occ_bits, G.info = key, g, cell_start, pts_sorted, occ_bits, info_h
        G.ranges_np, G.vsize_np = ranges_np, vsize_np
        self._grid = G
        return G

    # ------------------------------------------------------------------ candidates
    def candidate_ts(self, R: int, near: float, far: float, device) -> torch.Tensor:
        """mid-point parameters of the D depth candidates, produced with the SAME torch ops, order
        and RNG draw as near_far_linear_ray_generation (models/rendering/diff_ray_marching.py:349-392;
        jitter 0.3 when opt.is_train, :84-87) so the positions are the same floats.  Without jitter
        the result is identical for every ray, so only (1,1,D) is computed."""
        D = int(self.opt.z_depth_dim)
        jitter = 0.3 if getattr(self.opt, "is_train", False) else 0.0
        if self.inverse > 0:
            raise NotImplementedError("opt.inverse > 0 (disparity-linear candidates) is not used by any shipped config")
        tvals = torch.linspace(0, 1, D + 1, device=device).view(1, -1)
        tvals = near * (1 - tvals) + far * tvals
        rows = R if jitter > 0 else 1
        seg = (tvals[..., 1:] - tvals[..., :-1]) * (1 + jitter * (torch.rand((1, rows, D), device=device) - 0.5)) if jitter > 0 else \
              (tvals[..., 1:] - tvals[..., :-1]) * (1 + jitter * (torch.zeros((1, rows, D), device=device) - 0.5))
        end = torch.cumsum(seg, dim=2)
        end = torch.cat([torch.zeros((1, rows, 1), device=device), end], dim=2)
        end = near + end
        return ((end[:, :, :-1] + end[:, :, 1:]) / 2).contiguous()

    # ------------------------------------------------------------------ query
    def _buffers(self, R: int, SR: int, K: int, dev):
        key = (R, SR, K, str(dev))
        b = self._bufs.get(key)
        if b is None:
            i32 = dict(device=dev, dtype=torch.int32)
            b = dict(
                pidx_full=torch.empty((R, SR, K), **i32), nsamp=torch.empty(R, **i32), nvalid=torch.empty(R, **i32),
                keep=torch.empty(R, **i32), ray_off=torch.empty(R + 1, **i32), val_off=torch.empty(R + 1, **i32),
                scratch=torch.empty(int(lib().hnr_scan_scratch_elems(R)) + 1, **i32), counts=torch.empty(2, **i32),
                counts_host=[torch.empty(2, dtype=torch.int32).pin_memory() for _ in range(2)],
            )
            self._bufs = {key: b}          # keep only the latest shape
        return b

    def query_points(self, pixel_idx_tensor, point_xyz_pers_tensor, point_xyz_w_tensor, actual_numpoints_tensor, h, w, intrinsic,
                     near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor, cam_rot_tensor, ts: Optional[torch.Tensor] = None):
        """Same arguments as the reference (:80); `ts` optionally overrides the candidate parameters
        (tests feed the oracle and this kernel the same floats)."""
        return self.query_finish(self.query_launch(point_xyz_w_tensor, near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor, cam_rot_tensor, ts))

    def query_launch(self, point_xyz_w_tensor, near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor, cam_rot_tensor,
                     ts: Optional[torch.Tensor] = None) -> dict:
        """first half of query_points: enqueue the query kernels and the read-back of the two output counts, WITHOUT waiting for
        them.  A training loop launches the query of the NEXT frame before it issues the current frame's backward pass
        (NeuralPointsRayMarching.prefetch_query): the read-back has long completed when the next forward asks for it, so the host
        never idles the GPU at the query's synchronisation point."""
        opt = self.opt
        near, far = float(np.asarray(near_depth).item()), float(np.asarray(far_depth).item())
        G = self._ensure_grid(point_xyz_w_tensor)
        dev = point_xyz_w_tensor.device
        raydir = ray_dirs_tensor.reshape(-1, 3)
        if raydir.dtype != torch.float32 or not raydir.is_contiguous():
            raydir = raydir.float().contiguous()
        R, SR, K, D = raydir.shape[0], int(opt.SR), int(opt.K), int(opt.z_depth_dim)
        if ts is None:
            ts = self.candidate_ts(R, near, far, dev)
        ts = ts.reshape(-1, D).float().contiguous()
        ts_stride = D if ts.shape[0] > 1 else 0
        assert ts.shape[0] in (1, R)
        campos = cam_pos_tensor.reshape(-1)[:3].float().contiguous()
        camrot = cam_rot_tensor.reshape(-1)[:9].float().contiguous()
        b = self._buffers(R, SR, K, dev)
        f32 = dict(device=dev, dtype=torch.float32)
        sample_loc_full = torch.zeros((R, SR, 3), **f32)
        out_pidx = torch.empty((R, SR, K), device=dev, dtype=torch.int32)
        out_loc_pers = torch.empty((R, SR, 3), **f32)
        out_loc_w = torch.empty((R, SR, 3), **f32)
        out_dirs = torch.empty((R, SR, 3), **f32)
        ray_mask = torch.empty((R,), device=dev, dtype=torch.int8)
        ray_ids = torch.empty((R,), device=dev, dtype=torch.int32)
        vlist = torch.empty((R * SR,), device=dev, dtype=torch.int32)
        with ops.tag("query"), ops._launch(10):
          check(lib().hnr_query(ptr(campos), ptr(camrot), ptr(raydir), ptr(ts), ts_stride, R, D, SR, K, G.g, ptr(G.cell_start),
                              ptr(G.pts_sorted), ptr(G.occ_bits), ptr(sample_loc_full), ptr(b["pidx_full"]), ptr(b["nsamp"]),
                              ptr(b["nvalid"]), ptr(b["keep"]), ptr(b["ray_off"]), ptr(b["val_off"]), ptr(b["scratch"]), ptr(out_pidx),
                              ptr(out_loc_pers), ptr(out_loc_w), ptr(out_dirs), ptr(ray_mask), ptr(ray_ids), ptr(vlist), ptr(b["counts"]),
                              stream()), "query")
        # two pinned read-back slots alternate: a prefetched query may still be pending when the next one is launched
        self._slot = 1 - getattr(self, "_slot", 0)
        host = b["counts_host"][self._slot]
        host.copy_(b["counts"], non_blocking=True)
        ops.status_fetch_async(dev)                                # range guard of the previous frame's tensor-core kernels rides along
        ev = torch.cuda.Event()
        ev.record()
        return dict(event=ev, host=host, G=G, dev=dev, out=(out_pidx, out_loc_pers, out_loc_w, out_dirs, ray_mask), vlist=vlist, ray_ids=ray_ids)

    def query_finish(self, p: dict):
        """second half of query_points: wait for the read-back (the single host synchronisation of a query) and cut the outputs"""
        p["event"].synchronize()
        ops.status_check(p["dev"])
        Rk, Nv = int(p["host"][0]), int(p["host"][1])
        out_pidx, out_loc_pers, out_loc_w, out_dirs, ray_mask = p["out"]
        G = p["G"]
        self.last = QueryExtras(vlist=p["vlist"][:Nv], ray_ids=p["ray_ids"][:Rk], n_rays=Rk, n_valid=Nv)
        self.count += 1
        return (out_pidx[:Rk][None], out_loc_pers[:Rk][None], out_loc_w[:Rk][None], out_dirs[:Rk][None], ray_mask[None],
                G.vsize_np, G.ranges_np)

    def w2pers(self, point_xyz_w, camrotc2w, campos):
        """(:96-103) kept for API compatibility (plain tensor math on whatever device the inputs are)."""
        xyz_w_shift = point_xyz_w - campos[:, None, :]
        xyz_c = torch.sum(xyz_w_shift[..., None, :] * torch.transpose(camrotc2w, 1, 2)[:, None, None, ...], dim=-1)
        return torch.stack([xyz_c[..., 0] / xyz_c[..., 2], xyz_c[..., 1] / xyz_c[..., 2], xyz_c[..., 2]], dim=-1)
