"""Run the REFERENCE's own compiled query kernels (oracle/_ref/ref_query_k8.cubin) on a GPU.

TEST INFRASTRUCTURE ONLY.  pycuda is not installable, so the cubin that oracle/build_ref_query_cubin.py
compiled from the reference's CUDA source is loaded with cuda.bindings and launched on torch's
primary context / current stream.  The host orchestration between the kernels (allocation, fills,
masked_select compaction, cumsum) restates models/neural_points/query_point_indices_worldcoords.py
:540-711 with the same torch ops, so that the output is what `query_grid_point_index` returns:
(sample_pidx (1,R'',SR,K) in the reference's slot order, sample_loc_w (1,R'',SR,3), ray_mask (1,R) int8),
plus the grid tables for diagnostics (which voxel won occupied-slot 0, per-voxel counts).
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Dict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(HERE, "_ref", "ref_query_k8.cubin")
NAMES = ("claim_occ", "map_coor2occ", "fill_occ2pnts", "mask_raypos", "get_shadingloc", "query_neigh_along_ray_layered")


def available() -> bool:
    return os.path.exists(CUBIN) and torch.cuda.is_available()


class RefKernels:
    def __init__(self):
        from cuda.bindings import driver as cu
        self.cu = cu
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make sure the primary context is current
        err, self.mod = cu.cuModuleLoadData(open(CUBIN, "rb").read())
        assert err == cu.CUresult.CUDA_SUCCESS, err
        self.fn = {}
        for n in NAMES:
            err, f = cu.cuModuleGetFunction(self.mod, n.encode())
            assert err == cu.CUresult.CUDA_SUCCESS, (n, err)
            self.fn[n] = f

    def launch(self, name, n_threads, *args, block=1024):
        """args: torch tensors (passed as device pointers) or ctypes scalars"""
        vals = []
        for a in args:
            vals.append(ctypes.c_void_p(a.data_ptr()) if torch.is_tensor(a) else a)
        ptrs = (ctypes.c_void_p * len(vals))(*[ctypes.addressof(v) for v in vals])
        grid = int((n_threads + block - 1) // block)
        (err,) = self.cu.cuLaunchKernel(self.fn[name], grid, 1, 1, block, 1, 1, 0, torch.cuda.current_stream().cuda_stream,
                                        ctypes.addressof(ptrs), 0)
        assert err == self.cu.CUresult.CUDA_SUCCESS, (name, err)


def reference_query(xyz: torch.Tensor, raypos: torch.Tensor, gp, *, SR: int, K: int, P: int, max_o: int, kernel_size, query_size,
                    NN: int = 2) -> Dict[str, torch.Tensor]:
    """xyz (1,N,3) cuda f32; raypos (1,R,D,3) cuda f32; gp: oracle.query_oracle.GridParams (the values
    get_hyperparameters produces)."""
    rk = RefKernels()
    dev = xyz.device
    i32, c_int, c_f = torch.int32, ctypes.c_int, ctypes.c_float
    B, N = xyz.shape[0], xyz.shape[1]
    R, D = raypos.shape[1], raypos.shape[2]
    dims = [int(v) for v in gp.dims]
    vol = dims[0] * dims[1] * dims[2]
    shift = torch.tensor(gp.origin, device=dev, dtype=torch.float32)
    vsz = torch.tensor(gp.cell, device=dev, dtype=torch.float32)
    gsz = torch.tensor(dims, device=dev, dtype=i32)
    ksz = torch.tensor([int(v) for v in kernel_size], device=dev, dtype=i32)
    qsz = torch.tensor([int(v) for v in query_size], device=dev, dtype=i32)
    npts = torch.full((B,), N, device=dev, dtype=i32)
    secs = ctypes.c_ulong(int(time.time()))
    # build_occ_vox (:540-602)
    coor_occ = torch.zeros([B] + dims, dtype=i32, device=dev)
    occ_2_pnts = torch.full([B, max_o, P], -1, dtype=i32, device=dev)
    occ_2_coor = torch.full([B, max_o, 3], -1, dtype=i32, device=dev)
    occ_numpnts = torch.zeros([B, max_o], dtype=i32, device=dev)
    coor_2_occ = torch.full([B] + dims, -1, dtype=i32, device=dev)
    occ_idx = torch.zeros([B], dtype=i32, device=dev)
    rk.launch("claim_occ", B * N, xyz, npts, c_int(B), c_int(N), shift, vsz, gsz, c_int(vol), c_int(max_o), occ_idx, coor_2_occ, occ_2_coor, secs)
    coor_2_occ = torch.full([B] + dims, -1, dtype=i32, device=dev)
    rk.launch("map_coor2occ", B * max_o, c_int(B), gsz, qsz, c_int(vol), c_int(max_o), occ_idx, coor_occ, coor_2_occ, occ_2_coor)
    rk.launch("fill_occ2pnts", B * N, xyz, npts, c_int(B), c_int(N), c_int(P), shift, vsz, gsz, c_int(vol), c_int(max_o), coor_2_occ, occ_2_pnts,
              occ_numpnts, secs)
    # query_grid_point_index (:623-711)
    raypos_mask = torch.zeros([B, R, D], dtype=i32, device=dev)
    rk.launch("mask_raypos", B * R * D, raypos, coor_occ, c_int(B), c_int(R), c_int(D), c_int(vol), shift, gsz, vsz, raypos_mask)
    ray_mask = torch.max(raypos_mask, dim=-1)[0] > 0
    R1 = int(torch.max(torch.sum(ray_mask.to(i32))).item())
    sample_loc = torch.zeros([B, R1, SR, 3], dtype=torch.float32, device=dev)
    sample_pidx = torch.full([B, R1, SR, K], -1, dtype=i32, device=dev)
    if R1 > 0:
        rp = torch.masked_select(raypos, ray_mask[..., None, None].expand(-1, -1, D, 3)).reshape(B, R1, D, 3)
        rm = torch.masked_select(raypos_mask, ray_mask[..., None].expand(-1, -1, D)).reshape(B, R1, D)
        cum = torch.cumsum(rm, dim=-1).to(i32)
        rm = ((rm * cum * (cum <= SR)) - 1).to(i32).contiguous()
        loc_mask = torch.zeros([B, R1, SR], dtype=i32, device=dev)
        rk.launch("get_shadingloc", B * R1 * D, rp.contiguous(), rm, c_int(B), c_int(R1), c_int(D), c_int(SR), sample_loc, loc_mask)
        r2 = np.float32(np.float32(np.sqrt(gp.radius2)) ** 2) if False else np.float32(gp.radius2)
        rk.launch("query_neigh_along_ray_layered", B * R1 * SR, xyz, c_int(B), c_int(SR), c_int(R1), c_int(max_o), c_int(P), c_int(K), c_int(vol),
                  c_f(float(r2)), shift, gsz, vsz, ksz, occ_numpnts, occ_2_pnts, coor_2_occ, sample_loc, loc_mask, sample_pidx, secs, c_int(NN))
        valid_ray = torch.sum(sample_pidx.view(B, R1, -1) >= 0, dim=-1) > 0
        R2 = int(torch.max(torch.sum(valid_ray.to(i32), dim=-1)).item())
        ray_mask.masked_scatter_(ray_mask, valid_ray)
        sample_pidx = torch.masked_select(sample_pidx, valid_ray[..., None, None].expand(-1, -1, SR, K)).reshape(B, R2, SR, K)
        sample_loc = torch.masked_select(sample_loc, valid_ray[..., None, None].expand(-1, -1, SR, 3)).reshape(B, R2, SR, 3)
    torch.cuda.synchronize()
    n_occ = int(occ_idx[0].item())
    slot0 = occ_2_coor[0, 0].tolist()
    return dict(sample_pidx=sample_pidx, sample_loc_w=sample_loc, ray_mask=ray_mask.to(torch.int8), n_occupied=n_occ,
                slot0_cell=(slot0[0] * dims[1] + slot0[1]) * dims[2] + slot0[2] if n_occ > 0 else -1,
                max_cell_count=int(occ_numpnts.max().item()))
