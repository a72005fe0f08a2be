"""The reference's own training step (query -> NeuralPoints gather -> PointAggregator -> ray_march -> loss -> backward), timed.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py `--impl reference`, `cpu_baseline`, `reference_gpu`; never imported by the product).

What runs:
  * query: the reference's six CUDA kernels from its own source (oracle/_ref/ref_query_k8.cubin, oracle/ref_query_runner.py -- grid
    rebuilt on every call exactly like query_point_indices_worldcoords.py:616) when a GPU is present; the numpy oracle otherwise.
    The reference has no CPU query.
  * gather: NeuralPoints.forward :702-733 restated with the same torch ops (cat of the tables is skipped; index_select per table).
  * aggregation + compositing: the UNMODIFIED reference classes (models/aggregators/point_aggregators.py:1427-1522,
    models/rendering/diff_ray_marching.py:508-557) imported from /root/reference or from the copy staged under oracle/_ref/pyref
    (oracle/stage_reference.py; git-ignored, travels to the GPU box) -> kind "reference"; if neither exists, the oracle port
    (oracle/render_oracle.py, pinned to the reference by tests/golden) -> kind "port".
  * projections into the reference views: oracle restatement of neural_points_volumetric_model.py:248-310 (pinned by tests/golden/proj.npz).
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import query_oracle as qo
from . import ref_import
from . import render_oracle as ro

T = torch.from_numpy


def kind() -> str:
    return "reference" if ref_import.available() else "port"


class ReferenceStep:
    """configs[2]-shaped training step of the reference on `device` ("cpu": all host threads; "cuda": its ATen GPU path)."""

    def __init__(self, xyz: np.ndarray, att: Dict[str, np.ndarray], frame: Dict[str, np.ndarray], opt, P: Dict[str, torch.Tensor], device: str,
                 ray_ids: Optional[np.ndarray] = None, drop_ratio: float = 0.0):
        self.dev = torch.device(device)
        self.opt, self.frame = opt, frame
        self.xyz_np = xyz
        self.ids = np.arange(frame["raydir"].shape[1]) if ray_ids is None else np.asarray(ray_ids)
        V = int(opt.use_nearest)
        self.V = V
        c = lambda a: T(np.ascontiguousarray(a)).to(self.dev)
        self.tab = {k: c(att[k]).clone().requires_grad_(True) for k in ("emb", "color", "dir", "conf")}
        self.xyz = c(xyz)
        self.use_ref = ref_import.available()
        if self.use_ref:
            ropt = ref_import.shipped_opt(use_nearest=V, is_train=True, drop_ratio=drop_ratio, dilation_setup=str(opt.dilation_setup), SR=int(opt.SR))
            self.agg = ref_import.aggregator(ropt)
            self.agg.load_state_dict(P, strict=False)
            self.agg = self.agg.to(self.dev)
            self.dr, self.drf = ref_import.rendering()
            self.params = [p for p in self.agg.parameters()]
        else:
            self.P = {k: v.to(self.dev).clone().requires_grad_(True) for k, v in P.items()}
            self.cfg = ro.AggCfg(use_nearest=V, is_train=True, drop_ratio=drop_ratio, dilation_setup=str(opt.dilation_setup))
            self.params = list(self.P.values())
        self.gp = qo.grid_params(xyz, opt.vsize, opt.vscale, opt.kernel_size, opt.ranges, opt.radius_limit_scale)
        self.q = None

    # ------------------------------------------------------------------ query
    def query(self):
        """-> dict(sample_pidx (1,R'',SR,K), sample_loc_w (1,R'',SR,3), ray_mask (R,)) on self.dev; returns seconds"""
        opt, fr = self.opt, self.frame
        t0 = time.perf_counter()
        raydir = fr["raydir"][:, self.ids]
        ts = qo.candidate_ts(int(opt.z_depth_dim), float(fr["near"].min()), float(fr["far"].max()))[0, 0]
        if torch.cuda.is_available() and os.path.exists(os.path.join(os.path.dirname(__file__), "_ref", "ref_query_k8.cubin")):
            from . import ref_query_runner as rq
            g = torch.device("cuda")
            raypos = (T(fr["campos"]).to(g)[:, None, None, :] + T(np.ascontiguousarray(raydir)).to(g)[:, :, None, :] *
                      T(np.ascontiguousarray(ts)).to(g).view(1, 1, -1, 1)).contiguous()
            r = rq.reference_query(T(self.xyz_np).to(g)[None].contiguous(), raypos, self.gp, SR=int(opt.SR), K=int(opt.K), P=int(opt.P),
                                   max_o=int(opt.max_o), kernel_size=opt.kernel_size, query_size=opt.query_size)
            self.q = dict(sample_pidx=r["sample_pidx"].to(self.dev), sample_loc_w=r["sample_loc_w"].to(self.dev), ray_mask=r["ray_mask"][0].to(self.dev))
            self.query_kind = "reference CUDA kernels (oracle/_ref cubin)"
        else:
            r = qo.query(self.xyz_np, fr["campos"], fr["camrotc2w"], raydir, ts, vsize=opt.vsize, vscale=opt.vscale, kernel_size=opt.kernel_size,
                         query_size=opt.query_size, ranges=opt.ranges, radius_limit_scale=opt.radius_limit_scale, SR=opt.SR, K=opt.K, P=opt.P)
            self.q = dict(sample_pidx=T(r["sample_pidx"]).to(self.dev), sample_loc_w=T(r["sample_loc_w"]).to(self.dev),
                          ray_mask=T(r["ray_mask"].reshape(-1)).to(self.dev))
            self.query_kind = "numpy oracle"
        if self.dev.type == "cuda":
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    # ------------------------------------------------------------------ forward + backward
    def fwd_bwd(self):
        """one forward + loss + backward of the aggregation and compositing stages; returns (seconds fwd, seconds bwd, kept rays, loss)"""
        fr, q, dev = self.frame, self.q, self.dev
        sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
        for t in list(self.tab.values()) + self.params:
            t.grad = None
        sync()
        t0 = time.perf_counter()
        pidx = q["sample_pidx"]
        B, R, SR, K = pidx.shape
        c = lambda a: T(np.ascontiguousarray(a)).to(dev)
        campos, camrot = c(fr["campos"]), c(fr["camrotc2w"])
        # NeuralPoints.forward (:702-733)
        shift = self.xyz[None] - campos[:, None, :]
        xc = torch.sum(camrot[:, None, :, :] * shift[:, :, :, None], dim=-2)
        xyz_pers = torch.stack([xc[..., 0] / xc[..., 2], xc[..., 1] / xc[..., 2], xc[..., 2]], dim=-1)
        mask = pidx >= 0
        idx = torch.clamp(pidx, min=0).view(-1).long()
        g = lambda t: torch.index_select(t[None] if t.dim() == 2 else t, 1, idx).view(B, R, SR, K, -1)
        loc_w = q["sample_loc_w"]
        sh = loc_w - campos[:, None, None, :]
        lc = torch.sum(sh[..., None, :] * torch.transpose(camrot, 1, 2)[:, None, None, ...], dim=-1)
        sample_loc = torch.stack([lc[..., 0] / lc[..., 2], lc[..., 1] / lc[..., 2], lc[..., 2]], dim=-1)
        kept = torch.nonzero(q["ray_mask"] > 0).view(-1)
        dirs = c(fr["raydir"][:, self.ids])[:, kept][:, :, None, :].expand(-1, -1, SR, -1).contiguous()
        V = self.V
        xy = dv = img = None
        if V > 0:
            xy = ro.project_to_views(loc_w[0], c(fr["intrinsic_nearest"][0]), c(fr["c2w_nearest"][0, :V]))
            dv = ro.delta_viewdirs(loc_w[0], campos[0], c(fr["campos_nearest"][0, :V]))
            img = c(fr["images_nearest"][:, :V])
        vsize = np.asarray(self.opt.vsize, np.float32)
        args = (g(self.tab["color"]), torch.eye(3, device=dev), g(self.tab["dir"]), g(self.tab["conf"]), g(self.tab["emb"]),
                torch.index_select(xyz_pers, 1, idx).view(B, R, SR, K, 3), torch.index_select(self.xyz[None], 1, idx).view(B, R, SR, K, 3), mask,
                sample_loc, loc_w, dirs)
        if self.use_ref:
            out = self.agg(*args, vsize, 0, img_n=img, sample_loc_i_n=xy, delta_viewdir_n=dv, frame_weight_n=None, vid_angle_n=None)
        else:
            out = ro.aggregate(self.P, self.cfg, *args, img_n=img, sample_loc_i_n=xy, delta_viewdir_n=dv)
        decoded, ray_valid, _, cc = out[:4]
        vz = float(vsize[2])
        rd = ro.ray_dist_from_depth(sample_loc[..., 2], ray_valid, vz)
        bg = torch.ones(1, 3, device=dev)
        color = (self.dr.ray_march(rd, ray_valid, decoded, self.drf.radiance_render, self.drf.alpha_blend, bg)[0] if self.use_ref
                 else ro.ray_march(rd, ray_valid, decoded, bg)[0])
        gt = c(fr["gt_image"][:, self.ids])[:, kept]
        v = cc.clamp(1e-3, 1 - 1e-3)
        loss = torch.nn.functional.mse_loss(color, gt) + 1e-6 + 1e-4 * torch.mean(torch.log(v) + torch.log(1 - v))
        sync()
        t1 = time.perf_counter()
        loss.backward()
        sync()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, int(R), float(loss.detach())


def config0_reference(device: str = "cpu", repeats: int = 3, warmup: int = 1):
    """BASELINE.json configs[0] / BASELINE.md §3: the render stage alone -- PointAggregator.forward + ray_dist + ray_march, forward
    AND backward -- of the unmodified reference on synthetic 200K neural points (32 channels), 1024 rays x 80 samples, K = 8
    PRECOMPUTED neighbours (50 % of the samples empty), V = 4 reference images 120x160 (hybridneuralrendering_b200.synthetic.
    render_stage_inputs(seed=0), SURVEY.md §8d recipe).  -> dict(rays, s_fwd, s_bwd (medians), rays/s, valid samples / neighbours, kind)"""
    from hybridneuralrendering_b200 import synthetic as syn
    dev = torch.device(device)
    d = syn.render_stage_inputs(seed=0)
    g = syn.gather_neighbours(d)
    P = ro.random_params(0)
    R, SR, V = d["sample_pidx"].shape[1], d["sample_pidx"].shape[2], 4
    c = lambda a: T(np.ascontiguousarray(a)).to(dev)
    use_ref = ref_import.available()
    if use_ref:
        agg = ref_import.aggregator(ref_import.shipped_opt(use_nearest=V, is_train=False))
        agg.load_state_dict(P, strict=False)
        agg = agg.to(dev)
        dr, drf = ref_import.rendering()
        params = list(agg.parameters())
    else:
        Pd = {k: v.to(dev).clone().requires_grad_(True) for k, v in P.items()}
        cfg = ro.AggCfg(use_nearest=V)
        params = list(Pd.values())
    leaf = {k: c(g[k]).requires_grad_(True) for k in ("sampled_embedding", "sampled_color", "sampled_dir", "sampled_conf")}
    fixed = [c(g["sampled_xyz_pers"]), c(g["sampled_xyz"]), c(g["sample_pnt_mask"]), c(d["sample_loc"]), c(d["sample_loc_w"]), c(d["sample_ray_dirs"])]
    img, xy, dv = c(d["images_nearest"]), c(d["sample_loc_i_n"]), c(d["delta_viewdir_n"])
    gt = c(np.random.default_rng(7).random((1, R, 3), dtype=np.float32))
    vz = float(d["vsize"][2])
    sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
    tf, tb = [], []
    for i in range(warmup + repeats):
        for t in list(leaf.values()) + params:
            t.grad = None
        sync()
        t0 = time.perf_counter()
        a = (leaf["sampled_color"], torch.eye(3, device=dev), leaf["sampled_dir"], leaf["sampled_conf"], leaf["sampled_embedding"], *fixed)
        if use_ref:
            out = agg(*a, d["vsize"], 0, img_n=img, sample_loc_i_n=xy, delta_viewdir_n=dv, frame_weight_n=None, vid_angle_n=None)
        else:
            out = ro.aggregate(Pd, cfg, *a, img_n=img, sample_loc_i_n=xy, delta_viewdir_n=dv)
        decoded, ray_valid = out[0], out[1]
        rd = ro.ray_dist_from_depth(fixed[3][..., 2], ray_valid, vz)
        bg = torch.ones(1, 3, device=dev)
        color = dr.ray_march(rd, ray_valid, decoded, drf.radiance_render, drf.alpha_blend, bg)[0] if use_ref else ro.ray_march(rd, ray_valid, decoded, bg)[0]
        loss = torch.nn.functional.mse_loss(color, gt)
        sync()
        t1 = time.perf_counter()
        loss.backward()
        sync()
        t2 = time.perf_counter()
        if i >= warmup:
            tf.append(t1 - t0); tb.append(t2 - t1)
    sf, sb = float(np.median(tf)), float(np.median(tb))
    return {"rays": R, "samples_per_ray": SR, "valid_samples": int(ray_valid.sum()), "valid_neighbours": int((d["sample_pidx"] >= 0).sum()),
            "s_fwd": sf, "s_bwd": sb, "value": R / (sf + sb), "unit": "rays/s", "kind": "reference" if use_ref else "port", "loss": float(loss.detach())}
