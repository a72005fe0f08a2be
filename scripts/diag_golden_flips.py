"""Where do product gradients differ from the fp64 oracle at shipped shapes?  (diagnostic, GPU box)"""
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import build_aggregator, cuda
from hybridneuralrendering_b200 import synthetic as syn
from hybridneuralrendering_b200.diff_ray_marching import ray_march_from_depth
from oracle import render_oracle as ro
T = torch.from_numpy
torch.set_num_threads(16)

def case(R, SR, V, drop, setup, empty, seed=22, N=600, is_train=True):
    d = syn.render_stage_inputs(seed=seed, N=N, R=R, SR=SR, K=8, V=max(V, 1), H=48, W=64, empty_frac=empty)
    g = syn.gather_neighbours(d)
    P = ro.random_params(seed + 100)
    rng = np.random.default_rng(seed + 7)
    gt = rng.random((1, R, 3), dtype=np.float32)
    # ---- oracle fp64
    cfg = ro.AggCfg(use_nearest=V, is_train=is_train, drop_ratio=drop, dilation_setup=setup)
    c = lambda a: T(a).double()
    tab = {k: c(d[k]).clone().requires_grad_(True) for k in ("emb", "color", "dir", "conf")}
    idx = T(np.maximum(d["sample_pidx"], 0)).long()
    Pd = {k: v.double().clone().requires_grad_(True) for k, v in P.items()}
    out = ro.aggregate(Pd, cfg, tab["color"][idx], torch.eye(3, dtype=torch.float64), tab["dir"][idx], tab["conf"][idx], tab["emb"][idx], c(g["sampled_xyz_pers"]),
                       c(g["sampled_xyz"]), T(g["sample_pnt_mask"]), c(d["sample_loc"]), c(d["sample_loc_w"]), c(d["sample_ray_dirs"]),
                       img_n=c(d["images_nearest"]) if V else None, sample_loc_i_n=c(d["sample_loc_i_n"]) if V else None, delta_viewdir_n=c(d["delta_viewdir_n"]) if V else None)
    decoded, valid, w, cc = out
    rd = ro.ray_dist_from_depth(c(d["sample_loc"])[..., 2], valid, float(d["vsize"][2]))
    color = ro.ray_march(rd, valid, decoded, torch.ones(1, 3, dtype=torch.float64))[0]
    v = cc.clamp(1e-3, 1 - 1e-3)
    loss = torch.nn.functional.mse_loss(color, T(gt).double()) + 1e-4 * torch.mean(torch.log(v) + torch.log(1 - v))
    loss.backward()
    ref = {"gradT_" + k: t.grad.numpy() for k, t in tab.items()}
    ref.update({"gradP_" + k: p.grad.numpy() for k, p in Pd.items() if p.grad is not None})
    # ---- product
    agg = build_aggregator(P, use_nearest=V, is_train=is_train, drop_ratio=drop, dilation_setup=setup)
    tabg = {k: cuda(d[k]).clone().requires_grad_(True) for k in ("emb", "color", "dir", "conf")}
    idxg = cuda(np.maximum(d["sample_pidx"], 0)).long()
    outg = agg(tabg["color"][idxg], torch.eye(3).cuda(), tabg["dir"][idxg], tabg["conf"][idxg], tabg["emb"][idxg], cuda(g["sampled_xyz_pers"]),
               cuda(g["sampled_xyz"]), cuda(g["sample_pnt_mask"]), cuda(d["sample_loc"]), cuda(d["sample_loc_w"]), cuda(d["sample_ray_dirs"]),
               d["vsize"], 0, img_n=cuda(d["images_nearest"]), sample_loc_i_n=cuda(d["sample_loc_i_n"]), delta_viewdir_n=cuda(d["delta_viewdir_n"]))
    dg, vg, wg, ccg = outg[:4]
    colg, *_ = ray_march_from_depth(cuda(d["sample_loc"]), vg, dg, float(d["vsize"][2]), 1, torch.ones(1, 3).cuda())
    vv = ccg.clamp(1e-3, 1 - 1e-3)
    lossg = torch.nn.functional.mse_loss(colg, cuda(gt)) + 1e-4 * torch.mean(torch.log(vv) + torch.log(1 - vv))
    lossg.backward()
    got = {"gradT_" + k: t.grad.cpu().numpy() for k, t in tabg.items()}
    got.update({"gradP_" + k: p.grad.cpu().numpy() for k, p in agg.named_parameters() if p.grad is not None})
    fwd_err = float((dg.cpu().double() - decoded.detach()).abs().max())
    line = [f"R{R} SR{SR} V{V} drop{drop} {setup} empty{empty} train{int(is_train)} Nv={int(valid.sum())} fwd_maxerr={fwd_err:.1e}"]
    for key in ("gradT_emb", "gradT_color", "gradT_conf", "gradP_block1.0.weight", "gradP_color_feature_branch.0.weight", "gradP_aux_merge_weight_block.0.weight", "gradP_color_mixup_block.0.weight", "gradP_color_final_block.0.weight"):
        if key not in ref or key not in got: continue
        r, a = ref[key], got[key]; mx = np.abs(r).max()
        e = np.abs(a - r); bad = e > 1e-4 * np.abs(r) + 1e-4 * mx
        line.append(f"{key.split('_',1)[1]}: bad {int(bad.sum())}/{bad.size} max {e.max()/mx:.1e}")
    print(" | ".join(line), flush=True)

case(1792, 24, 8, 0.5, "7_8_1_8", 0.85)
case(1792, 24, 8, 0.0, "7_8_1_8", 0.85)
case(1792, 24, 0, 0.0, "7_8_1_8", 0.85)
case(256, 24, 8, 0.0, "7_8_1_8", 0.85)
case(1792, 24, 2, 0.0, "7_8_1_8", 0.85)
