#!/usr/bin/env python
"""Per-layer check of the 3xFP16 fused per-neighbour kernel against fp64 torch on the same gathered features
(prints the error of every layer; used while bringing the kernel up)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def layer_errors(R=64, SR=24, empty=0.4, seed=3, N=4000, emb_range=None, weight_scale=1.0, return_status=False, with_pp=False):
    from helpers import build_aggregator, cuda
    from hybridneuralrendering_b200 import mlp_tc, ops
    from hybridneuralrendering_b200 import synthetic as syn
    from oracle import render_oracle as ro
    d = syn.render_stage_inputs(seed=seed, N=N, R=R, SR=SR, K=8, V=2, H=32, W=40, empty_frac=empty)
    if emb_range is not None:           # wider embeddings than the U(-0.5, 0.5) initialisation (trained checkpoints)
        d["emb"] = ((np.random.default_rng(seed + 1).random(d["emb"].shape, dtype=np.float32) * 2 - 1) * emb_range).astype(np.float32)
    P = ro.random_params(seed)
    if weight_scale != 1.0:
        for k in list(P):
            if k.split(".")[0] in ("block1", "block3") and k.endswith("weight"):
                P[k] = P[k] * weight_scale
    agg = build_aggregator(P, use_nearest=2)
    S = R * SR
    xyz, emb, color, dirs, conf = (cuda(d[k]) for k in ("xyz", "emb", "color", "dir", "conf"))
    pidx = cuda(d["sample_pidx"]).reshape(S, 8).contiguous()
    loc_w = cuda(d["sample_loc_w"]).reshape(S, 3).contiguous()
    loc_pers = cuda(d["sample_loc"]).reshape(S, 3).contiguous()
    raydirs = cuda(d["sample_ray_dirs"]).reshape(S, 3).contiguous()
    cam = ops.make_cam(cuda(d["campos"]), cuda(d["camrotc2w"]), None)
    tables = (xyz, None, emb, color, dirs, conf.reshape(-1))
    with torch.no_grad():
        weight, confc, valid = ops.NbrWeightsFn.apply(xyz, conf.reshape(-1), pidx, None, loc_w)
        vlist = torch.nonzero(valid).view(-1).to(torch.int32)
        X0, E = ops.NbrFeaturesFn.apply(emb, color, dirs, xyz, None, pidx, None, vlist, loc_w, loc_pers, raydirs, cam)
        pack = mlp_tc.pack_mlp_f16(agg.block1, agg.block3)
        torch.cuda.synchronize()
        ops.status_word(xyz.device).zero_()
        sigma, X5, dbg, araw_k = mlp_tc.forward_f16(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, pack,
                                            agg.alpha_branch[0].weight, agg.alpha_branch[0].bias, debug=True)
        torch.cuda.synchronize()
        f = lambda lin, x: torch.nn.functional.leaky_relu(torch.nn.functional.linear(x, lin.weight.double(), lin.bias.double()), 0.01)
        h1 = f(agg.block1[0], X0.double())
        h2 = f(agg.block1[2], h1)
        h3 = f(agg.block3[0], torch.cat([h2, E.double()], 1))
        h4 = f(agg.block3[2], h3)
        refs = [h1, h2, h3, h4]
        errs = []
        for l in range(4):
            got = dbg[l].double()
            errs.append((float((got - refs[l]).abs().max()), float(refs[l].abs().max())))
        # heads
        Nv = vlist.shape[0]
        wrow = (weight * confc).index_select(0, vlist.long()).double()               # (Nv, 8)
        araw = torch.nn.functional.linear(h4, agg.alpha_branch[0].weight.double(), agg.alpha_branch[0].bias.double()).view(Nv, 8)
        sig_ref = (wrow * torch.nn.functional.softplus(araw - 1)).sum(1, keepdim=True)
        x5_ref = (h4.view(Nv, 8, 256) * wrow[..., None]).sum(1)
        errs.append((float((sigma.double() - sig_ref).abs().max()), float(sig_ref.abs().max())))
        errs.append((float((X5[:, :256].double() - x5_ref).abs().max()), float(x5_ref.abs().max())))
        # view encoding columns against the exact path
        _, X5_ref = ops.AlphaKSumFn.apply(h4.float(), confc, agg.alpha_branch[0].weight, agg.alpha_branch[0].bias, weight, vlist, raydirs, cam)
        errs.append((float((X5[:, 256:] - X5_ref[:, 256:]).abs().max()), 1.0))
        errs.append((float((araw_k.double().view(Nv, 8) - araw).abs().max()), float(araw.abs().max())))
        if with_pp:
            # inference with the per-point layer-0 partial (hnr_nbr_mlp_f16_forward_pp): same heads against the same fp64 references
            pp = mlp_tc.point_partial(emb, agg.block1[0].weight)
            sig2, X52 = mlp_tc.forward_f16(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, pack,
                                           agg.alpha_branch[0].weight, agg.alpha_branch[0].bias, pp=pp)
            torch.cuda.synchronize()
            errs.append((float((sig2.double() - sig_ref).abs().max()), float(sig_ref.abs().max())))
            errs.append((float((X52[:, :256].double() - x5_ref).abs().max()), float(x5_ref.abs().max())))
            errs.append((float((X52[:, 256:] - X5_ref[:, 256:]).abs().max()), 1.0))
        status = int(ops.status_word(xyz.device)[0])
        ops.status_word(xyz.device).zero_()
    return (errs, status) if return_status else errs


if __name__ == "__main__":
    names = ["layer0", "layer1", "layer2", "layer3", "sigma", "ksum", "viewpe", "araw"]
    for cfg in [dict(R=3, SR=5, empty=0.0), dict(R=64, SR=24, empty=0.4), dict(R=300, SR=80, empty=0.2)]:
        errs = layer_errors(**cfg)
        print(cfg, " ".join(f"{n}: {e:.2e}/{s:.2e}" for n, (e, s) in zip(names, errs)), flush=True)
