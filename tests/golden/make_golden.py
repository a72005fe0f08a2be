"""Generate the committed golden fixtures by running the UNMODIFIED reference (CPU, fp32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Inputs are produced by this repo's seeded generators (``hybridneuralrendering_b200.synthetic``) and
weights by ``oracle.render_oracle.random_params`` (numpy PCG64), so only OUTPUTS of the reference
need to be stored.  The reference ships no tests of its own, so these files are what
pins the oracle (prompt ③ / SURVEY.md §8c).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import, render_oracle as ro          # noqa: E402
from hybridneuralrendering_b200 import synthetic as syn      # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from frame_cases import FRAME_CASES, FRAME_CASES_CPU        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = torch.from_numpy


def agg_case(name, R, SR, V, H, W, is_train, drop_ratio, dilation_setup, seed, with_grad):
    d = syn.render_stage_inputs(seed=seed, N=600, R=R, SR=SR, K=8, V=V, H=H, W=W, empty_frac=0.4)
    g = syn.gather_neighbours(d)
    opt = ref_import.shipped_opt(use_nearest=V, is_train=is_train, drop_ratio=drop_ratio, dilation_setup=dilation_setup)
    agg = ref_import.aggregator(opt)
    P = ro.random_params(seed=seed + 100)
    missing = agg.load_state_dict(P, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith("color_branch") for k in missing.missing_keys), missing
    leaf = {k: T(g[k]).clone().requires_grad_(with_grad) for k in ("sampled_embedding", "sampled_color", "sampled_dir", "sampled_conf")}
    img = T(d["images_nearest"]).clone()
    out = agg(leaf["sampled_color"], torch.eye(3), leaf["sampled_dir"], leaf["sampled_conf"], leaf["sampled_embedding"],
              T(g["sampled_xyz_pers"]), T(g["sampled_xyz"]), T(g["sample_pnt_mask"]), T(d["sample_loc"]), T(d["sample_loc_w"]),
              T(d["sample_ray_dirs"]), d["vsize"], 0, img_n=img, sample_loc_i_n=T(d["sample_loc_i_n"]),
              delta_viewdir_n=T(d["delta_viewdir_n"]), frame_weight_n=None, vid_angle_n=None)
    decoded, ray_valid, weight, conf_coef = out[:4]
    dr, drf = ref_import.rendering()
    vz = float(d["vsize"][2])
    # C1 prologue exactly as the conductor does it (neural_points_volumetric_model.py:331-339)
    sl = T(d["sample_loc"])
    rd = torch.cummax(sl[..., 2], dim=-1)[0]
    rd = torch.cat([rd[..., 1:] - rd[..., :-1], torch.full((1, R, 1), vz)], dim=-1)
    m = torch.logical_or(rd < 1e-8, rd > 2 * vz).to(torch.float32)
    rd = rd * (1.0 - m) + m * vz
    rd = rd * ray_valid.float()
    bg = torch.ones(1, 3)
    color, _, opacity, accT, bw, bgT, _ = dr.ray_march(rd, ray_valid, decoded, drf.radiance_render, drf.alpha_blend, bg)
    res = dict(decoded=decoded, ray_valid=ray_valid, weight=weight, conf_coefficient=conf_coef, ray_dist=rd,
               ray_color=color, opacity=opacity, acc_transmission=accT, blend_weight=bw, bg_transmission=bgT)
    if with_grad:
        rng = np.random.default_rng(seed + 7)
        gt = T(rng.random((1, R, 3), dtype=np.float32))
        v = conf_coef.clamp(1e-3, 1 - 1e-3)
        loss = torch.nn.functional.mse_loss(color, gt) + 1e-4 * torch.mean(torch.log(v) + torch.log(1 - v))
        loss.backward()
        res["loss"] = loss.detach()
        res["gt"] = gt
        for k, t in leaf.items():
            res["grad_" + k] = t.grad
        for k, p in agg.named_parameters():
            if p.grad is not None:
                res["gradP_" + k] = p.grad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in res.items()},
                        meta=np.array([R, SR, V, H, W, int(is_train), seed]), drop_ratio=np.float32(drop_ratio), dilation_setup=np.array(dilation_setup))
    print(name, "Nv", int(ray_valid.sum()), "decoded", tuple(decoded.shape))


KINK_EPS = 1e-5      # |pre-activation| below which a LeakyReLU unit counts as "on the kink": ~5x the largest pre-activation error of the split-precision tensor-core kernels (4e-6 of the layer scale, test_f16_kernel_layers_match_fp64), ~20x fp32 summation noise


def _fragile_rays(agg, R, SR, K, pnt_mask, run_forward):
    """rays that contain a hidden unit whose pre-activation lies within KINK_EPS of 0 in the reference's own forward.  LeakyReLU' jumps
    there by a factor 100, so two correct fp32 implementations that sum in a different order may take different slopes; the
    training golden excludes those rays from the colour loss (their gradient paths are then exactly zero in every implementation)."""
    import torch.nn as nn
    rec = []
    hooks = []
    for name in ("block1", "block3", "color_feature_branch", "aux_merge_weight_block", "color_mixup_block"):
        seq = getattr(agg, name)
        mods = list(seq)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU):
                hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: rec.append((name, out.detach().abs().amin(dim=-1).reshape(-1).clone()))))
    with torch.no_grad():
        ray_valid = run_forward()
    for h in hooks:
        h.remove()
    nbr_ray = torch.nonzero(pnt_mask.reshape(-1)).reshape(-1) // (SR * K)          # ray of every compacted neighbour row (:921-939)
    smp_ray = torch.nonzero(ray_valid.reshape(-1)).reshape(-1) // SR               # ray of every compacted valid-sample row
    fragile = torch.zeros(R, dtype=torch.bool)
    for name, mn in rec:
        rows = nbr_ray if name in ("block1", "block3") else smp_ray
        assert mn.numel() == rows.numel(), (name, mn.numel(), rows.numel())
        fragile[rows[mn < KINK_EPS]] = True
    return fragile


def agg_case_tables(name, R, SR, V, H, W, is_train, drop_ratio, dilation_setup, seed, empty_frac, N=600):
    """Shipped-shape aggregator cases (SR 24 / 80, V 8 / 4, dilation_setup 7_8_1_8 incl. the out-of-range drop quirk of SURVEY B.16,
    thousands of valid samples).  Same call into the unmodified reference as agg_case, but the differentiable leaves are the POINT
    TABLES (gathered with torch indexing, exactly like NeuralPoints.forward :708-720), so the stored gradients stay small:
    (N, C) per table + every aggregator parameter.  Training case: rays with a hidden unit on a LeakyReLU kink are masked out of
    the colour loss (`keep`, stored), see _fragile_rays."""
    d = syn.render_stage_inputs(seed=seed, N=N, R=R, SR=SR, K=8, V=V, H=H, W=W, empty_frac=empty_frac)
    opt = ref_import.shipped_opt(use_nearest=V, is_train=is_train, drop_ratio=drop_ratio, dilation_setup=dilation_setup)
    agg = ref_import.aggregator(opt)
    P = ro.random_params(seed=seed + 100)
    missing = agg.load_state_dict(P, strict=False)
    assert not missing.unexpected_keys, missing
    tab = {k: T(d[k]).clone().requires_grad_(is_train) for k in ("emb", "color", "dir", "conf")}
    idx = T(np.maximum(d["sample_pidx"], 0)).long()
    g = syn.gather_neighbours(d)

    def fwd():
        return agg(tab["color"][idx], torch.eye(3), tab["dir"][idx], tab["conf"][idx], tab["emb"][idx],
                   T(g["sampled_xyz_pers"]), T(g["sampled_xyz"]), T(g["sample_pnt_mask"]), T(d["sample_loc"]), T(d["sample_loc_w"]),
                   T(d["sample_ray_dirs"]), d["vsize"], 0, img_n=T(d["images_nearest"]).clone(), sample_loc_i_n=T(d["sample_loc_i_n"]),
                   delta_viewdir_n=T(d["delta_viewdir_n"]), frame_weight_n=None, vid_angle_n=None)

    keep = torch.ones(1, R, 1)
    if is_train:
        fragile = _fragile_rays(agg, R, SR, 8, T(g["sample_pnt_mask"]), lambda: fwd()[1])
        keep[0, fragile, 0] = 0.0
    out = fwd()
    decoded, ray_valid, weight, conf_coef = out[:4]
    dr, drf = ref_import.rendering()
    vz = float(d["vsize"][2])
    sl = T(d["sample_loc"])
    rd = torch.cummax(sl[..., 2], dim=-1)[0]
    rd = torch.cat([rd[..., 1:] - rd[..., :-1], torch.full((1, R, 1), vz)], dim=-1)
    m = torch.logical_or(rd < 1e-8, rd > 2 * vz).to(torch.float32)
    rd = rd * (1.0 - m) + m * vz
    rd = rd * ray_valid.float()
    color, _, opacity, accT, bw, bgT, _ = dr.ray_march(rd, ray_valid, decoded, drf.radiance_render, drf.alpha_blend, torch.ones(1, 3))
    res = dict(decoded=decoded, ray_valid=ray_valid, ray_color=color, opacity=opacity, bg_transmission=bgT)
    if is_train:
        rng = np.random.default_rng(seed + 7)
        gt = T(rng.random((1, R, 3), dtype=np.float32))
        v = conf_coef.clamp(1e-3, 1 - 1e-3)
        loss = torch.nn.functional.mse_loss(color * keep, gt * keep) + 1e-4 * torch.mean(torch.log(v) + torch.log(1 - v))
        loss.backward()
        res["loss"] = loss.detach()
        res["gt"] = gt
        res["keep"] = keep
        for k, t in tab.items():
            res["gradT_" + k] = t.grad
        for k, p in agg.named_parameters():
            if p.grad is not None:
                res["gradP_" + k] = p.grad
        kept_valid = int((ray_valid.float() * keep[..., 0:1].reshape(1, R, 1)).sum())
        print(name, "rays kept", int(keep.sum()), "of", R, "valid samples in kept rays", kept_valid)
        assert kept_valid >= 2000
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in res.items()},
                        meta=np.array([R, SR, V, H, W, int(is_train), seed, N]), drop_ratio=np.float32(drop_ratio),
                        dilation_setup=np.array(dilation_setup), empty_frac=np.float32(empty_frac))
    print(name, "Nv", int(ray_valid.sum()), "decoded", tuple(decoded.shape))


def misc_cases():
    dr, drf = ref_import.rendering()
    nw = ref_import.networks()
    rng = np.random.default_rng(3)
    # positional encoding
    x = T(rng.standard_normal((7, 5)).astype(np.float32))
    pe = dict(x=x, pe3=nw.positional_encoding(x, 3), pe4_ori=nw.positional_encoding(x[:, :3], 4, ori=True))
    # ray generation (no jitter and with jitter); the jitter stream is torch's CPU generator
    campos = T(np.array([[0.3, -0.1, 2.0]], np.float32))
    raydir = T(rng.standard_normal((1, 9, 3)).astype(np.float32))
    rp0, seg0, _, ts0 = dr.near_far_linear_ray_generation(campos, raydir, 40, near=0.1, far=8.0, jitter=0.0)
    torch.manual_seed(11)
    rp1, seg1, _, ts1 = dr.near_far_linear_ray_generation(campos, raydir, 40, near=2.0, far=6.0, jitter=0.3)
    torch.manual_seed(11)
    noise = torch.rand((1, 9, 40))
    gen = dict(campos=campos, raydir=raydir, raypos0=rp0, ts0=ts0, raypos1=rp1, ts1=ts1, noise=noise)
    # standalone ray_march with hand-made tricky values (zero sigma, huge sigma, invalid samples)
    feats = T(rng.random((1, 6, 9, 4)).astype(np.float32))
    feats[0, 0, :, 0] = 0.0
    feats[0, 1, :, 0] = 1e4
    feats[0, 2, 3:, 0] *= 50
    valid = T(rng.random((1, 6, 9)) > 0.3)
    dist = T((rng.random((1, 6, 9)) * 0.02).astype(np.float32)) * valid.float()
    bg = T(np.array([[1.0, 0.5, 0.25]], np.float32))
    fr = feats.clone().requires_grad_(True)
    o = dr.ray_march(dist, valid, fr, drf.radiance_render, drf.alpha_blend, bg)
    G = T(rng.standard_normal((1, 6, 3)).astype(np.float32))
    (o[0] * G).sum().backward()
    rm = dict(feats=feats, valid=valid, dist=dist, bg=bg, G=G, ray_color=o[0], opacity=o[2], accT=o[3], bw=o[4], bgT=o[5], grad_feats=fr.grad)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **{"pe_" + k: v.numpy() for k, v in pe.items()},
                        **{"gen_" + k: v.numpy() for k, v in gen.items()}, **{"rm_" + k: v.detach().numpy() for k, v in rm.items()})
    print("misc ok")


def projection_case():
    """w2iproject + delta-view loop of the conductor, called as unbound code paths:
    neural_points_volumetric_model.py:248-255 and :296-310 need the class, which imports cleanly
    with a pytorch_msssim stub."""
    import types
    try:
        NeuralPointsRayMarching = ref_import.import_with_stubs("models.neural_points_volumetric_model").NeuralPointsRayMarching
    except Exception as e:  # pragma: no cover
        print("projection golden skipped:", repr(e))
        return
    rng = np.random.default_rng(5)
    fr = syn.room_frame(H=48, W=64, V=3, patch_num=2, patch_size=2, seed=1)
    loc_w = T((rng.random((1, 4, 5, 3)) * np.array([6, 5, 3])).astype(np.float32))
    xy = [NeuralPointsRayMarching.w2iproject(None, loc_w[0], T(fr["intrinsic_nearest"][0]), T(fr["c2w_nearest"][0, v]), None) for v in range(3)]
    # the delta-view loop is inline in NeuralPointsRayMarching.forward (:296-310): run THOSE source lines, read from the reference file
    import inspect, textwrap, types
    src = inspect.getsource(NeuralPointsRayMarching.forward).splitlines()
    a = next(i for i, l in enumerate(src) if "delta_sample_viewdir_nearest = []" in l)
    b = next(i for i, l in enumerate(src) if "delta_sample_viewdir_nearest = torch.stack(delta_sample_viewdir_nearest)" in l)
    env = dict(torch=torch, sample_loc_w=loc_w, campos=T(fr["campos"]), campos_nearest=T(fr["campos_nearest"]),
               self=types.SimpleNamespace(opt=types.SimpleNamespace(use_nearest=3)))
    exec(textwrap.dedent("\n".join(src[a:b + 1])), env)
    np.savez_compressed(os.path.join(OUT, "proj.npz"), loc_w=loc_w.numpy(), intrinsic=fr["intrinsic_nearest"][0], c2w_n=fr["c2w_nearest"][0],
                        xy=torch.stack(xy).numpy(), campos=fr["campos"], campos_n=fr["campos_nearest"][0],
                        delta_view=env["delta_sample_viewdir_nearest"].numpy())
    print("proj ok")


def blur_case():
    """blur_update_output is a method; call it unbound on a namespace carrying the attributes it
    reads (base_rendering_model.py:677-786).  Its `.cuda()` on the kernels is patched to a no-op
    for this CPU run -- the arithmetic is unchanged."""
    import types
    try:
        BaseRenderingModel = ref_import.import_with_stubs("models.base_rendering_model").BaseRenderingModel
    except Exception as e:  # pragma: no cover
        print("blur golden skipped:", repr(e))
        return
    rng = np.random.default_rng(9)
    PN, PS, Nk = 3, 8, 10
    S = PN * PS
    pred = T(rng.random((1, S * S, 3), dtype=np.float32)).requires_grad_(True)
    gt = T(rng.random((1, S * S, 3), dtype=np.float32))
    k = np.zeros((Nk, 9, 9), np.float32)
    for n in range(Nk):                       # sparse line-like kernels, rows sum to 1
        taps = rng.integers(2, 9)
        ii = rng.integers(0, 9, taps); jj = rng.integers(0, 9, taps)
        k[n, ii, jj] = rng.random(taps).astype(np.float32) + 0.1
        k[n, 4, 4] = 0.5                       # centre tap as in the shipped line kernels: border normalisation never 0/0
        k[n] /= k[n].sum()
    # make the argmin non-trivial: gt of patch p := pred blurred with kernel (3p mod Nk) + small noise,
    # except the last patch whose gt is pred itself (identity candidate wins)
    import torch.nn.functional as F
    with torch.no_grad():
        img = pred.detach().reshape(S, S, 3).permute(2, 0, 1).clone()
        gimg = gt.reshape(S, S, 3).permute(2, 0, 1).clone()
        for pi in range(PN):
            for pj in range(PN):
                pnum = pi * PN + pj
                blk = img[:, pi * PS:(pi + 1) * PS, pj * PS:(pj + 1) * PS].reshape(3, 1, PS, PS)
                if pnum == PN * PN - 1:
                    tgt = blk
                else:
                    kk = T(k[(3 * pnum) % Nk])[None, None]
                    tgt = F.conv2d(blk, kk, padding=4) / F.conv2d(torch.ones_like(blk), kk, padding=4)
                gimg[:, pi * PS:(pi + 1) * PS, pj * PS:(pj + 1) * PS] = tgt.reshape(3, PS, PS) + 0.01 * gimg[:, pi * PS:(pi + 1) * PS, pj * PS:(pj + 1) * PS]
        gt = gimg.permute(1, 2, 0).reshape(1, S * S, 3).contiguous()
    ns = types.SimpleNamespace(dilation_PatchNum=PN, dilation_PatchSize=PS, gt_image=gt, output={"coarse_raycolor": pred},
                               xv_patches=[], yv_patches=[], blur_kernels=T(k)[None])
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **kw: self
    try:
        BaseRenderingModel.blur_update_output(ns)
    finally:
        torch.Tensor.cuda = orig_cuda
    out = ns.output["coarse_raycolor"]
    G = T(rng.standard_normal((1, S * S, 3)).astype(np.float32))
    (out * G).sum().backward()
    np.savez_compressed(os.path.join(OUT, "blur.npz"), pred=pred.detach().numpy(), gt=gt.numpy(), kernels=k[None], out=out.detach().numpy(),
                        G=G.numpy(), grad_pred=pred.grad.numpy(), meta=np.array([PN, PS, Nk]))
    print("blur ok")


def learnable_blur_case():
    """learnable_blur_update_output (base_rendering_model.py:827-1020) called unbound on a namespace, with the predictor MLP
    built by the reference's own PointAggregator(opt.learnable_blur_kernel=1) (point_aggregators.py:715-749).  One file, four
    option sets: the shipped one (mode 4, /sum, boundary 1) and the other norm / mode / boundary branches."""
    import types
    BaseRenderingModel = ref_import.import_with_stubs("models.base_rendering_model").BaseRenderingModel
    rng = np.random.default_rng(21)
    PN, PS, KS = 3, 8, 9
    S = PN * PS
    pred0 = rng.random((1, S * S, 3), dtype=np.float32)
    gt = T(rng.random((1, S * S, 3), dtype=np.float32))
    G = T(rng.standard_normal((1, S * S, 3)).astype(np.float32))
    res = dict(pred=pred0, gt=gt.numpy(), G=G.numpy(), meta=np.array([PN, PS, KS]))
    cases = [(4, 0, 1), (4, 0, 0), (0, 1, 2), (4, 1, 2)]            # (kernel_mode, kernel_norm, boundary_mode)
    res["cases"] = np.array(cases)
    for ci, (mode, norm, bmode) in enumerate(cases):
        opt = ref_import.shipped_opt(is_train=True, learnable_blur_kernel=1, learnable_blur_kernel_size=KS, learnable_blur_patch_size=PS,
                                     learnable_blur_kernel_mode=mode, learnable_blur_kernel_norm=norm, learnable_blur_kernel_conv=0,
                                     boundary_mode=bmode)
        agg = ref_import.aggregator(opt)
        blk = agg.learn_blur_kernel_block
        lins = [m for m in blk if isinstance(m, torch.nn.Linear)]
        with torch.no_grad():
            for m, (W, b) in zip(lins, ro.blur_predictor_params(100 + ci, PS, KS, mode)):
                m.weight.copy_(W)
                m.bias.copy_(b)
        pred = T(pred0.copy()).requires_grad_(True)
        ns = types.SimpleNamespace(dilation_PatchNum=PN, dilation_PatchSize=PS, gt_image=gt, output={"coarse_raycolor": pred},
                                   xv_patches=[], yv_patches=[], opt=opt)
        BaseRenderingModel.learnable_blur_update_output(ns, blk)
        out = ns.output["coarse_raycolor"]
        (out * G).sum().backward()
        res[f"c{ci}_out"] = out.detach().numpy()
        res[f"c{ci}_grad_pred"] = pred.grad.numpy().copy()
        for li, m in enumerate([m for m in blk if isinstance(m, torch.nn.Linear)]):
            res[f"c{ci}_gW{li}"] = m.weight.grad.numpy().copy()
            res[f"c{ci}_gb{li}"] = m.bias.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "blur_learn.npz"), **res)
    print("learnable blur ok")


def frame_case():
    """ScannetFtDataset.__getitem__ (data/scannet_ft_dataset.py:736-976), the UNMODIFIED method, called unbound on a namespace that
    carries the attributes it reads, over a temporary scene directory (lossless PNG payload under the .jpg names the method
    opens -- PIL detects the format from the content; LANCZOS resize to the same size is the identity)."""
    import random
    import tempfile
    import types
    from PIL import Image
    from torchvision import transforms
    mod = ref_import.import_with_stubs("data.scannet_ft_dataset")
    images, c2w, vids, K, train_ids, test_ids = syn.frame_scene()
    F, H, W = images.shape[:3]
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "scene", "exported", "color"))
        os.makedirs(os.path.join(tmp, "scene", "exported", "pose"))
        for i, v in enumerate(vids):
            with open(os.path.join(tmp, "scene", "exported", "color", f"{v}.jpg"), "wb") as f:
                Image.fromarray(images[i]).save(f, format="PNG")
            np.savetxt(os.path.join(tmp, "scene", "exported", "pose", f"{v}.txt"), c2w[i].astype(np.float64), fmt="%.18e")
        for name, split, idx, over, seed, bg in FRAME_CASES + FRAME_CASES_CPU:
            o = dict(use_frame_weight=0, weight_exp=1.0, dynamic_nearest=0, select_high_quality=0, downweight_blurry_feats=0,
                     random_sample_size=32, dilation_setup="8_8_1_8")
            o.update(over)
            opt = types.SimpleNamespace(**o)
            ns = types.SimpleNamespace(id_list=train_ids if split == "train" else test_ids, train_id_list=train_ids, data_dir=tmp, scan="scene",
                                       img_wh=(W, H), transform=transforms.ToTensor(), intrinsic=K.copy(), opt=opt, split=split,
                                       train_weight_list=None, total_num_image=vids[-1] + 1, step=5, near_far=[0.1, 8.0],
                                       bg_color=bg, blur_kernels=np.zeros((1, 9, 9), np.float32))
            random.seed(seed)
            np.random.seed(seed)
            item = mod.ScannetFtDataset.__getitem__(ns, idx)
            for k in ("pixel_idx", "raydir", "gt_image", "images_nearest", "c2w_nearest", "campos_nearest", "camrotc2w_nearest",
                      "vid_angle_nearest", "c2w", "campos", "camrotc2w", "middle", "near", "far", "bg_color", "intrinsic"):
                v = item[k]
                res[f"{name}_{k}"] = v.numpy() if torch.is_tensor(v) else np.asarray(v)
            res[f"{name}_meta"] = np.array([item["vid"], item["h"], item["w"], opt.use_nearest])
            res[f"{name}_after"] = np.array([random.random(), np.random.rand()])     # both streams must be left in the same state
            # images_nearest is large: keep which frames were chosen (exact match against the scene) instead of the pixels
            chosen = []
            imgs_n = res.pop(f"{name}_images_nearest")
            res[f"{name}_images_nearest_absmax"] = np.float32(np.abs(imgs_n).max())
            if opt.use_nearest <= 0:                                                    # zeroed placeholder frame (:849-851)
                assert imgs_n.shape[0] == 1 and not imgs_n.any()
                res[f"{name}_vid_nearest"] = np.array([0])
                continue
            for img in imgs_n:
                hit = [i for i in range(F) if np.array_equal(img, images[i].astype(np.float32) / np.float32(255))]
                assert len(hit) == 1, hit
                chosen.append(vids[hit[0]])
            res[f"{name}_vid_nearest"] = np.array(chosen)
    np.savez_compressed(os.path.join(OUT, "frame.npz"), **res)
    print("frame ok", {n: res[f"{n}_vid_nearest"].tolist() for n, *_ in FRAME_CASES + FRAME_CASES_CPU})


if __name__ == "__main__":
    # bit-reproducible files: torch's multi-threaded CPU index_put_(accumulate=True) -- the backward of the table gathers `tab[k][idx]`
    # of agg_case_tables -- adds with atomics in arrival order; deterministic mode makes it (and nothing else used here) sequential
    torch.use_deterministic_algorithms(True, warn_only=True)
    if "--only-frame" in sys.argv:
        frame_case()
        sys.exit(0)
    if "--only-shipped-shapes" in sys.argv:
        torch.set_num_threads(8)
        # shipped shapes (dev_scripts/*: SR 80 / V 4 synthetic, SR 24 / V 8 + dilation_setup 7_8_1_8 + drop 0.5 ScanNet); R = 1792 > 1759 =
        # the largest row index the reference's (misaligned) patch drop touches
        agg_case_tables("agg_eval_sr80", R=64, SR=80, V=4, H=60, W=80, is_train=False, drop_ratio=0.0, dilation_setup="7_8_1_8", seed=21, empty_frac=0.5)
        agg_case_tables("agg_train_sr24", R=1792, SR=24, V=8, H=48, W=64, is_train=True, drop_ratio=0.5, dilation_setup="7_8_1_8", seed=22, empty_frac=0.9)
        sys.exit(0)
    if "--only-proj" in sys.argv:
        projection_case()
        sys.exit(0)
    if "--only-learnable-blur" in sys.argv:
        learnable_blur_case()
        sys.exit(0)
    torch.set_num_threads(4)
    agg_case("agg_eval", R=12, SR=5, V=2, H=20, W=24, is_train=False, drop_ratio=0.0, dilation_setup="7_8_1_8", seed=1, with_grad=False)
    agg_case("agg_train", R=16, SR=6, V=3, H=24, W=20, is_train=True, drop_ratio=0.5, dilation_setup="2_2_1_8", seed=2, with_grad=True)
    misc_cases()
    projection_case()
    blur_case()
    learnable_blur_case()
    frame_case()
    torch.set_num_threads(8)
    agg_case_tables("agg_eval_sr80", R=64, SR=80, V=4, H=60, W=80, is_train=False, drop_ratio=0.0, dilation_setup="7_8_1_8", seed=21, empty_frac=0.5)
    agg_case_tables("agg_train_sr24", R=1792, SR=24, V=8, H=48, W=64, is_train=True, drop_ratio=0.5, dilation_setup="7_8_1_8", seed=22, empty_frac=0.9)
