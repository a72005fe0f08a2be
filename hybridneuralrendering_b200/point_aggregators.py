"""PointAggregator -- host-side mirror of models/aggregators/point_aggregators.py.

Same constructor ``PointAggregator(opt)``, same sub-module names and layer shapes (``block1``,
``block3``, ``alpha_branch``, ``color_branch``, ``color_feature_branch``, ``aux_merge_weight_block``,
``aux_block_s1/s2/s3``, ``color_mixup_block``, ``color_final_block``) so ``state_dict``s round-trip
with the reference, same ``forward`` signature and return tuple (:1427-1522).

The arithmetic runs in libhnr kernels: neighbour features are generated straight from the point
tables (fused path) or from the gathered tensors (drop-in ``forward``), the dense layers run on the
hand-written tcgen05 kernels (fused per-neighbour MLP ``mlp_tc`` / ``ops.NbrMlpTrainFn``, per-sample
chains ``chain``), the image features are read from the conv pyramid (own kernels, ``csrc/pyramid.cu``)
with the bilinear-upsample arithmetic folded in.  The cuDNN pyramid (``_ExactConvPyramid``) and the
exact-fp32 layer kernels (``ops.linear``, ``mlp_engine = "simt"``) are kept only as the tests'
independent cross-checks of those kernels.

Supported configuration = the one every shipped script uses: ``viewmlp``,
``agg_intrp_order=2``, ``agg_distance_kernel=linear``, ``agg_dist_pers=20``, ``LeakyReLU``,
``feature_guidance``, ``use_delta_view``, ``mixup_mode=partial``, ``learn_residuals``; anything else
raises instead of silently computing something different.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ACT_COLOR, ACT_LRELU, ACT_NONE, ACT_SIGMOID, X5_W


def drop_patch_rays(patch_size, patch_num, drop_ratio):
    """raster positions of the first floor(P^2*ratio) patches (reference :14-23)."""
    S = patch_size * patch_num
    n = int(patch_num * patch_num * drop_ratio)
    rows, cols = divmod(n, patch_num)
    flag = np.zeros((S, S), dtype=bool)
    flag[: rows * patch_size, :] = True
    flag[rows * patch_size:(rows + 1) * patch_size, : cols * patch_size] = True
    return np.nonzero(flag.reshape(-1))[0]


def _mlp(sizes, act_last=True, final: Optional[nn.Module] = None):
    layers = []
    for i in range(len(sizes) - 1):
        layers.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2 or act_last:
            layers.append(nn.LeakyReLU(inplace=True))
    if final is not None:
        layers.append(final)
    return nn.Sequential(*layers)


def _conv_block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=2, padding=1), nn.LeakyReLU(inplace=True),
                         nn.Conv2d(cout, cout, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(inplace=True))


def _init_seq(s):
    """Xavier-uniform with the gain of the following activation, zero bias (helpers/networks.py init_seq)."""
    mods = list(s)
    for a, b in zip(mods[:-1], mods[1:]):
        if isinstance(a, (nn.Linear, nn.Conv2d)):
            gain = nn.init.calculate_gain('leaky_relu', b.negative_slope) if isinstance(b, nn.LeakyReLU) else 1.0
            nn.init.xavier_uniform_(a.weight, gain=gain)
            nn.init.zeros_(a.bias)
    if isinstance(mods[-1], (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(mods[-1].weight)
        nn.init.zeros_(mods[-1].bias)


def _pyramid(x, ps):
    outs = []
    for i in range(3):
        w0, b0, w1, b1 = ps[4 * i:4 * i + 4]
        x = F.leaky_relu(F.conv2d(x, w0, b0, stride=2, padding=1), 0.01)
        x = F.leaky_relu(F.conv2d(x, w1, b1, stride=1, padding=1), 0.01)
        outs.append(x)
    return outs


class _ExactConvPyramid(torch.autograd.Function):
    """the six pyramid convolutions (reference :598-630, :1059-1063) through cuDNN in exact fp32, forward
    AND backward: cuDNN's default TF32 path (10-bit mantissa) would break the rtol 1e-4 bar, and a plain
    `cudnn.flags` context around the forward would not cover the backward pass."""

    @staticmethod
    def forward(ctx, x, *params):
        with torch.enable_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ps = [p.detach().requires_grad_(True) for p in params]
            outs = _pyramid(x.detach(), ps)
        ctx.ps, ctx.outs = ps, outs
        return tuple(o.detach() for o in outs)

    @staticmethod
    def backward(ctx, *gs):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            grads = torch.autograd.grad(ctx.outs, ctx.ps, [g.contiguous() for g in gs], allow_unused=True)
        return (None, *grads)


class PointAggregator(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self._check_supported(opt)
        Fd, H = opt.point_features_dim, opt.shading_feature_num
        self.dist_dim = 6
        in0 = Fd + 2 * opt.num_feat_freqs * Fd + 2 * abs(opt.dist_xyz_freq) * self.dist_dim          # 284
        self.viewdir_channels = 2 * opt.num_viewdir_freqs * 3                                        # 24
        self.block1 = _mlp([in0] + [H] * opt.shading_feature_mlp_layer1)
        self.block3 = _mlp([H + 3 + 4] + [H] * opt.shading_feature_mlp_layer3)
        self.alpha_branch = nn.Sequential(nn.Linear(H, 1))
        cin = H + self.viewdir_channels
        self.color_branch = _mlp([cin, H // 2, H // 2, H // 2, 3], act_last=False)         # built, unused (as in the reference)
        self.color_feature_branch = _mlp([cin, H // 2, H // 2, H // 2])
        aux_c = 45
        self.aux_merge_weight_block = _mlp([aux_c + H // 2 + 3, H // 4, H // 4, H // 4, 1], act_last=False, final=nn.Sigmoid())
        # the reference uses non-inplace LeakyReLU in this block; numerically identical
        self.aux_block_s1, self.aux_block_s2, self.aux_block_s3 = _conv_block(3, 6), _conv_block(6, 12), _conv_block(12, 24)
        self.color_mixup_block = _mlp([2 * aux_c, aux_c, aux_c, aux_c], act_last=False)
        self.color_final_block = nn.Sequential(nn.Linear(H // 2, 3))
        self.learn_blur_kernel_block = None
        if getattr(opt, "learnable_blur_kernel", 0):
            # blur-kernel predictor (reference :715-749): [gray gt patch | gray rendered patch] -> ks^2 taps (+1 combine weight)
            if getattr(opt, "learnable_blur_kernel_conv", 0):
                raise NotImplementedError("learnable_blur_kernel_conv=1 is not implemented (no shipped script selects it)")
            kk = int(getattr(opt, "learnable_blur_kernel_size", 9)) ** 2
            if int(getattr(opt, "learnable_blur_kernel_mode", 4)) in (2, 4):
                kk += 1
            self.learn_blur_kernel_block = _mlp([2 * int(getattr(opt, "learnable_blur_patch_size", 8)) ** 2, 128, 128, 128, kk],
                                                act_last=False, final=nn.Sigmoid())
            _init_seq(self.learn_blur_kernel_block)
        for m in (self.block1, self.block3, self.alpha_branch, self.color_branch, self.color_feature_branch,
                  self.aux_merge_weight_block, self.aux_block_s1, self.aux_block_s2, self.aux_block_s3, self.color_mixup_block,
                  self.color_final_block):
            _init_seq(m)
        self.shading_patch_size = 1
        # "tc": the product path -- fused tcgen05 kernels (3xFP16 forward, 3xBF16 backward).  "simt": exact-fp32 layer kernels, kept ONLY
        # as an independent cross-check of the tensor-core path in tests/ (not a user-selectable backend: there is one engine)
        self.mlp_engine = "tc"
        # valid samples decoded per pass in no-grad mode (bounds activation memory); HNR_MAX_VALID_CHUNK = A/B override
        self.max_valid_chunk = int(os.environ.get("HNR_MAX_VALID_CHUNK", "262144"))
        # image pyramid on own kernels (csrc/pyramid.cu); False = the six convolutions through cuDNN in exact fp32 (_ExactConvPyramid, kept
        # as the cross-check of the own kernels in the tests)
        self.own_pyramid = os.environ.get("HNR_OWN_PYRAMID", "1") != "0"
        self.fused_train_forward = True      # graph-recording forwards of the per-neighbour stage also use the fused kernel
        # backward of the per-neighbour stage: fused data-gradient chain + one image-fed weight-gradient launch (nbr_bwd_f16.cu,
        # wgrad_img.cu); False = layer-by-layer tensor-core kernels (kept as the cross-check of the fused path in the tests)
        self.fused_backward = os.environ.get("HNR_FUSED_BWD", "1") != "0"

    @staticmethod
    def _check_supported(opt):
        req = dict(which_agg_model="viewmlp", agg_intrp_order=2, agg_distance_kernel="linear", agg_dist_pers=20,
                   act_type="LeakyReLU", shading_feature_mlp_layer1=2, shading_feature_mlp_layer2=0, shading_feature_mlp_layer3=2,
                   shading_alpha_mlp_layer=1, shading_color_mlp_layer=4, shading_feature_num=256, point_features_dim=32,
                   num_feat_freqs=3, dist_xyz_freq=5, num_viewdir_freqs=4, feature_guidance=1, use_delta_view=1,
                   mixup_mode="partial", learn_residuals=1, apply_pnt_mask=1, agg_weight_norm=1)
        bad = {k: getattr(opt, k, v) for k, v in req.items() if getattr(opt, k, v) != v}
        off = [k for k in ("tradition_attention", "refine_blend", "dynamic_weight", "add_idx", "separate_color_decoder",
                           "large_color_final_block", "use_2D_CNN", "disable_viewdirs", "disable_color_feature", "learnable_blur_kernel_conv",
                           "downweight_blurry_feats", "dist_xyz_deno") if getattr(opt, k, 0)]
        if getattr(opt, "xyz_grad", 0):
            raise NotImplementedError("xyz_grad > 0 is not implemented: no kernel produces a gradient w.r.t. the point positions (every shipped "
                                      "script trains with xyz_grad=0; the distance encodings and inverse-distance weights are treated as data)")
        if bad or off:
            raise NotImplementedError(f"PointAggregator: unsupported options {bad} / enabled flags {off}; only the shipped "
                                      "configuration (SURVEY.md §8d) is implemented, there is no fallback path")

    # ------------------------------------------------------------------ reference helpers kept for API parity
    def raw2out_density(self, raw_density):
        return F.softplus(raw_density - 1)

    def raw2out_color(self, raw_color):
        return torch.sigmoid(raw_color) * (1 + 2 * 0.001) - 0.001

    def gradiant_clamp(self, sampled_conf, min=0.0001, max=1):
        diff = sampled_conf - torch.clamp(sampled_conf, min=min, max=max)
        return sampled_conf - diff.detach()

    # ------------------------------------------------------------------ image pyramid (own kernels; cuDNN = cross-check)
    def feature_pyramid(self, img_n: torch.Tensor):
        """img_n (1,V,H,W,3) -> NHWC levels [(V,H,W,3),(V,h1,w1,6),(V,h2,w2,12),(V,h3,w3,24)]"""
        img = img_n[0]
        params = []
        for blk in (self.aux_block_s1, self.aux_block_s2, self.aux_block_s3):
            params += [blk[0].weight, blk[0].bias, blk[2].weight, blk[2].bias]
        if self.own_pyramid:
            imgc = img.contiguous()
            return [imgc] + list(ops.PyramidFn.apply(imgc, *params))
        s1, s2, s3 = _ExactConvPyramid.apply(img.permute(0, 3, 1, 2), *params)
        return [img.contiguous()] + [t.permute(0, 2, 3, 1).contiguous() for t in (s1, s2, s3)]

    # ------------------------------------------------------------------ core
    def _keep_mask(self, R: int, SR: int, vlist: torch.Tensor) -> Optional[torch.Tensor]:
        """train-time image-feature drop (:1222-1237): uint8 (Nv,) keep flags, or None."""
        opt = self.opt
        if not (getattr(opt, "is_train", False) and opt.drop_ratio > 0 and opt.random_position == 1 and opt.use_nearest > 0):
            return None
        if not (opt.ray_points and opt.drop_patch and opt.drop_disturb_range == 0):
            raise NotImplementedError("only the deterministic patch drop (ray_points=1, drop_patch=1, drop_disturb_range=0) is implemented")
        # the dropped-ray flags depend on the options only: built once per (R, setup, ratio, device).  Uploading them every step is
        # a pageable host-to-device copy, i.e. a host stall until the whole forward so far has executed (1.8 ms in the step trace)
        key = (R, opt.dilation_setup, float(opt.drop_ratio), str(vlist.device))
        cache = self.__dict__.setdefault("_keep_cache", {})
        keep_ray = cache.get(key)
        if keep_ray is None:
            toks = opt.dilation_setup.split('_')
            pos = drop_patch_rays(int(toks[1]), int(toks[0]), opt.drop_ratio)
            if len(pos) and pos.max() >= R:
                raise IndexError(f"index {int(pos.max())} is out of bounds for axis 0 with size {R}")   # what numpy raises in the reference
            flags = np.ones(R, dtype=np.uint8)
            flags[pos] = 0
            keep_ray = cache[key] = torch.from_numpy(flags).to(vlist.device)
        return keep_ray.index_select(0, vlist.long() // SR).contiguous()

    def _run(self, tables, pidx, mask, loc_pers, loc_w, raydirs, cam, R, SR, levels, xy, delta, vlist=None):
        """tables = (xyz (N,3), xyz_pers|None, emb (N,32), color (N,3), dir (N,3), conf (N,)).
        pidx (S,K) i32; mask (S,K) u8|None; loc_* / raydirs (S,3).  Returns decoded (S,4), valid (S) bool,
        weight (S,K), conf_coefficient (S,K)."""
        xyz, xyz_pers, emb, color, dirs, conf = tables
        S, K = pidx.shape
        with ops.tag("gather"):
            weight, confc, valid = ops.NbrWeightsFn.apply(xyz, conf, pidx, mask, loc_w)
        if vlist is None:
            vlist = torch.nonzero(valid, as_tuple=False).view(-1).to(torch.int32)        # sync (drop-in path only)
        Nv = vlist.shape[0]
        decoded = torch.zeros((S, 4), device=pidx.device, dtype=torch.float32)
        self._last = (pidx, mask, vlist)
        if Nv == 0:
            return decoded, valid.bool(), weight, confc
        args = (tables, pidx, mask, loc_pers, loc_w, raydirs, cam, R, SR, levels, xy, delta, weight, confc)
        if torch.is_grad_enabled():
            decoded = ops.ScatterRowsFn.apply(self._decode(vlist, *args), vlist.long(), S)
        elif Nv <= self.max_valid_chunk:
            decoded = decoded.index_copy(0, vlist.long(), self._decode(vlist, *args))
        else:
            # inference over many samples (full frames): bound the activation memory by walking the valid-sample
            # list in slices; sizes are known on the host, so no synchronisation is involved
            for v0 in range(0, Nv, self.max_valid_chunk):
                part = vlist[v0:v0 + self.max_valid_chunk]
                decoded.index_copy_(0, part.long(), self._decode(part, *args))
        return decoded, valid.bool(), weight, confc

    def _decode(self, vlist, tables, pidx, mask, loc_pers, loc_w, raydirs, cam, R, SR, levels, xy, delta, weight, confc):
        """[sigma, rgb] (len(vlist),4) of the valid samples in `vlist`"""
        opt = self.opt
        xyz, xyz_pers, emb, color, dirs, conf = tables
        S, K = pidx.shape
        Nv = vlist.shape[0]
        b1, b3 = self.block1, self.block3
        # per-neighbour MLP: the fused tensor-core kernel; graph-recording forwards run it in training mode (operands saved as split
        # images for the fused backward).  The layer kernels below serve the gathered-tensor drop-in forward (mask given) and the tests
        use_tc = self.mlp_engine == "tc" and not torch.is_grad_enabled() and K == 8 and mask is None
        if use_tc:
            from . import mlp_tc
            pack = self._packed_weights()
            with ops.tag("nbr_mlp"):
                sigma, X5 = mlp_tc.forward_f16(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, pack,
                                               self.alpha_branch[0].weight, self.alpha_branch[0].bias, pp=self._point_partial(emb, Nv))
        elif self.mlp_engine == "tc" and torch.is_grad_enabled() and K == 8 and mask is None and self.fused_train_forward:
            # training: the same fused kernel, with the four layers' activations saved for the tensor-core backward
            tp = self._train_packs()
            with ops.tag("nbr_mlp"):
                if self.fused_backward:
                    sigma, X5 = ops.NbrMlpTrainFn.apply(emb, color, dirs, confc, b1[0].weight, b1[0].bias, b1[2].weight, b1[2].bias,
                                                        b3[0].weight, b3[0].bias, b3[2].weight, b3[2].bias, self.alpha_branch[0].weight,
                                                        self.alpha_branch[0].bias,
                                                        (xyz, xyz_pers, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight,
                                                         tp.nbr_pack, tp.nbr_packT))
                else:
                    sigma, X5 = ops.NbrMlpFusedFn.apply(emb, color, dirs, confc, b1[0].weight, b1[0].bias, b1[2].weight, b1[2].bias,
                                                        b3[0].weight, b3[0].bias, b3[2].weight, b3[2].bias, self.alpha_branch[0].weight,
                                                        self.alpha_branch[0].bias,
                                                        (xyz, xyz_pers, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, tp.nbr_pack))
        else:
            with ops.tag("gather"):
                X0, E = ops.NbrFeaturesFn.apply(emb, color, dirs, xyz, xyz_pers, pidx, mask, vlist, loc_w, loc_pers, raydirs, cam)
            with ops.tag("nbr_mlp"):
                h = ops.linear([X0], b1[0].weight, b1[0].bias, ACT_LRELU)
                h = ops.linear([h], b1[2].weight, b1[2].bias, ACT_LRELU)
                h = ops.linear([h, E], b3[0].weight, b3[0].bias, ACT_LRELU)
                h = ops.linear([h], b3[2].weight, b3[2].bias, ACT_LRELU)
            with ops.tag("ksum"):
                sigma, X5 = ops.AlphaKSumFn.apply(h, confc, self.alpha_branch[0].weight, self.alpha_branch[0].bias, weight, vlist, raydirs, cam)
        cf = self.color_feature_branch
        V = int(opt.use_nearest)
        # per-sample MLPs: fused tensor-core chains (chain_f16.cu) for no-grad forwards, layer kernels otherwise
        fused = self.mlp_engine == "tc" and not torch.is_grad_enabled()
        if fused:
            from . import chain
            with ops.tag("sample_mlp"):
                pc = chain.packed_chain(self, "cf", [cf[0], cf[2], cf[4]], [ACT_LRELU] * 3, X5_W)
                g = chain.chain_forward(pc, [X5])[0]
        # graph-recording forwards: the same chain kernel with every layer's output kept for the tensor-core backward
        fused_t = self.mlp_engine == "tc" and torch.is_grad_enabled() and self.fused_train_forward and Nv >= 128
        if fused_t:
            from . import chain
            tp = self._train_packs()
            with ops.tag("sample_mlp"):
                # dX5: only the 256 K-sum columns carry a gradient (the last 24 are the view-direction encoding)
                g = chain.chain_train(tp.pc["cf"], [cf[0], cf[2], cf[4]], [ACT_LRELU] * 3, [X5], pb=tp.pb["cf"])[0]
        elif not fused:
            with ops.tag("sample_mlp"):
                g = ops.linear([X5], cf[0].weight, cf[0].bias, ACT_LRELU)
                g = ops.linear([g], cf[2].weight, cf[2].bias, ACT_LRELU)
                g = ops.linear([g], cf[4].weight, cf[4].bias, ACT_LRELU)
        if V > 0 and fused:
            # inference: 48-wide rows ([aux 45 | dview 3], [merged 45 | 0 0 0]) keep every chain input 16-byte aligned
            with ops.tag("image_gather"):
                aux48, ok = ops.image_gather_padded(levels, xy, vlist, delta)
            am = self.aux_merge_weight_block
            with ops.tag("sample_mlp"):
                # first layer split (packer.AMG_COLS / AM_COLS48): W0g.g once per sample (one-layer chain), the chain over the V views
                # reads only the 48 view-dependent columns [aux | dview] and adds row (m mod Nv) of that product
                from .packer import AM_COLS48, AMG_COLS, _NoBias
                pcg = chain.packed_chain(self, "amg", [_NoBias(am[0])], [ACT_NONE], 128, cols0=AMG_COLS)
                G = chain.chain_forward(pcg, [g])[0]
                pc = chain.packed_chain(self, "am", [am[0], am[2], am[4]], [ACT_LRELU] * 3, 48, cols0=AM_COLS48)
                sig = chain.chain_forward(pc, [aux48.view(V * Nv, ops.AUX_LD)], M=V * Nv, out=False,
                                          head=(am[6].weight, am[6].bias, ACT_SIGMOID), add0=(G, Nv))[1]
            with ops.tag("blend"):
                merged = ops.blend_padded(aux48, sig, ok, self._keep_mask(R, SR, vlist))
        elif V > 0 and fused_t:
            # training: the same 48-wide aligned rows [aux 45 | dview 3] as in inference (the blend-weight chain reads them with vector
            # loads; 45-wide rows cost it 4x the load instructions); gradient rows are 48 wide, columns 45..47 are ignored
            with ops.tag("image_gather"):
                aux, ok = ops.ImageGatherFn.apply(levels[0], levels[1], levels[2], levels[3], xy, vlist, delta)
            am = self.aux_merge_weight_block
            with ops.tag("sample_mlp"):
                # first layer split as in inference: W0g.g once per sample (its gradient = the sum of dZ_0 over the V views), the chain
                # over the views reads / differentiates only the 48 view-dependent columns
                from .packer import AM_COLS48, AMG_COLS, _NoBias
                if "amg" in tp.pc:
                    G = chain.chain_train(tp.pc["amg"], [_NoBias(am[0])], [ACT_NONE], [g], cols0=AMG_COLS, pb=tp.pb["amg"])[0]
                    sig = chain.chain_train(tp.pc["am"], [am[0], am[2], am[4]], [ACT_LRELU] * 3, [aux.view(V * Nv, ops.AUX_LD)], M=V * Nv,
                                            head=(am[6], ACT_SIGMOID), cols0=AM_COLS48, pb=tp.pb["am"], add0=(G, Nv))[1]
                else:       # layer-by-layer backward (cross-check): unsplit first layer, kernel source order [g | aux | dview]
                    sig = chain.chain_train(tp.pc["am"], [am[0], am[2], am[4]], [ACT_LRELU] * 3, [g, aux.view(V * Nv, ops.AUX_LD)], M=V * Nv,
                                            mods=(Nv, 0), head=(am[6], ACT_SIGMOID), cols0=self._AM_COLS0, pb=None)[1]
            with ops.tag("blend"):
                merged = ops.BlendFn.apply(aux, sig, ok, self._keep_mask(R, SR, vlist))
        elif V > 0:
            with ops.tag("image_gather"):
                aux, ok = ops.ImageGatherFn.apply(levels[0], levels[1], levels[2], levels[3], xy, vlist)
            dv = delta.reshape(V, S, 3).index_select(1, vlist.long()).reshape(V * Nv, 3)
            am = self.aux_merge_weight_block
            with ops.tag("sample_mlp"):
                if fused_t:
                    c0 = self._AM_COLS0                                                  # kernel source order [g | aux | dview]
                    sig = chain.chain_train(tp.pc["am"], [am[0], am[2], am[4]], [ACT_LRELU] * 3, [g, aux.view(V * Nv, 45), dv], M=V * Nv,
                                            mods=(Nv, 0, 0), head=(am[6], ACT_SIGMOID), cols0=c0, pb=tp.pb["am"])[1]
                else:
                    t = ops.linear([aux.view(V * Nv, 45), g, dv], am[0].weight, am[0].bias, ACT_LRELU, mods=(0, Nv, 0), M=V * Nv)
                    t = ops.linear([t], am[2].weight, am[2].bias, ACT_LRELU)
                    t = ops.linear([t], am[4].weight, am[4].bias, ACT_LRELU)
                    sig = ops.linear([t], am[6].weight, am[6].bias, ACT_SIGMOID)
            with ops.tag("blend"):
                merged = ops.BlendFn.apply(aux, sig, ok, self._keep_mask(R, SR, vlist))
        else:
            merged = torch.zeros((Nv, ops.AUX_LD if fused else 45), device=pidx.device, dtype=torch.float32)
        gi, gv = g[:, :45], g[:, 45:]
        cm = self.color_mixup_block
        with ops.tag("sample_mlp"):
            if fused:
                # sources g[:, :48] and merged (48 wide): the three padding columns of each meet zero weights
                pc = chain.packed_chain(self, "cm", [cm[0], cm[2], cm[4]], [ACT_LRELU, ACT_LRELU, ACT_NONE], 96,
                                        cols0=list(range(45)) + [-1] * 3 + list(range(45, 90)) + [-1] * 3)
                m = chain.chain_forward(pc, [g[:, :ops.AUX_LD], merged], res=gi)[0]
            elif fused_t:
                m = chain.chain_train(tp.pc["cm"], [cm[0], cm[2], cm[4]], [ACT_LRELU, ACT_LRELU, ACT_NONE], [gi, merged], res=gi, pb=tp.pb["cm"])[0]
            else:
                m = ops.linear([gi, merged], cm[0].weight, cm[0].bias, ACT_LRELU)
                m = ops.linear([m], cm[2].weight, cm[2].bias, ACT_LRELU)
                m = ops.linear([m], cm[4].weight, cm[4].bias, ACT_NONE, res=gi)
            rgb = ops.linear([m, gv], self.color_final_block[0].weight, self.color_final_block[0].bias, ACT_COLOR)
        return torch.cat([sigma, rgb], dim=-1)

    _AM_COLS0 = list(range(45, 173)) + list(range(45)) + [173, 174, 175]        # blend-weight net, kernel source order [g | aux | dview]

    def prepack(self):
        """Bring the tensor-core weight images of a graph-recording forward up to date NOW (one launch, packer.TrainPacker).  The
        conductor calls this BEFORE the query so that the launch is issued while the GPU still executes the previous step."""
        if torch.is_grad_enabled() and self.mlp_engine == "tc" and self.fused_train_forward and int(self.opt.K) == 8:
            self._train_packs()

    def _train_packs(self):
        """packs of the training forward / backward (built once with the tensor-op packers, then refreshed by ONE kernel launch
        whenever a parameter version changed)"""
        from .packer import TrainPacker
        tp = getattr(self, "_tp", None)
        if tp is None or not tp.current(self, self.fused_backward):
            tp = self._tp = TrainPacker(self, self.fused_backward)
        else:
            tp.refresh()
        return tp

    def prepare_views(self, img_n, c2w_n):
        """query-independent part of the image branch (feature pyramid I1, world->camera matrices of the reference views),
        so that the conductor can issue it before the query's read-back.  -> (levels, w2c)"""
        levels = self.feature_pyramid(img_n)
        w2c = torch.linalg.inv_ex(c2w_n.reshape(-1, 4, 4).float())[0]        # inv() reads its info flag back (host sync)
        return levels, w2c

    def _point_partial(self, emb, Nv: int):
        """per-point layer-0 partial of the per-neighbour MLP for no-grad forwards (mlp_tc.point_partial), cached per (embedding table,
        first-layer weight) version.  OFF by default (HNR_POINT_PARTIAL=1 enables it): measured on the 800x800 frame it is neutral
        within run-to-run noise (nbr_mlp 40.4 vs 41.6 ms on the same box, the whole frame 69.8 vs 71.8 ms with every other stage
        equally faster in that run) -- skipping 14 of layer 0's 18 generated chunks is paid back by the 1 KB gather per neighbour row
        in the layer-0 epilogue, and the table costs 1 KB per point.  Kept as a verified A/B (tests/test_gpu_mlp_tc.py)."""
        if os.environ.get("HNR_POINT_PARTIAL", "0") != "1" or emb.shape[-1] != 32:
            return None
        W1 = self.block1[0].weight
        N = emb.shape[0] if emb.dim() == 2 else emb.shape[-2]
        key = (emb.data_ptr(), emb._version, N, W1.data_ptr(), W1._version)
        ent = getattr(self, "_pp_cache", None)
        if ent is not None and ent[0] == key:
            return ent[1]
        if Nv * 8 < N // 4:        # building the table costs one 224 -> 256 layer per POINT: only when enough neighbour rows use it
            return None
        from . import mlp_tc
        self._pp_cache = (key, mlp_tc.point_partial(emb, W1))
        return self._pp_cache[1]

    def _packed_weights(self):
        """fp16 hi/lo images of block1/block3 for the fused tensor-core kernel, re-packed when a weight changes"""
        ps = [self.block1[0].weight, self.block1[2].weight, self.block3[0].weight, self.block3[2].weight,
              self.block1[0].bias, self.block1[2].bias, self.block3[0].bias, self.block3[2].bias]
        train = torch.is_grad_enabled()
        key = (self.mlp_engine, train) + tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_pack_key", None) != key:
            from . import mlp_tc
            # graph-recording forwards re-pack after every optimiser step: fixed weight scale, no host read-back
            self._wpack_cache = mlp_tc.pack_mlp_f16(self.block1, self.block3, weight_scale=mlp_tc.TRAIN_WEIGHT_SCALE if train else None)
            self._pack_key = key
        return self._wpack_cache

    def _packed_weights_bwd(self):
        """bf16 hi/lo images of W^T for the fused data-gradient chain, re-packed when a weight changes"""
        ps = [self.block1[0].weight, self.block1[2].weight, self.block3[0].weight, self.block3[2].weight]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_packT_key", None) != key:
            from . import mlp_tc
            self._wpackT_cache = mlp_tc.pack_mlp_bwd(self.block1, self.block3)
            self._packT_key = key
        return self._wpackT_cache

    def last_valid_neighbours(self) -> int:
        """number of valid (sample, neighbour) pairs of the last call (profiling only; syncs)"""
        pidx, mask, vlist = self._last
        rows = (mask if mask is not None else (pidx >= 0)).index_select(0, vlist.long())
        return int(rows.sum())

    # ------------------------------------------------------------------ drop-in forward (gathered tensors)
    def forward(self, sampled_color, sampled_Rw2c, sampled_dir, sampled_conf, sampled_embedding, sampled_xyz_pers, sampled_xyz,
                sample_pnt_mask, sample_loc, sample_loc_w, sample_ray_dirs, vsize, grid_vox_sz, aux_image=None, pixel_idx=None,
                img_n=None, vid_angle_n=None, sample_loc_i_n=None, delta_viewdir_n=None, frame_weight_n=None):
        opt = self.opt
        B, R, SR, K = sample_pnt_mask.shape
        assert B == 1, "batch size 1 (as everywhere in the reference)"
        in_shape = sample_loc_w.shape
        S = R * SR
        dev = sample_pnt_mask.device
        if S == 0:
            return self._pack(torch.zeros(in_shape[:-1] + (4,), device=dev), torch.zeros(in_shape[:-1], dtype=torch.bool, device=dev), None, None)
        if sampled_Rw2c.dim() != 2:
            raise NotImplementedError("per-point Rw2c is not supported (normview is off in every shipped config)")
        if sampled_conf is None or sampled_color is None or sampled_dir is None:
            raise NotImplementedError("point_{conf,color,dir}_mode must be '1' (shipped configuration)")
        f2 = lambda t, c: t.reshape(S * K, c)
        tables = (f2(sampled_xyz, 3), f2(sampled_xyz_pers, 3), f2(sampled_embedding, sampled_embedding.shape[-1]), f2(sampled_color, 3),
                  f2(sampled_dir, 3), sampled_conf.reshape(S * K))
        pidx = torch.arange(S * K, device=dev, dtype=torch.int32).view(S, K)
        mask = sample_pnt_mask.reshape(S, K).to(torch.uint8).contiguous()
        cam = ops.make_cam(torch.zeros(3, device=dev), torch.eye(3, device=dev), sampled_Rw2c)
        levels = xy = delta = None
        if opt.use_nearest > 0:
            V = int(opt.use_nearest)
            levels = self.feature_pyramid(img_n)
            xy = sample_loc_i_n.reshape(V, S, 2).float().contiguous()
            delta = delta_viewdir_n.reshape(V, S, 3).float()
        decoded, valid, weight, confc = self._run(tables, pidx, mask, sample_loc.reshape(S, 3).float().contiguous(),
                                                  sample_loc_w.reshape(S, 3).float().contiguous(),
                                                  sample_ray_dirs.reshape(S, 3).float().contiguous(), cam, R, SR, levels, xy, delta)
        if int(valid.sum()) == 0:   # mirrors the reference's early return (:1444-1446)
            return self._pack(decoded.view(in_shape[:-1] + (4,)), valid.view(in_shape[:-1]), None, None)
        return self._pack(decoded.view(in_shape[:-1] + (4,)), valid.view(in_shape[:-1]), weight.view(B, R, SR, K), confc.view(B, R, SR, K))

    def _pack(self, decoded, ray_valid, weight, confc):
        opt = self.opt
        if weight is not None and (opt.sparse_loss_weight <= 0) and ("conf_coefficient" not in opt.zero_one_loss_items) and opt.prob == 0:
            weight, confc = None, None
        if getattr(opt, "is_train", False):
            return decoded, ray_valid, weight, confc, self.learn_blur_kernel_block
        return decoded, ray_valid, weight, confc

    # ------------------------------------------------------------------ fused forward (point tables + indices)
    def forward_fused(self, points, sample_pidx, sample_loc, sample_loc_w, sample_ray_dirs, campos, camrotc2w, extras=None, img_n=None,
                      c2w_n=None, intrinsic_n=None, campos_n=None, views=None):
        """points: NeuralPoints.  sample_* as returned by NeuralPoints.query.  Projection into the
        reference views (P1) runs in-kernel.  Returns decoded (1,R,SR,4), ray_valid, weight, conf_coefficient."""
        opt = self.opt
        B, R, SR, K = sample_pidx.shape
        S = R * SR
        dev = sample_pidx.device
        if S == 0:
            return torch.zeros((1, 0, SR, 4), device=dev), torch.zeros((1, 0, SR), dtype=torch.bool, device=dev), None, None
        if points.Rw2c is not None and points.Rw2c.dim() != 2:
            raise NotImplementedError("per-point Rw2c is not supported")
        # (N, C) VIEWS of the (1, N, C) parameters: the backward of a view is a view, whereas `param[0]` (select) costs a zero fill and
        # a copy of the whole 39*N-float gradient tables in backward
        v2 = lambda p: p.view(p.shape[-2], p.shape[-1])
        tables = (points.xyz, None, v2(points.points_embeding), v2(points.points_color), v2(points.points_dir), points.points_conf.reshape(-1))
        cam = ops.make_cam(campos, camrotc2w, points.Rw2c)
        loc_w = sample_loc_w.reshape(S, 3)
        levels = xy = delta = None
        if opt.use_nearest > 0:
            levels, w2c = views if views is not None else self.prepare_views(img_n, c2w_n)
            xy, delta = ops.project_views(loc_w, w2c, intrinsic_n.reshape(3, 3), campos.reshape(-1)[:3], campos_n.reshape(-1, 3))
        decoded, valid, weight, confc = self._run(tables, sample_pidx.reshape(S, K), None, sample_loc.reshape(S, 3), loc_w,
                                                  sample_ray_dirs.reshape(S, 3), cam, R, SR, levels, xy, delta,
                                                  vlist=None if extras is None else extras.vlist)
        return decoded.view(1, R, SR, 4), valid.view(1, R, SR), weight.view(1, R, SR, K), confc.view(1, R, SR, K)
