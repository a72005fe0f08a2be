"""Host side of the fused tensor-core per-neighbour MLP (csrc/nbr_mlp_f16.cu forward, csrc/nbr_bwd_f16.cu + csrc/wgrad_img.cu
backward): weight packing, split-image helpers, launches.

Forward packing: every layer's (256, K) nn.Linear weight is (1) column-permuted / zero-padded to the K order the kernel generates
its operand chunks in, (2) multiplied by a power of two and split into fp16 `hi` and `lo = fp16(w - hi)`, and (3) tiled into the
canonical no-swizzle K-major UMMA layout, one 16 KB image [hi 8 KB | lo 8 KB] per 16-column chunk, so that the kernel fetches a
chunk with a single cp.async.bulk.  (The first-generation 3xTF32 kernel mlp_tc.cu was retired in round 2: the 3xFP16 path's range
is covered by tests/test_gpu_mlp_tc.py::test_f16_kernel_wide_range_embeddings_and_weights.)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import check, lib, ptr, stream

_PERM_CACHE = {}


def _permute_pad(W: torch.Tensor, cols: Sequence[int]) -> torch.Tensor:
    key = (str(W.device), tuple(cols))
    ent = _PERM_CACHE.get(key)
    if ent is None:         # built once per column order and device: creating a tensor from a Python list is a blocking host-to-device copy
        ent = _PERM_CACHE[key] = (torch.tensor([c if c >= 0 else 0 for c in cols], device=W.device, dtype=torch.long),
                                  torch.tensor([c < 0 for c in cols], device=W.device))
    idx, pad = ent
    return W.index_select(1, idx).masked_fill(pad[None, :], 0.0)


# =====================================================================================================
# forward: 3xFP16 (csrc/nbr_mlp_f16.cu)
# =====================================================================================================
KC16 = 16
ACT_SCALE = 64.0          # power-of-two input scale of every layer: keeps hi AND lo of O(1e-3..1e3) activations in fp16's normal range


def layer1_column_order_f16() -> List[int]:
    """kernel K order of block1's 284 inputs for 16-wide chunks: [emb 32 | per 8-channel block and octave f:
    sin x 8, cos x 8 | dist: f*12 + (sin|cos)*6 + j | 4 x pad].  Reference order (point_aggregators.py:931-939, networks.py:175-189):
    [emb 32 | (c*3+f)*2+{sin,cos} 192 | (j*5+f)*2+{sin,cos} 60]."""
    cols = list(range(32))
    for cblk in range(4):
        for f in range(3):
            for sc in range(2):
                for i in range(8):
                    cols.append(32 + ((cblk * 8 + i) * 3 + f) * 2 + sc)
    for idx in range(64):
        if idx < 60:
            f, rem = divmod(idx, 12)
            sc, j = divmod(rem, 6)
            cols.append(224 + (j * 5 + f) * 2 + sc)
        else:
            cols.append(-1)
    assert len(cols) == 288 and sorted(c for c in cols if c >= 0) == list(range(284))
    return cols


def split_f16(x: torch.Tensor):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi, lo


def pack_layer_f16(W: torch.Tensor, cols: Optional[Sequence[int]] = None, weight_scale: Optional[float] = None):
    """W (256, K) fp32 -> (uint8 image of Kp/16 chunks, weight scale).  The weights are multiplied by a power of two
    that brings max|w| just below 2^14, split into fp16 hi / lo and tiled per 16-column chunk as [hi 8 KB | lo 8 KB],
    each part [k block (2)][row group (32)][row (8)][8 halves] = canonical no-swizzle K-major UMMA core matrices."""
    W = W.detach().float()
    assert W.shape[0] == 256
    if cols is None:
        Kp = (W.shape[1] + KC16 - 1) // KC16 * KC16
        cols = list(range(W.shape[1])) + [-1] * (Kp - W.shape[1])
    Wp = _permute_pad(W, cols).contiguous()
    Kp = Wp.shape[1]
    assert Kp % KC16 == 0
    if weight_scale is None:
        wmax = float(Wp.abs().max())                      # host read-back: fine for inference (packed once per weight version)
        sw = 2.0 ** math.floor(math.log2(16384.0 / wmax)) if wmax > 0 else 1.0
    else:
        sw = float(weight_scale)                          # training re-packs every step: a fixed scale avoids the synchronisation
        # ... and the range check rides on the device status word (reported at the host's next synchronisation point)
        ops.status_word(W.device).bitwise_or_((Wp.abs().max() * sw > 60000.0).to(torch.int32) * 4)
    hi, lo = split_f16(Wp * sw)
    tile = lambda x: x.view(32, 8, Kp // KC16, 2, 8).permute(2, 3, 0, 1, 4)          # (C, 2, 32, 8, 8)
    img = torch.stack([tile(hi), tile(lo)], dim=1).contiguous()                        # (C, 2[hi|lo], 2, 32, 8, 8)
    return img.view(torch.uint8).reshape(-1), sw


TRAIN_WEIGHT_SCALE = 1024.0     # fixed power-of-two weight scale of graph-recording forwards: full hi/lo precision for 1e-4 <= |w| < 63


def pack_mlp_f16(block1, block3, act_scale: float = ACT_SCALE, weight_scale: Optional[float] = None):
    """-> (wpack uint8, bias (4,256) pre-scaled, mul [4] floats, scale0, scale2) for hnr_nbr_mlp_f16_forward"""
    l2 = list(range(256, 263)) + [-1] * 9 + list(range(256))        # extras chunk first, then the 256 hidden columns
    layers = [(block1[0], layer1_column_order_f16()), (block1[2], None), (block3[0], l2), (block3[2], None)]
    parts, sws = [], []
    for lin, cols in layers:
        img, sw = pack_layer_f16(lin.weight, cols, weight_scale)
        parts.append(img)
        sws.append(sw)
    wpack = torch.cat(parts).contiguous()
    assert wpack.numel() == lib().hnr_nbr_mlp_f16_packed_bytes()
    s_in = [act_scale] * 4
    mul = [s_in[l + 1] / (s_in[l] * sws[l]) for l in range(3)] + [1.0 / (s_in[3] * sws[3])]
    bias = torch.stack([lin.bias.detach().float() * (s_in[l + 1] if l < 3 else 1.0) for l, (lin, _) in enumerate(layers)]).contiguous()
    return wpack, bias, mul, s_in[0], s_in[2]


def point_partial(emb: torch.Tensor, W1: torch.Tensor) -> torch.Tensor:
    """per-point layer-0 partial pp (N, 256) = W1[:, :224] . [emb | PE(emb)] (reference column order, point_aggregators.py:931-939,
    networks.py:175-189: [emb 32 | (c*3+f)*2+{sin,cos} of 2^f e_c]), through the own 3xTF32 layer kernel.  Computed once per (weights,
    point set) for inference: the fused kernel then skips 14 of layer 0's 18 operand chunks (hnr_nbr_mlp_f16_forward_pp)."""
    with torch.no_grad():
        e = emb.detach().float().reshape(-1, emb.shape[-1])
        assert e.shape[1] == 32 and W1.shape == (256, 284)
        out = torch.empty((e.shape[0], 256), device=e.device, dtype=torch.float32)
        Wp = W1.detach()[:, :224].contiguous()
        freqs = torch.tensor([1.0, 2.0, 4.0], device=e.device)
        step = 1 << 18                                     # bounds the temporary [emb | PE(emb)] rows (224 floats per point)
        for i0 in range(0, e.shape[0], step):
            ec = e[i0:i0 + step]
            arg = ec[:, :, None] * freqs                   # (n, 32, 3)
            pe = torch.stack([torch.sin(arg), torch.cos(arg)], dim=-1).reshape(ec.shape[0], 192)
            out[i0:i0 + step] = ops.linear([torch.cat([ec, pe], dim=1)], Wp, None, ops.ACT_NONE)
        return out


def forward_f16(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, pack, w_alpha, b_alpha, debug: bool = False, pp=None):
    """fused gather + per-neighbour MLP + density head + weighted K-sum, 3xFP16 tensor-core kernel.
    Returns sigma (Nv,1), X5 (Nv,280) [, acts (4, Nv*8, 256) activations of the four layers, araw (Nv*8) density
    pre-activations when debug -- what the training forward saves for the backward pass]."""
    import ctypes as C
    xyz, xyz_pers, emb, color, dirs, _ = tables
    wpack, bias, mul, s0, s2 = pack
    Nv, K = vlist.shape[0], pidx.shape[1]
    sigma = torch.empty((Nv, 1), device=pidx.device, dtype=torch.float32)
    X5 = torch.empty((Nv, 280), device=pidx.device, dtype=torch.float32)
    dbg = torch.empty((4, Nv * K, 256), device=pidx.device, dtype=torch.float32) if debug else None
    araw = torch.empty((Nv * K,), device=pidx.device, dtype=torch.float32) if debug else None
    wa = w_alpha.detach().float().contiguous().view(-1)
    ba = b_alpha.detach().float().contiguous().view(-1)
    mul_c = (C.c_float * 4)(*mul)
    if pp is not None and not debug:
        # per-point layer-0 partial (inference): the kernel generates only the distance-encoding chunks of layer 0
        with ops._launch():
            check(lib().hnr_nbr_mlp_f16_forward_pp(ptr(xyz), ptr(xyz_pers), ptr(pp), ptr(color), ptr(dirs), ptr(pidx), ptr(vlist), ptr(loc_w),
                                                   ptr(loc_pers), ptr(raydirs), ptr(cam), ptr(weight), ptr(confc), ptr(wpack), ptr(bias), ptr(wa),
                                                   ptr(ba), mul_c, float(s0), float(s2), 1.0 / ACT_SCALE, Nv, K, ptr(sigma), ptr(X5),
                                                   ptr(ops.status_word(pidx.device)), stream()), "nbr_mlp_f16_forward_pp")
        return sigma, X5
    with ops._launch():
        check(lib().hnr_nbr_mlp_f16_forward(ptr(xyz), ptr(xyz_pers), ptr(emb), ptr(color), ptr(dirs), ptr(pidx), ptr(vlist), ptr(loc_w),
                                            ptr(loc_pers), ptr(raydirs), ptr(cam), ptr(weight), ptr(confc), ptr(wpack), ptr(bias), ptr(wa),
                                            ptr(ba), mul_c, float(s0), float(s2), 1.0 / ACT_SCALE, Nv, K, ptr(sigma), ptr(X5), ptr(dbg),
                                            ptr(araw), ptr(ops.status_word(pidx.device)), stream()), "nbr_mlp_f16_forward")
    return (sigma, X5, dbg, araw) if debug else (sigma, X5)


# =====================================================================================================
# fused training path (csrc/nbr_bwd_f16.cu, csrc/wgrad_img.cu): split images + W^T images
# =====================================================================================================
X0_IMG_W, E_IMG_W = 288, 16
X0_GRAD_W = 224           # leading reference columns of the layer-0 input that carry a gradient: [emb 32 | sin/cos(2^j emb) 192]


def rows_padded(rows: int) -> int:
    return (rows + 127) // 128 * 128


def image_empty(rows: int, cols: int, device) -> torch.Tensor:
    """uninitialised split image (csrc/img_common.cuh) of `rows` x `cols` values: 4 bytes per element, rows padded to 128"""
    return torch.empty(rows_padded(rows) * cols * 4, device=device, dtype=torch.uint8)


def image_to_dense(imgt: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """split image -> (rows, cols) fp32 (hi + lo); tests / debugging only"""
    rp = rows_padded(rows)
    v = imgt.view(torch.bfloat16).view(rp // 32, 2, cols // 8, 32, 8).float()
    return (v[:, 0] + v[:, 1]).permute(0, 2, 1, 3).reshape(rp, cols)[:rows]


def dense_to_image(x: torch.Tensor) -> torch.Tensor:
    """(rows, cols % 16 == 0) fp32 -> split image with zero padding rows; tests / debugging only"""
    rows, cols = x.shape
    rp = rows_padded(rows)
    xp = torch.zeros((rp, cols), device=x.device, dtype=torch.float32)
    xp[:rows] = x
    hi = xp.bfloat16()
    lo = (xp - hi.float()).bfloat16()
    t = lambda a: a.view(rp // 32, 32, cols // 8, 8).permute(0, 2, 1, 3)
    return torch.stack([t(hi), t(lo)], dim=1).contiguous().view(torch.uint8).reshape(-1)


def _pack_wT_bf16(W: torch.Tensor) -> torch.Tensor:
    """W (256 outputs n, K <= 256 input columns k) -> 16 chunk images of the B operand of dX = dZ . W: rows = k (zero padded to
    256), reduction index n in 16-wide chunks, bf16 hi | lo, canonical no-swizzle K-major UMMA core matrices"""
    W = W.detach().float()
    assert W.shape[0] == 256 and W.shape[1] <= 256
    B = torch.zeros((256, 256), device=W.device, dtype=torch.float32)
    B[:W.shape[1]] = W.t()
    hi = B.bfloat16()
    lo = (B - hi.float()).bfloat16()
    tile = lambda x: x.view(32, 8, 16, 2, 8).permute(2, 3, 0, 1, 4)                  # (chunk, k block, row group, row, 8)
    return torch.stack([tile(hi), tile(lo)], dim=1).contiguous().view(torch.uint8).reshape(-1)


def pack_mlp_bwd(block1, block3) -> torch.Tensor:
    """W^T images of the data-gradient chain in consumption order: W4, W3[:, :256], W2, W1[:, :224] (reference column order)"""
    parts = [_pack_wT_bf16(block3[2].weight), _pack_wT_bf16(block3[0].weight[:, :256]), _pack_wT_bf16(block1[2].weight),
             _pack_wT_bf16(block1[0].weight[:, :X0_GRAD_W])]
    out = torch.cat(parts).contiguous()
    assert out.numel() == lib().hnr_nbr_bwd_f16_packed_bytes()
    return out


def forward_f16_train(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, pack, w_alpha, b_alpha):
    """training forward of the fused per-neighbour stage.  Returns sigma (Nv,1), X5 (Nv,280), images dict {x0, e, h0..h3}, araw."""
    import ctypes as C
    xyz, xyz_pers, emb, color, dirs, _ = tables
    wpack, bias, mul, s0, s2 = pack
    Nv, K = vlist.shape[0], pidx.shape[1]
    dev = pidx.device
    rows = Nv * K
    sigma = torch.empty((Nv, 1), device=dev, dtype=torch.float32)
    X5 = torch.empty((Nv, 280), device=dev, dtype=torch.float32)
    imgs = {"x0": image_empty(rows, X0_IMG_W, dev), "e": image_empty(rows, E_IMG_W, dev)}
    for l in range(4):
        imgs[f"h{l}"] = image_empty(rows, 256, dev)
    araw = torch.empty((rows,), device=dev, dtype=torch.float32)
    wa = w_alpha.detach().float().contiguous().view(-1)
    ba = b_alpha.detach().float().contiguous().view(-1)
    mul_c = (C.c_float * 4)(*mul)
    with ops._launch():
        check(lib().hnr_nbr_mlp_f16_forward_train(ptr(xyz), ptr(xyz_pers), ptr(emb), ptr(color), ptr(dirs), ptr(pidx), ptr(vlist), ptr(loc_w),
                                                  ptr(loc_pers), ptr(raydirs), ptr(cam), ptr(weight), ptr(confc), ptr(wpack), ptr(bias),
                                                  ptr(wa), ptr(ba), mul_c, float(s0), float(s2), 1.0 / ACT_SCALE, Nv, K, ptr(sigma), ptr(X5),
                                                  ptr(imgs["x0"]), ptr(imgs["e"]), ptr(imgs["h0"]), ptr(imgs["h1"]), ptr(imgs["h2"]),
                                                  ptr(imgs["h3"]), ptr(araw), ptr(ops.status_word(dev)), stream()),
              "nbr_mlp_f16_forward_train")
    return sigma, X5, imgs, araw
