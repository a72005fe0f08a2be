"""Host side of the fused dense-layer chain on tensor cores (csrc/chain_f16.cu): weight packing + launch.

A chain is a list of nn.Linear layers (widths <= 128) applied to the concatenation of up to three fp32 row-major
sources; everything between the first input and the last output stays on the SM.  Packing (cached per weight
version): every W_l (N,K) is zero-padded to (Np % 16 == 0, Kp % 16 == 0), multiplied by a power of two that brings
max|w| just below 2^14, split into fp16 hi / lo = fp16(w - hi) and tiled per 16-column K chunk as
[hi | lo], each part [k block (2)][row group (Np/8)][row (8)][8 halves] (canonical no-swizzle K-major UMMA layout).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import check, i64_array, lib, ptr, ptr_array, stream

ACT_SCALE = 64.0      # power-of-two scale of every layer input (keeps fp16 hi AND lo of O(1e-3..1e3) values normal)
NMAX = 128


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


_IDX_CACHE = {}


def _cols_index(cols0, device):
    """(index tensor, 0/1 mask tensor) for a column order, cached per device: creating them from Python lists is a blocking
    host-to-device copy, which must not happen on every training step"""
    key = (str(device), tuple(cols0))
    ent = _IDX_CACHE.get(key)
    if ent is None:
        idx = torch.tensor([max(c, 0) for c in cols0], device=device)
        mask = torch.tensor([1.0 if c >= 0 else 0.0 for c in cols0], device=device)
        ent = _IDX_CACHE[key] = (idx, mask)
    return ent


def _cols_map32(cols0, kp, device):
    """device int32 map kernel input column -> reference column (-1 = padding), padded to kp entries; cached per device"""
    key = (str(device), tuple(cols0), kp, "map32")
    ent = _IDX_CACHE.get(key)
    if ent is None:
        ent = _IDX_CACHE[key] = torch.tensor(list(cols0) + [-1] * (kp - len(cols0)), device=device, dtype=torch.int32)
    return ent


def _cols_real(cols0, device):
    """(reference column, kernel column) index tensors of the real (non-padding) columns of a column order, cached per device"""
    key = (str(device), tuple(cols0), "real")
    ent = _IDX_CACHE.get(key)
    if ent is None:
        ent = _IDX_CACHE[key] = (torch.tensor([c for c in cols0 if c >= 0], device=device, dtype=torch.long),
                                 torch.tensor([k for k, c in enumerate(cols0) if c >= 0], device=device, dtype=torch.long))
    return ent


def pack_chain_layer(W: torch.Tensor, Kp: int, weight_scale: Optional[float] = None):
    W = W.detach().float()
    N, K = W.shape
    Np = _pad16(N)
    assert Np <= NMAX and K <= Kp and Kp % 16 == 0
    Wp = torch.zeros((Np, Kp), device=W.device, dtype=torch.float32)
    Wp[:N, :K] = W
    if weight_scale is None:
        wmax = float(Wp.abs().max())                      # host read-back: inference packs once per weight version
        sw = 2.0 ** math.floor(math.log2(16384.0 / wmax)) if wmax > 0 else 1.0
    else:                                                 # training re-packs every step: fixed scale, range check on the status word
        sw = float(weight_scale)
        ops.status_word(W.device).bitwise_or_((Wp.abs().max() * sw > 60000.0).to(torch.int32) * 4)
    Ws = Wp * sw
    hi = Ws.half()
    lo = (Ws - hi.float()).half()
    tile = lambda x: x.view(Np // 8, 8, Kp // 16, 2, 8).permute(2, 3, 0, 1, 4)
    img = torch.stack([tile(hi), tile(lo)], dim=1).contiguous().view(torch.uint8).reshape(-1)     # (C, hi|lo, 2, Np/8, 8, 8)
    assert img.numel() == (Kp // 16) * lib().hnr_chain_f16_chunk_bytes(Np)
    return img, sw, Np


class PackedChain:
    """packed weights + scale bookkeeping of a chain; rebuilt when any weight / bias changes"""

    def __init__(self, layers: Sequence[torch.nn.Linear], acts: Sequence[int], k_in: int, act_scale: float = ACT_SCALE,
                 cols0: Optional[Sequence[int]] = None, weight_scale: Optional[float] = None):
        """cols0: order in which the kernel's concatenated sources present the first layer's input columns"""
        assert 1 <= len(layers) <= 4 and len(acts) == len(layers)
        self.nlayer = len(layers)
        Kp = [_pad16(k_in)]
        imgs, sws, Nps, Ns = [], [], [], []
        for l, lin in enumerate(layers):
            assert (lin.weight.shape[1] == k_in or (l == 0 and cols0 is not None)) if l == 0 else lin.weight.shape[1] == Ns[-1]
            W = lin.weight
            if l == 0 and cols0 is not None:
                # cols0[k'] = reference input column presented at kernel column k' (-1 = padding column with zero weight)
                # a permutation of the reference columns, or (partial chains: layer-0 addend) a subset of them
                real = [c for c in cols0 if c >= 0]
                assert len(set(real)) == len(real) and max(real) < W.shape[1] and len(cols0) == k_in
                idx, mask = _cols_index(cols0, W.device)
                W = W.detach().index_select(1, idx) * mask
            img, sw, Np = pack_chain_layer(W, Kp[l], weight_scale)
            imgs.append(img); sws.append(sw); Nps.append(Np); Ns.append(lin.weight.shape[0])
            Kp.append(Np)
        self.Kp, self.N, self.Np, self.acts = Kp[:-1], Ns, Nps, list(acts)
        offs, o = [], 0
        for img in imgs:
            offs.append(o)
            o += img.numel()
        self.w_off = offs
        self.wpack = torch.cat(imgs).contiguous()
        s_in = [act_scale] * (self.nlayer + 1)
        s_in[self.nlayer] = 1.0                                     # the last layer's output is not re-scaled
        self.in_scale = s_in[0]
        self.mul = [s_in[l + 1] / (s_in[l] * sws[l]) for l in range(self.nlayer)]
        self.inv_next = [1.0 / s_in[l + 1] for l in range(self.nlayer)]
        dev = layers[0].weight.device
        bias = torch.zeros((4, NMAX), device=dev, dtype=torch.float32)
        for l, lin in enumerate(layers):
            if lin.bias is not None:
                bias[l, :Ns[l]] = lin.bias.detach().float() * s_in[l + 1]
        self.bias = bias.contiguous()


TRAIN_WEIGHT_SCALE = 1024.0


def packed_chain(owner, name: str, layers, acts, k_in: int, cols0=None, weight_scale=None) -> PackedChain:
    """cache on `owner` (a module) keyed by the parameters' versions"""
    key = tuple((p.data_ptr(), p._version) for lin in layers for p in (lin.weight, lin.bias) if p is not None)
    cache = owner.__dict__.setdefault("_chain_cache", {})
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        ent = (key, PackedChain(layers, acts, k_in, cols0=cols0, weight_scale=weight_scale))
        cache[name] = ent
    return ent[1]


def chain_forward(pc: PackedChain, srcs: Sequence[torch.Tensor], M: Optional[int] = None, mods: Sequence[int] = (),
                  out: bool = True, res: Optional[torch.Tensor] = None, head=None, keep_inner: bool = False, save_images: bool = False,
                  add0=None):
    """run the chain over M rows.  srcs: 2-D fp32 tensors with unit inner stride (row stride free); concat widths must sum to
    the first layer's K.  head = (weight (1,N) , bias (1,), act) fuses a 1-output layer on the last output.
    add0 = (A (rows, >= Np[0]) fp32 with 16-byte aligned rows, mod): A[m % mod] is added to layer 0's pre-activation.
    Returns (Y_last or None, head_out or None, [inner Y_l] if keep_inner)."""
    srcs = [ops._rows2d(s) for s in srcs]
    assert 1 <= len(srcs) <= 3
    if M is None:
        M = srcs[0].shape[0]
    dev = srcs[0].device
    ks = [s.shape[1] for s in srcs] + [0] * (3 - len(srcs))
    padded = list(srcs) + [None] * (3 - len(srcs))
    lds = [s.stride(0) if s is not None else 0 for s in padded]
    modl = list(mods) + [0] * (3 - len(mods))
    assert _pad16(sum(ks)) == pc.Kp[0], (ks, pc.Kp)
    nl = pc.nlayer
    Ys: List[Optional[torch.Tensor]] = [None] * nl
    if keep_inner:
        for l in range(nl - 1):
            Ys[l] = torch.empty((M, pc.N[l]), device=dev, dtype=torch.float32)
    if out:
        Ys[nl - 1] = torch.empty((M, pc.N[nl - 1]), device=dev, dtype=torch.float32)
    head_out = hw = hb = None
    hact = 0
    if head is not None:
        hw, hb, hact = ops._f32c(head[0]).view(-1), ops._f32c(head[1]).view(-1), int(head[2])
        assert hw.numel() == pc.N[nl - 1]
        hw = torch.nn.functional.pad(hw.detach(), (0, NMAX - hw.numel())).contiguous()      # the kernel reads whole 32-column blocks
        head_out = torch.empty((M, 1), device=dev, dtype=torch.float32)
    resv = ops._rows2d(res) if res is not None else None
    f4 = lambda v: (C.c_float * len(v))(*v)
    i4 = lambda v: (C.c_int * len(v))(*[int(x) for x in v])
    if add0 is not None:
        A0, amod = add0
        assert A0.dtype == torch.float32 and A0.stride(1) == 1 and A0.shape[1] >= pc.Np[0] and not keep_inner
        x0img = himg = None
        if save_images:
            from . import mlp_tc
            x0img = mlp_tc.image_empty(M, pc.Kp[0], dev)
            himg = [mlp_tc.image_empty(M, pc.Np[l], dev) for l in range(nl - 1)]
        with ops._launch():
            check(lib().hnr_chain_f16_forward_add0(ptr_array(padded), i64_array(lds), i64_array(ks), i64_array(modl), float(pc.in_scale), nl,
                                                   i64_array(pc.Kp), i64_array(pc.N), i64_array(pc.Np), i4(pc.acts), ptr(pc.wpack),
                                                   i64_array(pc.w_off), ptr(pc.bias), f4(pc.mul), f4(pc.inv_next), ptr_array(Ys),
                                                   i64_array([y.stride(0) if y is not None else 0 for y in Ys]), ptr(resv),
                                                   resv.stride(0) if resv is not None else 0, ptr(hw), ptr(hb), hact, ptr(head_out), M,
                                                   ptr(ops.status_word(dev)), ptr(x0img),
                                                   ptr_array(himg + [None] * (4 - len(himg))) if himg is not None else None, ptr(A0), A0.stride(0),
                                                   int(amod), float(ACT_SCALE), stream()), "chain_f16_forward_add0")
        return Ys[nl - 1], head_out, ((x0img, himg) if save_images else Ys[:-1])
    if save_images:
        # training forward: the concatenated input and the inner outputs go to HBM as split images (csrc/img_common.cuh)
        from . import mlp_tc
        x0img = mlp_tc.image_empty(M, pc.Kp[0], dev)
        himg = [mlp_tc.image_empty(M, pc.Np[l], dev) for l in range(nl - 1)]
        with ops._launch():
            check(lib().hnr_chain_f16_forward_train(ptr_array(padded), i64_array(lds), i64_array(ks), i64_array(modl), float(pc.in_scale), nl,
                                                    i64_array(pc.Kp), i64_array(pc.N), i64_array(pc.Np), i4(pc.acts), ptr(pc.wpack),
                                                    i64_array(pc.w_off), ptr(pc.bias), f4(pc.mul), f4(pc.inv_next), ptr_array(Ys),
                                                    i64_array([y.stride(0) if y is not None else 0 for y in Ys]), ptr(resv),
                                                    resv.stride(0) if resv is not None else 0, ptr(hw), ptr(hb), hact, ptr(head_out), M,
                                                    ptr(ops.status_word(dev)), ptr(x0img), ptr_array(himg + [None] * (4 - len(himg))), stream()),
                  "chain_f16_forward_train")
        return Ys[nl - 1], head_out, (x0img, himg)
    with ops._launch():
        check(lib().hnr_chain_f16_forward(ptr_array(padded), i64_array(lds), i64_array(ks), i64_array(modl), float(pc.in_scale), nl,
                                          i64_array(pc.Kp), i64_array(pc.N), i64_array(pc.Np), i4(pc.acts), ptr(pc.wpack),
                                          i64_array(pc.w_off), ptr(pc.bias), f4(pc.mul), f4(pc.inv_next), ptr_array(Ys),
                                          i64_array([y.stride(0) if y is not None else 0 for y in Ys]), ptr(resv),
                                          resv.stride(0) if resv is not None else 0, ptr(hw), ptr(hb), hact, ptr(head_out), M,
                                          ptr(ops.status_word(dev)), stream()),
              "chain_f16_forward")
    return Ys[nl - 1], head_out, Ys[:-1]


# ------------------------------------------------------------------------------------------------------------------
# fused backward (csrc/chain_bwd_f16.cu + csrc/wgrad_img.cu)
# ------------------------------------------------------------------------------------------------------------------
FUSED_BWD = __import__("os").environ.get("HNR_FUSED_BWD", "1") != "0"


def _pack_wT_chunks(B: torch.Tensor) -> torch.Tensor:
    """B (rows % 8 == 0, red % 16 == 0) fp32 = the B operand of a data-gradient MMA (rows = input columns k, reduction = output
    units n) -> red/16 chunk images [hi | lo], each part [k block (2)][row group][row (8)][8 bf16] (canonical K-major UMMA)"""
    rows, red = B.shape
    hi = B.bfloat16()
    lo = (B - hi.float()).bfloat16()
    tile = lambda x: x.view(rows // 8, 8, red // 16, 2, 8).permute(2, 3, 0, 1, 4)
    return torch.stack([tile(hi), tile(lo)], dim=1).contiguous().view(torch.uint8).reshape(-1)


class PackedChainBwd:
    """W^T images of a chain for hnr_chain_bwd_f16: layer l's operand has rows = Np[l-1] (l > 0) or NX (l = 0, the first NX
    kernel-order input columns) and the reduction over Np[l] output units."""

    def __init__(self, layers, pc: PackedChain, NX: int, cols0=None):
        dev = layers[0].weight.device
        self.NX = NX
        imgs, offs, o = [], [], 0
        for l, lin in enumerate(layers):
            W = lin.weight.detach().float()                        # (N_l, K_l)
            if l == 0 and cols0 is not None:
                idx, mask = _cols_index(cols0, dev)
                W = W.index_select(1, idx) * mask                    # kernel source order
            rows = NX if l == 0 else pc.Np[l - 1]
            B = torch.zeros((rows, pc.Np[l]), device=dev, dtype=torch.float32)
            kk = min(rows, W.shape[1])
            B[:kk, :W.shape[0]] = W.t()[:kk]
            img = _pack_wT_chunks(B)
            imgs.append(img)
            offs.append(o)
            o += img.numel()
        self.w_off = offs
        self.wpack = torch.cat(imgs).contiguous()


def packed_chain_bwd(owner, name: str, layers, pc: PackedChain, NX: int, cols0=None) -> PackedChainBwd:
    key = (NX,) + tuple((lin.weight.data_ptr(), lin.weight._version) for lin in layers)
    cache = owner.__dict__.setdefault("_chain_bwd_cache", {})
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        ent = (key, PackedChainBwd(layers, pc, NX, cols0))
        cache[name] = ent
    return ent[1]


WG_LDO = 320


def chain_backward_fused(pc: PackedChain, pb: PackedChainBwd, Ws, images, y_top: torch.Tensor, dY: torch.Tensor, M: int, acts, ks, mods,
                         need_src, cols0, params=None, add0_mod: int = 0):
    """data gradients (one fused launch) + weight / bias gradients (one image-fed launch) of a chain.
    add0_mod > 0: also returns the gradient of the layer-0 addend (sum over the M / add0_mod views of dZ_0).
    Returns ([d_src_i | None], [dW_l], [db_l][, d_add0])."""
    from . import mlp_tc
    x0img, himg = images
    nl = pc.nlayer
    dev = dY.device
    dY, y_top = ops._rows2d(dY), ops._rows2d(y_top)
    NX = pb.NX
    ktot = sum(ks)
    ldx = max(NX, (ktot + 3) // 4 * 4)
    dX = torch.empty((M, ldx), device=dev, dtype=torch.float32)
    if ldx > NX:
        dX[:, NX:].zero_()
    dz = [mlp_tc.image_empty(M, pc.Np[l], dev) for l in range(nl)]
    act_top = int(acts[nl - 1])
    with ops._launch(name="chain_bwd"):
        check(lib().hnr_chain_bwd_f16(nl, i64_array(pc.Np), i64_array(pc.N), NX, act_top, ptr(dY), dY.stride(0), ptr(y_top), y_top.stride(0),
                                      ptr_array(himg + [None] * (4 - len(himg))), ptr_array(dz + [None] * (4 - nl)), ptr(pb.wpack),
                                      i64_array(pb.w_off), ptr(dX), ldx, M, stream()), "chain_bwd_f16")
    # ONE zero fill for the chain's parameter-shaped gradients; the weight-gradient kernel accumulates straight into them (layer 0 through
    # the column map of the kernel's source order)
    sizes = []
    for l in range(nl):
        sizes += [Ws[l].numel(), Ws[l].shape[0]]
    pad4 = lambda n_: (n_ + 3) // 4 * 4                           # 16-byte aligned views
    flat = torch.zeros(sum(pad4(n_) for n_ in sizes), device=dev, dtype=torch.float32)
    dWs, dbs, o = [], [], 0
    for l in range(nl):
        dWs.append(flat[o:o + sizes[2 * l]].view_as(Ws[l])); o += pad4(sizes[2 * l])
        dbs.append(flat[o:o + sizes[2 * l + 1]]); o += pad4(sizes[2 * l + 1])
    cmap = _cols_map32(cols0, pc.Kp[0], dev) if cols0 is not None else None
    rp = mlp_tc.rows_padded(M)
    aW, ab = [ops.alias(t) for t in dWs], [ops.alias(t) for t in dbs]       # a parked launch refers to aliases only (ops.alias)
    shapes = ([w.shape[0] for w in Ws], [w.shape[1] for w in Ws])

    def launch_wgrad():
        with ops._launch(name="wgrad_img"):
            check(lib().hnr_wgrad_img_jobs(nl, ptr_array(dz), i64_array(pc.Np), ptr_array([x0img] + himg), None, i64_array([pc.Kp[0]] + pc.Np[:-1]),
                                           i64_array([0] * nl), ptr_array(aW), ptr_array(ab), ptr_array([cmap] + [None] * (nl - 1)),
                                           i64_array(shapes[0]), i64_array(shapes[1]), i64_array([rp] * nl), stream()),
                  "wgrad_img_jobs")
    pairs = []
    if params is not None:              # (weight, bias) parameters of the layers, in order: lets a training loop defer this launch
        for l in range(nl):
            pairs += [(params[2 * l], aW[l])] + ([(params[2 * l + 1], ab[l])] if params[2 * l + 1] is not None else [])
        ops._wgrad_launch(launch_wgrad, pairs)
    else:
        launch_wgrad()
    d_srcs, off = [], 0
    for i, k in enumerate(ks):
        g = None
        if need_src[i]:
            g = dX[:, off:off + k]
            if mods[i] > 0:
                g = g.reshape(-1, mods[i], k).sum(dim=0)
        d_srcs.append(g)
        off += k
    if add0_mod > 0:
        assert M % add0_mod == 0
        d_add0 = torch.empty((add0_mod, pc.Np[0]), device=dev, dtype=torch.float32)
        with ops._launch(name="img_sum_views"):
            check(lib().hnr_img_sum_views(ptr(dz[0]), pc.Np[0], add0_mod, M // add0_mod, ptr(d_add0), pc.Np[0], stream()), "img_sum_views")
        return d_srcs, dWs, dbs, d_add0
    return d_srcs, dWs, dbs


class ChainFn(torch.autograd.Function):
    """Graph-recording forward of a fused chain: one chain_f16 launch with every layer's output kept, backward layer by
    layer on the tensor-core gradient kernels (ops.linear_backward).  Same arithmetic and gradients as the equivalent
    sequence of ops.linear calls.  apply(pc, layers_params..., ) is wrapped by chain_train()."""

    @staticmethod
    def forward(ctx, pc, acts, mods, M, has_res, head_act, nlayer, nsrc, cols0, pb, add0_mod, *tensors):
        # tensors = [W_0, b_0, ..., W_{n-1}, b_{n-1}] + ([head_W, head_b] if head_act >= 0) + srcs + ([res] if has_res) + ([add0] if add0_mod)
        k = 2 * nlayer
        Ws, bs = list(tensors[0:k:2]), list(tensors[1:k:2])
        head = None
        if head_act >= 0:
            head = (tensors[k], tensors[k + 1], head_act)
            k += 2
        srcs = list(tensors[k:k + nsrc])
        res = tensors[k + nsrc] if has_res else None
        add0 = (tensors[k + nsrc + (1 if has_res else 0)], add0_mod) if add0_mod else None
        # gradients of outputs nobody differentiates (the last layer's output when only the head is used) arrive as None instead of a
        # materialised zero tensor (a 154 MB fill per step for the blend-weight net)
        ctx.set_materialize_grads(False)
        fused = FUSED_BWD and pb is not None
        ctx.fused = fused
        ctx.add0_mod = add0_mod
        assert add0 is None or fused, "the layer-0 addend needs the fused backward"
        if fused:
            y, h, images = chain_forward(pc, srcs, M=M, mods=mods, out=True, res=res, head=head, save_images=True, add0=add0)
            ctx.cfg = (acts, mods, M, has_res, head_act, nlayer, nsrc, cols0)
            ctx.pc, ctx.pb, ctx.ks = pc, pb, [s_.shape[1] for s_ in srcs]
            ctx.wparams = list(tensors[0:2 * nlayer])
            ctx.nimg = len(images[1])
            ctx.save_for_backward(*Ws, *([head[0]] if head else []), images[0], *images[1], y, *([h] if head else []))
            if head is not None:
                ctx.mark_non_differentiable(y)
                return y, h
            return y, y.new_empty(0)
        y, h, inner = chain_forward(pc, srcs, M=M, mods=mods, out=True, res=res, head=head, keep_inner=True)
        ctx.cfg = (acts, mods, M, has_res, head_act, nlayer, nsrc, cols0)
        ctx.save_for_backward(*Ws, *([head[0]] if head else []), *srcs, *inner, y, *([h] if head else []), *([res] if has_res else []))
        if head is not None:
            ctx.mark_non_differentiable(y)
            return y, h
        return y, y.new_empty(0)

    @staticmethod
    def backward(ctx, dY, dH):
        acts, mods, M, has_res, head_act, nlayer, nsrc, cols0 = ctx.cfg
        sv = list(ctx.saved_tensors)
        if head_act < 0 and dY is None:                      # the chain's output was not used at all
            dY = torch.zeros((M, ctx.pc.N[-1] if hasattr(ctx, "pc") else sv[nlayer - 1].shape[0]), device=sv[0].device, dtype=torch.float32)
        if head_act >= 0 and dH is None:
            dH = torch.zeros((M, 1), device=sv[0].device, dtype=torch.float32)
        if ctx.fused:
            return ChainFn._backward_fused(ctx, sv, dY, dH)
        Ws = sv[:nlayer]; p = nlayer
        head_W = None
        if head_act >= 0:
            head_W = sv[p]; p += 1
        srcs = sv[p:p + nsrc]; p += nsrc
        inner = sv[p:p + nlayer - 1]; p += nlayer - 1
        y = sv[p]; p += 1
        h = sv[p] if head_act >= 0 else None
        p += 1 if head_act >= 0 else 0
        res = sv[p] if has_res else None
        Ys = inner + [y]
        g_head = [None, None]
        if head_act >= 0:
            # head: h = act(y_last . w + b); its input is y_last AFTER the residual (none in that configuration)
            (dY_from_head,), dWh, dbh = ops.linear_backward(head_W, h, [y], (), dH, head_act, [True])
            g_head = [dWh, dbh]
            dcur = dY_from_head
        else:
            dcur = dY
        d_res = dcur if has_res else None
        gW = [None] * nlayer
        gb = [None] * nlayer
        d_srcs = [None] * nsrc
        for l in reversed(range(nlayer)):
            ins = srcs if l == 0 else [Ys[l - 1]]
            need = [ctx.needs_input_grad[11 + 2 * nlayer + (2 if head_act >= 0 else 0) + i] for i in range(nsrc)] if l == 0 else [True]
            Wl = Ws[l]
            if l == 0 and cols0 is not None:         # the kernel's source order is a column permutation of the reference weight
                idx = _cols_index(cols0, Wl.device)[0]
                Wl = Wl.index_select(1, idx)
            d_in, gW[l], gb[l] = ops.linear_backward(Wl, Ys[l], ins, mods if l == 0 else (), dcur, acts[l], need, M=M)
            if l == 0 and cols0 is not None and gW[l] is not None:
                gW[l] = torch.zeros_like(gW[l]).index_copy_(1, idx, gW[l])
            if l == 0:
                d_srcs = d_in
            else:
                dcur = d_in[0]
        grads = []
        for l in range(nlayer):
            grads += [gW[l], gb[l]]
        if head_act >= 0:
            grads += g_head
        grads += list(d_srcs)
        if has_res:
            grads.append(d_res)
        return (None,) * 11 + tuple(grads)

    @staticmethod
    def _backward_fused(ctx, sv, dY, dH):
        acts, mods, M, has_res, head_act, nlayer, nsrc, cols0 = ctx.cfg
        Ws = sv[:nlayer]; p = nlayer
        head_W = None
        if head_act >= 0:
            head_W = sv[p]; p += 1
        x0img = sv[p]; p += 1
        himg = list(sv[p:p + ctx.nimg]); p += ctx.nimg
        y = sv[p]; p += 1
        h = sv[p] if head_act >= 0 else None
        g_head = [None, None]
        if head_act >= 0:
            dcur, dWh, dbh = ops.linear_head_backward(head_W, h, y, dH, head_act)
            g_head = [dWh, dbh]
        else:
            dcur = dY
        d_res = dcur if has_res else None
        need = [ctx.needs_input_grad[11 + 2 * nlayer + (2 if head_act >= 0 else 0) + i] for i in range(nsrc)]
        modl = list(mods) + [0] * (nsrc - len(mods))
        r = chain_backward_fused(ctx.pc, ctx.pb, Ws, (x0img, himg), y, dcur, M, acts, ctx.ks, modl, need, cols0, params=ctx.wparams,
                                 add0_mod=ctx.add0_mod)
        d_srcs, gW, gb = r[0], r[1], r[2]
        grads = []
        for l in range(nlayer):
            grads += [gW[l], gb[l] if ctx.wparams[2 * l + 1] is not None else None]
        if head_act >= 0:
            grads += g_head
        grads += list(d_srcs)
        if has_res:
            grads.append(d_res)
        if ctx.add0_mod:
            grads.append(r[3])
        return (None,) * 11 + tuple(grads)


def chain_train(pc: PackedChain, layers, acts, srcs, M=None, mods=(), res=None, head=None, cols0=None, pb: Optional[PackedChainBwd] = None,
                add0=None):
    """autograd-aware fused chain.  layers: the nn.Linear modules (their parameters receive gradients); head = (nn.Linear, act) or
    None.  Returns (y_last, head_out | None).  NOTE: with a residual the last activation must be 'none' (as in the mix-up block):
    the saved output then includes the residual, which the identity derivative never reads."""
    srcs = [ops._rows2d(s) for s in srcs]
    if M is None:
        M = srcs[0].shape[0]
    assert res is None or acts[-1] == ops.ACT_NONE
    tensors = []
    for lin in layers:
        tensors += [lin.weight, lin.bias]
    head_act = -1
    if head is not None:
        tensors += [head[0].weight, head[0].bias]
        head_act = int(head[1])
    tensors += srcs
    if res is not None:
        tensors.append(res)
    add0_mod = 0
    if add0 is not None:                     # (A (rows, Np[0]) fp32, mod): added to layer 0's pre-activation, receives a gradient
        tensors.append(add0[0])
        add0_mod = int(add0[1])
    y, h = ChainFn.apply(pc, tuple(acts), tuple(mods), M, res is not None, head_act, len(layers), len(srcs),
                         tuple(cols0) if cols0 is not None else None, pb, add0_mod, *tensors)
    return y, (h if head is not None else None)
