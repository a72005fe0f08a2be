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


def pack_chain_layer(W: torch.Tensor, Kp: int):
    W = W.detach().float()
    N, K = W.shape
    Np = _pad16(N)
    assert Np <= NMAX and K <= Kp and Kp % 16 == 0
    Wp = torch.zeros((Np, Kp), device=W.device, dtype=torch.float32)
    Wp[:N, :K] = W
    wmax = float(Wp.abs().max())
    sw = 2.0 ** math.floor(math.log2(16384.0 / wmax)) if wmax > 0 else 1.0
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
                 cols0: Optional[Sequence[int]] = None):
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
                assert sorted(c for c in cols0 if c >= 0) == list(range(W.shape[1])) and len(cols0) == k_in
                idx = torch.tensor([max(c, 0) for c in cols0], device=W.device)
                W = W.detach().index_select(1, idx) * torch.tensor([1.0 if c >= 0 else 0.0 for c in cols0], device=W.device)
            img, sw, Np = pack_chain_layer(W, Kp[l])
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


def packed_chain(owner, name: str, layers, acts, k_in: int, cols0=None) -> PackedChain:
    """cache on `owner` (a module) keyed by the parameters' versions"""
    key = tuple((p.data_ptr(), p._version) for lin in layers for p in (lin.weight, lin.bias) if p is not None)
    cache = owner.__dict__.setdefault("_chain_cache", {})
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        ent = (key, PackedChain(layers, acts, k_in, cols0=cols0))
        cache[name] = ent
    return ent[1]


def chain_forward(pc: PackedChain, srcs: Sequence[torch.Tensor], M: Optional[int] = None, mods: Sequence[int] = (),
                  out: bool = True, res: Optional[torch.Tensor] = None, head=None, keep_inner: bool = False):
    """run the chain over M rows.  srcs: 2-D fp32 tensors with unit inner stride (row stride free); concat widths must sum to
    the first layer's K.  head = (weight (1,N) , bias (1,), act) fuses a 1-output layer on the last output.
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
    with ops._launch():
        check(lib().hnr_chain_f16_forward(ptr_array(padded), i64_array(lds), i64_array(ks), i64_array(modl), float(pc.in_scale), nl,
                                          i64_array(pc.Kp), i64_array(pc.N), i64_array(pc.Np), i4(pc.acts), ptr(pc.wpack),
                                          i64_array(pc.w_off), ptr(pc.bias), f4(pc.mul), f4(pc.inv_next), ptr_array(Ys),
                                          i64_array([y.stride(0) if y is not None else 0 for y in Ys]), ptr(resv),
                                          resv.stride(0) if resv is not None else 0, ptr(hw), ptr(hb), hact, ptr(head_out), M,
                                          ptr(ops.status_word(dev)), stream()),
              "chain_f16_forward")
    return Ys[nl - 1], head_out, Ys[:-1]
