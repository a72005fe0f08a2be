"""Host side of the fused tensor-core per-neighbour MLP (csrc/mlp_tc.cu): weight packing + launch.

Packing: every layer's (256, K) nn.Linear weight is (1) column-permuted / zero-padded to the K
order the kernel generates its operand chunks in, (2) split into TF32 `hi` (low 13 mantissa bits
cleared) and `lo = w - hi`, and (3) tiled into the canonical no-swizzle K-major UMMA layout, one
16 KB image [hi 8 KB | lo 8 KB] per 8-column chunk, so that the kernel fetches a chunk with a single
cp.async.bulk.  Layout of one 8 KB part: [k half (2)][row group (32)][row (8)][4 floats].
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import check, lib, ptr, stream

KC = 8
KP = (288, 256, 264, 256)


def layer1_column_order() -> List[int]:
    """reference column (in the 284-wide block1 input) of every kernel-order column k' (-1 = zero pad).
    reference order (point_aggregators.py:931-939, networks.py:175-189): [emb 32 | (c*3+f)*2+{sin,cos} 192 |
    (j*5+f)*2+{sin,cos} 60].  kernel order: [emb 32 | per 8-channel block: f, sin|cos, 8 channels |
    dist: f*12 + (sin|cos)*6 + j | 4 x pad]."""
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


def _permute_pad(W: torch.Tensor, cols: Sequence[int]) -> torch.Tensor:
    idx = torch.tensor([c if c >= 0 else 0 for c in cols], device=W.device, dtype=torch.long)
    out = W.index_select(1, idx)
    pad = torch.tensor([c < 0 for c in cols], device=W.device)
    return out.masked_fill(pad[None, :], 0.0)


def pack_layer(W: torch.Tensor, cols: Optional[Sequence[int]] = None) -> torch.Tensor:
    """W (256, K) fp32 -> uint8 image of ceil(K/8) chunks (see module docstring)."""
    W = W.detach().float()
    assert W.shape[0] == 256
    if cols is None:
        Kp = (W.shape[1] + KC - 1) // KC * KC
        cols = list(range(W.shape[1])) + [-1] * (Kp - W.shape[1])
    Wp = _permute_pad(W, cols).contiguous()
    Kp = Wp.shape[1]
    hi = (Wp.view(torch.int32) & -8192).view(torch.float32)          # clear the 13 low mantissa bits
    lo = Wp - hi
    def tile(x):                                                     # (256,Kp) -> (C, 2, 32, 8, 4)
        return x.view(32, 8, Kp // KC, 2, 4).permute(2, 3, 0, 1, 4)
    img = torch.stack([tile(hi), tile(lo)], dim=1).contiguous()      # (C, 2[hi|lo], 2, 32, 8, 4)
    return img.view(torch.uint8).reshape(-1)


def pack_mlp(block1, block3) -> torch.Tensor:
    """packed image of the four dense layers of the per-neighbour MLP + (4,256) biases"""
    l3 = list(range(263)) + [-1]
    parts = [pack_layer(block1[0].weight, layer1_column_order()), pack_layer(block1[2].weight),
             pack_layer(block3[0].weight, l3), pack_layer(block3[2].weight)]
    for p, kp in zip(parts, KP):
        assert p.numel() == lib().hnr_mlp_tc_packed_bytes(kp)
    bias = torch.stack([block1[0].bias, block1[2].bias, block3[0].bias, block3[2].bias]).detach().float().contiguous()
    return torch.cat(parts).contiguous(), bias


def gemm_test(a: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """self test: a (rows,K) fp32, W (256,K) -> a @ W^T through the tcgen05 3xTF32 pipeline"""
    a = a.float().contiguous()
    img = pack_layer(W)
    Kp = (W.shape[1] + KC - 1) // KC * KC
    out = torch.empty((a.shape[0], 256), device=a.device, dtype=torch.float32)
    with ops._launch():
        check(lib().hnr_mlp_tc_gemm_test(ptr(a), a.shape[0], a.shape[1], ptr(img), Kp, ptr(out), stream()), "mlp_tc_gemm_test")
    return out


def forward(tables, pidx, vlist, loc_w, loc_pers, raydirs, cam, weight, confc, wpack, bias, w_alpha, b_alpha):
    """fused gather + per-neighbour MLP + density head + weighted K-sum (inference).  Returns sigma (Nv,1), X5 (Nv,280)."""
    xyz, xyz_pers, emb, color, dirs, _ = tables
    Nv, K = vlist.shape[0], pidx.shape[1]
    sigma = torch.empty((Nv, 1), device=pidx.device, dtype=torch.float32)
    X5 = torch.empty((Nv, 280), device=pidx.device, dtype=torch.float32)
    wa = w_alpha.detach().float().contiguous().view(-1)
    ba = b_alpha.detach().float().contiguous().view(-1)
    with ops._launch():
        check(lib().hnr_mlp_tc_forward(ptr(xyz), ptr(xyz_pers), ptr(emb), ptr(color), ptr(dirs), ptr(pidx), ptr(vlist), ptr(loc_w),
                                       ptr(loc_pers), ptr(raydirs), ptr(cam), ptr(weight), ptr(confc), ptr(wpack), ptr(bias), ptr(wa),
                                       ptr(ba), Nv, K, ptr(sigma), ptr(X5), stream()), "mlp_tc_forward")
    return sigma, X5
