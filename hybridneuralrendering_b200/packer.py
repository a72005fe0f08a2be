"""One-launch weight re-packing for the training step (csrc/pack.cu).

Every optimiser step changes all MLP weights, and every fused tensor-core kernel reads them as split hi/lo chunk images
(mlp_tc.py / chain.py describe the layouts).  Re-packing with tensor ops costs ~270 tiny launches per step; the `TrainPacker`
builds the images once with those (reference) packers, records where every layer's image lives in a static device-resident job
table, and from then on refreshes all of them -- forward fp16 images, transposed bf16 images of the data-gradient chains, the
pre-scaled bias tables -- with ONE kernel launch whenever a parameter version changed."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import chain, mlp_tc, ops
from ._lib import check, lib, ptr, stream
from .ops import ACT_LRELU, ACT_NONE

vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class PackJob(C.Structure):         # mirrors csrc/pack.cu
    _fields_ = [("W", vp), ("colmap", vp), ("dst", vp), ("piece0", i64), ("scale", f32), ("ldw", i32), ("n_src", i32), ("k_src", i32),
                ("rows_p", i32), ("red_p", i32), ("transpose", i32), ("fmt", i32), ("map_len", i32), ("pad_", i32)]


class BiasJob(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("scale", f32), ("n", i32)]


AM_COLS0 = list(range(45, 173)) + list(range(45)) + [173, 174, 175]     # blend-weight net, kernel source order [g | aux | dview]
# the blend-weight net's first layer split in two (chain.py add0): the 128 columns that multiply g (shared by the V views of a sample)
# run as a one-layer chain per SAMPLE, the 48 view-dependent columns [aux 45 | dview 3] per (view, sample)
AMG_COLS = list(range(45, 173))
AM_COLS48 = list(range(45)) + [173, 174, 175]


class _NoBias:
    """a layer that shares another layer's weight tensor and has no bias (the g part of the blend-weight net's first layer)"""

    def __init__(self, lin):
        self.weight, self.bias = lin.weight, None


class TrainPacker:
    """packs of a graph-recording forward of `agg` (a PointAggregator) + the job table that refreshes them in one launch"""

    def __init__(self, agg, with_bwd: bool):
        assert C.sizeof(PackJob) == lib().hnr_pack_job_bytes() and C.sizeof(BiasJob) == lib().hnr_bias_job_bytes()
        TS = mlp_tc.TRAIN_WEIGHT_SCALE
        b1, b3 = agg.block1, agg.block3
        cf, am, cm = agg.color_feature_branch, agg.aux_merge_weight_block, agg.color_mixup_block
        self.dev = b1[0].weight.device
        self.with_bwd = with_bwd
        self.params = self._params(agg)
        # ---- build everything once with the tensor-op packers (they define the layouts)
        self.nbr_pack = mlp_tc.pack_mlp_f16(b1, b3, weight_scale=TS)
        self.nbr_packT = mlp_tc.pack_mlp_bwd(b1, b3) if with_bwd else None
        self.pc: Dict[str, chain.PackedChain] = {}
        self.pb: Dict[str, Optional[chain.PackedChainBwd]] = {}
        specs = [("cf", [cf[0], cf[2], cf[4]], [ACT_LRELU] * 3, ops.X5_W, None, 256)]
        if int(agg.opt.use_nearest) > 0 and with_bwd:
            specs.append(("amg", [_NoBias(am[0])], [ACT_NONE], 128, AMG_COLS, 128))
            specs.append(("am", [am[0], am[2], am[4]], [ACT_LRELU] * 3, 48, AM_COLS48, 48))
        elif int(agg.opt.use_nearest) > 0:
            # layer-by-layer backward (the tests' cross-check of the fused path): the unsplit 176-wide first layer, sources [g | aux | dview]
            specs.append(("am", [am[0], am[2], am[4]], [ACT_LRELU] * 3, 176, AM_COLS0, 176))
        specs.append(("cm", [cm[0], cm[2], cm[4]], [ACT_LRELU, ACT_LRELU, ACT_NONE], 90, None, 96))
        for name, layers, acts, k_in, cols0, nx in specs:
            self.pc[name] = chain.PackedChain(layers, acts, k_in, cols0=cols0, weight_scale=TS)
            self.pb[name] = chain.PackedChainBwd(layers, self.pc[name], nx, cols0=cols0) if with_bwd else None
        # ---- job table
        self._keep: List[torch.Tensor] = []           # column maps
        jobs: List[PackJob] = []
        bias: List[BiasJob] = []
        self._pieces = 0

        def cmap(cols):
            if cols is None:
                return None, 0
            t = torch.tensor(list(cols), dtype=torch.int32, device=self.dev)
            self._keep.append(t)
            return t, len(cols)

        def add(W, dst_t, dst_off, rows_p, red_p, transpose, fmt, scale, cols=None, k_src=None):
            cm_t, ml = cmap(cols)
            j = PackJob()
            j.W, j.colmap, j.dst = W.data_ptr(), (cm_t.data_ptr() if cm_t is not None else None), dst_t.data_ptr() + dst_off
            j.piece0, j.scale, j.ldw, j.n_src = self._pieces, float(scale), W.stride(0), W.shape[0]
            j.k_src = W.shape[1] if k_src is None else k_src
            j.rows_p, j.red_p, j.transpose, j.fmt, j.map_len = rows_p, red_p, transpose, fmt, ml
            assert W.stride(1) == 1 and rows_p % 8 == 0 and red_p % 16 == 0
            jobs.append(j)
            self._pieces += (red_p // 16) * 2 * rows_p
            return (red_p // 16) * rows_p * 64           # bytes of the image

        def add_bias(src, dst, scale):
            b = BiasJob()
            b.src, b.dst, b.scale, b.n = src.data_ptr(), dst.data_ptr(), float(scale), src.numel()
            bias.append(b)

        # per-neighbour MLP, forward (fp16 x TS; mlp_tc.pack_mlp_f16)
        wpack, nbias = self.nbr_pack[0], self.nbr_pack[1]
        l2 = list(range(256, 263)) + [-1] * 9 + list(range(256))
        off = 0
        for l, (lin, cols, kp) in enumerate([(b1[0], mlp_tc.layer1_column_order_f16(), 288), (b1[2], None, 256), (b3[0], l2, 272), (b3[2], None, 256)]):
            off += add(lin.weight, wpack, off, 256, kp, 0, 0, TS, cols)
            add_bias(lin.bias, nbias[l], mlp_tc.ACT_SCALE if l < 3 else 1.0)
        assert off == wpack.numel()
        if with_bwd:                                     # mlp_tc.pack_mlp_bwd
            off = 0
            for lin, ks in ((b3[2], None), (b3[0], 256), (b1[2], None), (b1[0], mlp_tc.X0_GRAD_W)):
                off += add(lin.weight, self.nbr_packT, off, 256, 256, 1, 1, 1.0, None, ks)
            assert off == self.nbr_packT.numel()
        for name, layers, acts, k_in, cols0, nx in specs:
            pc, pb = self.pc[name], self.pb[name]
            nl = pc.nlayer
            for l, lin in enumerate(layers):
                n = add(lin.weight, pc.wpack, pc.w_off[l], pc.Np[l], pc.Kp[l], 0, 0, TS, cols0 if l == 0 else None)
                assert pc.w_off[l] + n == (pc.w_off[l + 1] if l + 1 < nl else pc.wpack.numel())
                if lin.bias is not None:
                    add_bias(lin.bias, pc.bias[l], chain.ACT_SCALE if l < nl - 1 else 1.0)
            if pb is not None:
                for l, lin in enumerate(layers):
                    rows = nx if l == 0 else pc.Np[l - 1]
                    n = add(lin.weight, pb.wpack, pb.w_off[l], rows, pc.Np[l], 1, 1, 1.0, cols0 if l == 0 else None)
                    assert pb.w_off[l] + n == (pb.w_off[l + 1] if l + 1 < nl else pb.wpack.numel())
        self.njobs, self.nbias = len(jobs), len(bias)
        self._jobs = torch.frombuffer(bytearray(bytes((PackJob * len(jobs))(*jobs))), dtype=torch.uint8).to(self.dev)
        self._bias = torch.frombuffer(bytearray(bytes((BiasJob * len(bias))(*bias))), dtype=torch.uint8).to(self.dev)
        self.ptr_key = tuple(p.data_ptr() for p in self.params)
        self.ver_key = tuple(p._version for p in self.params)

    @staticmethod
    def _params(agg):
        mods = [agg.block1[0], agg.block1[2], agg.block3[0], agg.block3[2]]
        for seq in (agg.color_feature_branch, agg.aux_merge_weight_block, agg.color_mixup_block):
            mods += [seq[0], seq[2], seq[4]]
        return [p for m in mods for p in (m.weight, m.bias)]

    def refresh(self) -> None:
        """bring every image up to date with the current parameter values (one launch); no-op when nothing changed"""
        ver = tuple(p._version for p in self.params)
        if ver == self.ver_key:
            return
        with ops._launch(name="pack_weights"):
            check(lib().hnr_pack_weights(ptr(self._jobs), self.njobs, self._pieces, ptr(self._bias), self.nbias, ptr(ops.status_word(self.dev)),
                                         stream()), "pack_weights")
        self.ver_key = ver

    def current(self, agg, with_bwd: bool) -> bool:
        """False when the parameters were re-allocated (load_state_dict to new storage, .to(device), ...) or the mode changed"""
        return with_bwd == self.with_bwd and tuple(p.data_ptr() for p in self._params(agg)) == self.ptr_key
