"""GPU parity of the device frame-dict producer against the numpy oracle and against golden outputs of the
unmodified reference ``ScannetFtDataset.__getitem__`` (tests/golden/frame.npz).  Integer/byte work (pixel grid, chosen views,
uint8 -> fp32 frames, ground-truth lookup) must be bit-exact; ray directions: |err| <= 3e-7 (the reference's ``dirs @ rot.T``
goes through BLAS, whose 3-term summation order is not specified)."""
import random

import numpy as np
import pytest
import torch

from frame_cases import FRAME_CASES
from helpers import load_golden
from hybridneuralrendering_b200 import frame_producer as fp
from hybridneuralrendering_b200 import synthetic as syn
from oracle import frame_oracle as fo

pytestmark = pytest.mark.gpu
RAYDIR_ATOL = 3e-7


def _producer(split, over, bg, images, c2w, vids, K, train_ids, test_ids):
    bank = fp.FrameBank(images, c2w, vids, K, "cuda")
    opt = fp.default_opt(**over)
    return fp.FrameProducer(bank, train_ids if split == "train" else test_ids, train_ids, opt, split=split, near_far=(0.1, 8.0),
                            bg_color=bg, blur_kernels=np.zeros((1, 9, 9), np.float32), total_num_image=vids[-1] + 1)


def test_items_match_reference_golden_and_oracle():
    G = load_golden("frame")
    scene = syn.frame_scene()
    images, c2w, vids, K, train_ids, test_ids = scene
    for name, split, idx, over, seed, bg in FRAME_CASES:
        prod = _producer(split, over, bg, *scene)
        random.seed(seed)
        np.random.seed(seed)
        it = prod.item(idx)
        after = np.array([random.random(), np.random.rand()])
        random.seed(seed)
        np.random.seed(seed)
        ref = fo.frame_item(images, c2w, vids, K, train_ids if split == "train" else test_ids, train_ids, idx, split=split, bg_color=bg,
                            total_num_image=vids[-1] + 1, **over)
        np.testing.assert_array_equal(after, G[f"{name}_after"], err_msg=f"{name}: RNG streams not consumed like the reference")
        assert it["vid_nearest"].tolist() == G[f"{name}_vid_nearest"].tolist() == ref["vid_nearest"].tolist()
        assert [it["vid"], int(it["h"][0]), int(it["w"][0])] == G[f"{name}_meta"][:3].tolist()
        for k in ("pixel_idx", "gt_image", "c2w_nearest", "campos_nearest", "camrotc2w_nearest", "c2w", "campos", "camrotc2w", "bg_color"):
            got = it[k][0].cpu().numpy()
            np.testing.assert_array_equal(got, G[f"{name}_{k}"].astype(np.float32).reshape(got.shape), err_msg=f"{name}:{k} vs reference")
            np.testing.assert_array_equal(got, np.asarray(ref[k], np.float32).reshape(got.shape), err_msg=f"{name}:{k} vs oracle")
        np.testing.assert_array_equal(it["images_nearest"][0].cpu().numpy(), ref["images_nearest"])
        rd = it["raydir"][0].cpu().numpy()
        np.testing.assert_allclose(rd, G[f"{name}_raydir"], rtol=0, atol=RAYDIR_ATOL, err_msg=name)
        np.testing.assert_allclose(rd, ref["raydir"], rtol=0, atol=RAYDIR_ATOL, err_msg=name)
        np.testing.assert_allclose(it["vid_angle_nearest"], G[f"{name}_vid_angle_nearest"], rtol=1e-12)
        np.testing.assert_allclose(float(it["middle"]), float(G[f"{name}_middle"].reshape(())), rtol=1e-6)
        assert float(it["near"]) == float(G[f"{name}_near"].reshape(())) and float(it["far"]) == float(G[f"{name}_far"].reshape(()))
        assert it["blur_kernels"].shape == (1, 1, 9, 9)


def test_full_size_frame_properties():
    """BASELINE configs[2] frame size (640x480, V=8): size-independent properties instead of the (slow) python oracle loop."""
    rng = np.random.default_rng(3)
    F, H, W = 12, 480, 640
    images = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    vids = [5 * i for i in range(F)]
    c2w = np.stack([syn.look_at(np.array([2.0 + 0.05 * t, 2.0, 1.4]), np.array([5.4, 2.75 + 0.2 * t, 1.0])) for t in range(F)]).astype(np.float32)
    K = syn.intrinsic_matrix(577.87, 577.87, 320.0, 240.0).astype(np.float32)
    bank = fp.FrameBank(images, c2w, vids, K, "cuda")
    # (1) test item: full frame, every pixel once, gt == frame / 255 exactly, views == bank rows / 255 exactly
    prod = fp.FrameProducer(bank, vids, vids, fp.default_opt(use_nearest=8, random_sample="no_crop", find_nearest_mode=0), split="test")
    it = prod.item(6)
    assert it["raydir"].shape == (1, H * W, 3) and it["pixel_idx"].shape == (1, H, W, 2)
    # reference values by IEEE division on the host (torch's CUDA `x / 255.0` multiplies by a rounded reciprocal instead)
    u8_to_f32 = lambda a: torch.from_numpy(a.astype(np.float32) / np.float32(255)).cuda()
    assert torch.equal(it["gt_image"][0], u8_to_f32(images[6]).reshape(-1, 3))
    px, py = it["pixel_idx"][0, ..., 0], it["pixel_idx"][0, ..., 1]
    assert torch.equal(px, torch.arange(W, device="cuda", dtype=torch.float32)[None].expand(H, W))
    assert torch.equal(py, torch.arange(H, device="cuda", dtype=torch.float32)[:, None].expand(H, W))
    rows = [bank.row_of_vid[int(v)] for v in it["vid_nearest"]]
    assert 6 not in rows and len(set(rows)) == 8
    assert torch.equal(it["images_nearest"][0], u8_to_f32(images[rows]))
    # ray directions against the same formula in fp64
    x = (px.double() + 0.5 - float(K[0, 2])) / float(K[0, 0])
    y = (py.double() + 0.5 - float(K[1, 2])) / float(K[1, 1])
    d = torch.stack([x, y, torch.ones_like(x)], -1).reshape(-1, 3) @ torch.from_numpy(c2w[6, :3, :3]).double().cuda().T
    assert (it["raydir"][0].double() - d).abs().max().item() < 3e-7
    # (2) training item: 8x8 patches of 8x8, dilations 1..8: patch structure and bounds
    prod = fp.FrameProducer(bank, vids, vids, fp.default_opt(use_nearest=8, dilation_setup="8_8_1_8", edge_filter=10, dir_norm=1), split="train")
    random.seed(1)
    np.random.seed(1)
    it = prod.item(3)
    pix = it["pixel_idx"][0]
    assert pix.shape == (64, 64, 2) and it["raydir"].shape == (1, 4096, 3)
    assert pix[..., 0].min() >= 10 and pix[..., 0].max() < W - 10 and pix[..., 1].min() >= 10 and pix[..., 1].max() < H - 10
    blocks = pix.reshape(8, 8, 8, 8, 2).permute(0, 2, 1, 3, 4)                     # (pi, pj, a, b, 2)
    dx = blocks[:, :, :, 1:, 0] - blocks[:, :, :, :-1, 0]
    dy = blocks[:, :, 1:, :, 1] - blocks[:, :, :-1, :, 1]
    dil = dx[:, :, 0, 0]
    assert ((dil >= 1) & (dil <= 8)).all() and (dx == dil[:, :, None, None]).all() and (dy == dil[:, :, None, None]).all()
    assert torch.allclose(it["raydir"][0].norm(dim=-1), torch.ones(4096, device="cuda"), atol=2e-5)     # dir_norm: 1/(|d| + 1e-5)
    gt = u8_to_f32(images[3])[pix[..., 1].long(), pix[..., 0].long()]
    assert torch.equal(it["gt_image"][0], gt.reshape(-1, 3))


def test_producer_item_feeds_the_renderer():
    """the item goes straight into NeuralPointsRayMarching.forward: same output as the host-built dict of the same frame"""
    from hybridneuralrendering_b200 import make_opt, NeuralPoints, NeuralPointsRayMarching, PointAggregator
    from oracle import render_oracle as ro
    H, W, V = 48, 64, 3
    rng = np.random.default_rng(11)
    fr = syn.room_frame(H=H, W=W, V=V, patch_num=4, patch_size=4, seed=2)
    # a 1+V frame scene whose poses are room_frame's: frame 0 = the item, frames 5, 10, 15 = its reference views
    images = rng.integers(0, 256, (1 + V, H, W, 3), dtype=np.uint8)
    c2w = np.concatenate([fr["c2w"], fr["c2w_nearest"][0]]).astype(np.float32)
    vids = [0, 5, 10, 15]
    bank = fp.FrameBank(images, c2w, vids, fr["intrinsic"][0], "cuda")
    prod = fp.FrameProducer(bank, [0], vids, fp.default_opt(use_nearest=V, random_sample="no_crop", find_nearest_mode=0), split="test")
    it = prod.item(0)
    assert it["vid_nearest"].tolist() == [5, 10, 15]
    opt = make_opt("scannet", use_nearest=V, SR=24)
    xyz = syn.room_scene(40000, 5)
    att = syn.point_attributes(np.random.default_rng(5), len(xyz))
    dev = torch.device("cuda")
    c = lambda a: torch.from_numpy(a).cuda()
    pts = NeuralPoints(32, len(xyz), opt, dev)
    pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                   points_conf=c(att["conf"])[None], parameter=True)
    agg = PointAggregator(opt).cuda()
    agg.load_state_dict(ro.random_params(7), strict=False)
    net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).cuda()
    with torch.no_grad():
        a = net(**it)
        px, py = syn.full_frame_pixels(H, W)
        host = dict(campos=c(fr["campos"]), camrotc2w=c(fr["camrotc2w"]), c2w=c(fr["c2w"]), raydir=it["raydir"].clone(),
                    pixel_idx=c(np.stack([px, py], -1))[None], near=c(fr["near"]), far=c(fr["far"]), h=fr["h"], w=fr["w"], intrinsic=c(fr["intrinsic"]),
                    bg_color=c(fr["bg_color"]), images_nearest=c(images[1:].astype(np.float32) / np.float32(255))[None], c2w_nearest=c(fr["c2w_nearest"]),
                    campos_nearest=c(fr["campos_nearest"]), intrinsic_nearest=c(fr["intrinsic_nearest"]))
        b = net(**host)
    assert a["coarse_raycolor"].shape[1] > 100
    # the host dict reuses the producer's ray directions (their parity is the subject of the first test; numpy's BLAS summation
    # order differs between hosts), everything else is built independently on the host -> same rays, same colours
    assert (it["raydir"][0] - c(syn.rays_for_pixels(px, py, fr["intrinsic"][0], fr["c2w"][0]))).abs().max().item() < 1e-6
    same = (a["ray_mask"] == b["ray_mask"]).float().mean().item()
    assert same > 0.995, same
    if torch.equal(a["ray_mask"], b["ray_mask"]):
        diff = (a["coarse_raycolor"] - b["coarse_raycolor"]).abs().amax(dim=-1)[0]
        assert (diff > 1e-4).float().mean().item() < 0.01
