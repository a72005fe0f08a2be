"""CPU restatement (numpy) of the reference's frame-dict producer -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm may import this module; the product
(``hybridneuralrendering_b200.frame_producer`` + ``csrc/frame.cu``) never does.

Follows ``ScannetFtDataset.__getitem__`` (data/scannet_ft_dataset.py:736-976) and ``get_dtu_raydir``
(data/data_utils.py:57-71) of the reference from the decoded uint8 frames on.  Pinned by
``tests/golden/frame.npz``: outputs of the UNMODIFIED reference ``__getitem__`` run on a temporary scene
(``tests/golden/make_golden.py::frame_case``), see ``tests/test_oracle_golden.py::test_frame_producer_matches_reference``.
"""
import math
import random

import numpy as np


def nearest_views(train_id_list, vid, use_nearest, find_nearest_mode=0, split="train", weights=None, select_high_quality=0):
    """data/scannet_ft_dataset.py:771-812"""
    if use_nearest <= 0:
        return np.array([0])
    ids = np.array(train_id_list)
    id_dist = np.abs(ids - vid)
    min_idx = np.argsort(id_dist)
    drop_self = id_dist[min_idx[0]] == 0 and (find_nearest_mode == 0 or (find_nearest_mode == 1 and split == "train"))
    if find_nearest_mode not in (0, 1):
        raise NotImplementedError
    lo = 1 if drop_self else 0
    if select_high_quality > 0:
        num = int(use_nearest * 1.5)
        cand = ids[min_idx[lo:num + lo]]
        w = np.array(weights)[min_idx[lo:num + lo]]
        return cand[np.argsort(-w)[0:use_nearest]]
    return ids[min_idx[lo:use_nearest + lo]]


def dilated_pixels(width, height, margin, patch_num, patch_size, dilations):
    """data/scannet_ft_dataset.py:917-940; draws from the GLOBAL `random` / `np.random` streams like the reference"""
    S = patch_num * patch_size
    px, py = np.zeros((S, S)), np.zeros((S, S))
    for i in range(patch_num):
        for j in range(patch_num):
            d = int(random.choice(dilations))
            gx, gy = np.meshgrid(np.arange(patch_size).astype(np.float32), np.arange(patch_size).astype(np.float32))
            x0 = np.random.randint(margin, width - margin - (patch_size - 1) * d)
            y0 = np.random.randint(margin, height - margin - (patch_size - 1) * d)
            px[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size] = x0 + d * gx
            py[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size] = y0 + d * gy
    return px, py


def dtu_raydir(pixelcoords, intrinsic, rot, dir_norm):
    """data/data_utils.py:57-71 (rot = camera-to-world rotation); fp32 like the reference's inputs"""
    x = (pixelcoords[..., 0] + 0.5 - intrinsic[0, 2]) / intrinsic[0, 0]
    y = (pixelcoords[..., 1] + 0.5 - intrinsic[1, 2]) / intrinsic[1, 1]
    dirs = np.stack([x, y, np.ones_like(x)], axis=-1) @ rot.T
    if dir_norm:
        dirs = dirs / (np.linalg.norm(dirs, axis=-1, keepdims=True) + 1e-5)
    return dirs


def frame_item(images_u8, c2ws, vids, intrinsic, id_list, train_id_list, id, split="train", use_nearest=4, find_nearest_mode=0,
               dynamic_nearest=0, edge_filter=0, random_sample="dilated", random_sample_size=32, dilation_setup="8_8_1_8", dir_norm=0,
               near_far=(0.1, 8.0), total_num_image=None, bg_color=(1.0, 1.0, 1.0)):
    """one item; images_u8 (F,H,W,3) uint8, c2ws (F,4,4) fp32, vids (F,) frame numbers.  Same global-RNG consumption
    order as the reference: dynamic_nearest draw, sampler draws, bg_color draw."""
    row = {int(v): i for i, v in enumerate(vids)}
    vid = id_list[id]
    total = total_num_image if total_num_image is not None else max(vids) + 1
    if dynamic_nearest:
        use_nearest = np.random.randint(2, 8) if split == "train" else 4
    vn = nearest_views(train_id_list, vid, use_nearest, find_nearest_mode, split)
    imgs_n = np.stack([images_u8[row[int(v)]].astype(np.float32) / np.float32(255) for v in vn])
    c2w_n = np.stack([c2ws[row[int(v)]] for v in vn]).astype(np.float32)
    if use_nearest <= 0:
        imgs_n = imgs_n * 0
    c2w = c2ws[row[vid]].astype(np.float32)
    H, W = images_u8.shape[1:3]
    m = edge_filter
    out = dict(vid=vid, vid_nearest=np.asarray(vn), images_nearest=imgs_n, c2w_nearest=c2w_n, campos_nearest=c2w_n[:, :3, 3],
               camrotc2w_nearest=c2w_n[:, :3, :3], vid_angle_nearest=np.stack([(int(v) / total) * 2 * math.pi for v in vn]),
               c2w=c2w, campos=c2w[:3, 3], camrotc2w=c2w[:3, :3], middle=np.float32(np.linalg.norm(c2w[:3, 3]) + 0.7),
               near=np.float32(near_far[0]), far=np.float32(near_far[1]), h=H, w=W)
    if random_sample == "patch":
        s = random_sample_size
        x0 = np.random.randint(m, W - m - s + 1)
        y0 = np.random.randint(m, H - m - s + 1)
        px, py = np.meshgrid(np.arange(x0, x0 + s).astype(np.float32), np.arange(y0, y0 + s).astype(np.float32))
    elif random_sample == "dilated":
        st = dilation_setup.split("_")
        px, py = dilated_pixels(W, H, m, int(st[0]), int(st[1]), np.arange(float(st[2]), float(st[3]) + 1))
    else:
        px, py = np.meshgrid(np.arange(m, W - m).astype(np.float32), np.arange(m, H - m).astype(np.float32))
    pix = np.stack((px, py), axis=-1).astype(np.float32)
    out["pixel_idx"] = pix
    out["raydir"] = dtu_raydir(pix, intrinsic.astype(np.float32), c2w[:3, :3], dir_norm > 0).reshape(-1, 3).astype(np.float32)
    full = images_u8[row[vid]].astype(np.float32) / np.float32(255)
    out["gt_image"] = full[py.astype(np.int32), px.astype(np.int32)].reshape(-1, 3)
    if bg_color == "random":
        out["bg_color"] = np.full(3, 1.0 if np.random.rand() > 0.5 else 0.0, np.float32)
    elif bg_color is not None:
        out["bg_color"] = np.asarray(bg_color, np.float32)
    return out
