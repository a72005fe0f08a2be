"""option sets of the frame-producer golden (tests/golden/frame.npz): shared by the generator and the tests"""
FRAME_CASES = [   # (name, split, id, opt overrides, rng seed, bg_color)
    ("a", "train", 4, dict(use_nearest=3, find_nearest_mode=0, random_sample="dilated", dilation_setup="3_4_1_4", dir_norm=0, edge_filter=2), 5, (1.0, 1.0, 1.0)),
    ("b", "test", 1, dict(use_nearest=4, find_nearest_mode=1, random_sample="no_crop", dir_norm=1, edge_filter=3), 6, (0.0, 0.0, 0.0)),
    ("c", "train", 0, dict(use_nearest=2, find_nearest_mode=1, random_sample="patch", random_sample_size=8, dir_norm=0, edge_filter=0), 7, "random"),
    ("d", "train", 7, dict(use_nearest=4, find_nearest_mode=0, dynamic_nearest=1, random_sample="dilated", dilation_setup="2_8_1_3", dir_norm=0, edge_filter=0), 8, (1.0, 1.0, 1.0)),
    ("e", "test", 2, dict(use_nearest=4, find_nearest_mode=0, random_sample="no_crop", dir_norm=0, edge_filter=0), 9, (1.0, 1.0, 1.0)),
]

# further option sets checked on the CPU only (oracle vs reference, stubbed host path of the product)
FRAME_CASES_CPU = [
    ("f", "train", 2, dict(use_nearest=0, find_nearest_mode=0, random_sample="dilated", dilation_setup="2_4_2_2", dir_norm=0, edge_filter=1), 10, (1.0, 1.0, 1.0)),
    ("g", "test", 0, dict(use_nearest=6, find_nearest_mode=1, random_sample="patch", random_sample_size=16, dir_norm=1, edge_filter=4), 11, "random"),
]
