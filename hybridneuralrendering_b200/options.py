"""The `opt` fields the hot path reads, with the values every shipped dev_script uses
(reference: dev_scripts/w_scannet_etf/scene241_full.sh, dev_scripts/w_n360/lego_hybrid.sh;
SURVEY.md §8d).  The host modules take any object with these attributes (e.g. the reference's own
argparse namespace), so this is only a convenience for tests, the bench and standalone use."""
from __future__ import annotations

from types import SimpleNamespace

SCANNET = dict(
    # query
    vsize=[0.008, 0.008, 0.008], vscale=[2, 2, 2], kernel_size=[3, 3, 3], query_size=[3, 3, 3],
    ranges=[-10.0, -10.0, -10.0, 10.0, 10.0, 10.0], radius_limit_scale=4.0, depth_limit_scale=0.0,
    SR=24, K=8, P=26, max_o=610000, NN=2, z_depth_dim=400, inverse=0, gpu_maxthr=1024, wcoord_query=1,
    near_plane=0.1, far_plane=8.0,
    # points
    point_features_dim=32, point_conf_mode="1", point_dir_mode="1", point_color_mode="1",
    xyz_grad=0, feat_grad=1, conf_grad=1, dir_grad=1, color_grad=1, load_points=0, default_conf=-1.0,
    # aggregator
    agg_dist_pers=20, agg_intrp_order=2, agg_distance_kernel="linear", agg_weight_norm=1, agg_axis_weight=None,
    act_type="LeakyReLU", act_super=1, apply_pnt_mask=1, which_agg_model="viewmlp",
    shading_feature_mlp_layer0=1, shading_feature_mlp_layer1=2, shading_feature_mlp_layer2=0, shading_feature_mlp_layer3=2,
    shading_alpha_mlp_layer=1, shading_color_mlp_layer=4, shading_feature_num=256, shading_color_channel_num=3,
    num_feat_freqs=3, dist_xyz_freq=5, dist_xyz_deno=0, num_pos_freqs=10, num_viewdir_freqs=4, view_ori=0,
    agg_feat_xyz_mode="None", agg_alpha_xyz_mode="None", agg_color_xyz_mode="None",
    use_nearest=4, dynamic_nearest=0, feature_guidance=1, use_delta_view=1, mixup_mode="partial", learn_residuals=1,
    refine_blend=0, dynamic_weight=0, tradition_attention=0, add_idx=0, downweight_blurry_feats=0,
    separate_color_decoder=0, large_color_final_block=0, use_2D_CNN=0, disable_viewdirs=0, disable_color_feature=0,
    search_size=0, search_dilation=0,
    drop_ratio=0.5, drop_patch=1, ray_points=1, random_position=1, drop_disturb_range=0, dilation_setup="7_8_1_8",
    learnable_blur_kernel=0, learnable_blur_kernel_conv=0, learnable_blur_kernel_size=9, learnable_blur_patch_size=8,
    learnable_blur_kernel_mode=4, learnable_blur_kernel_norm=0, boundary_mode=0,
    # conductor / compositing
    raydist_mode_unit=1, which_render_func="radiance", which_blend_func="alpha", which_tonemap_func="off",
    is_train=False, prob=0, zero_one_loss_items="conf_coefficient", sparse_loss_weight=0, bg_color="white",
    # blur
    add_blur_sim=1, blur_kernel_version=3, blur_kernel_size=9, num_move_dirs=8, move_dists="1,2,4",
)

LEGO = dict(SCANNET, vsize=[0.004, 0.004, 0.004], ranges=[-0.638, -1.141, -0.346, 0.634, 1.149, 1.141], SR=80, P=12,
            max_o=2000000, near_plane=2.0, far_plane=6.0)


def make_opt(base: str = "scannet", **over) -> SimpleNamespace:
    d = dict(SCANNET if base == "scannet" else LEGO)
    d.update(over)
    return SimpleNamespace(**d)
