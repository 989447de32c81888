"""Resolved motion-estimation parameter sets of the BASELINE.json configurations.

The reference derives these per picture in its control plane (set_me_hme_params_* /
eb_vp9_signal_derivation_me_kernel_*, Source/Lib/Codec/EbMotionEstimationProcess.c:55-327, :541-722, tables
Source/Lib/Codec/EbDefinitions.h:6364-6744, multipliers :989-1005) -- out of scope here; the values below were read
back from the unmodified reference with oracle/ref_seam_check.c (dump mode) for 3840x2160 / 1920x1080, enc-mode 8,
tune 1, and are what the reference-side shim passes through SvtB200MeParams at run time.
"""
from . import api

# hme_level_0_search_area_multiplier_x/y[hierarchical_levels = 4][temporal layer] (EbDefinitions.h:995, :1004)
HME_L0_MULT_4LEVELS = [350, 200, 100, 100, 100]

# temporal layer of picture n (1..16) inside a 16-picture mini-GOP with 4 hierarchical levels
MINIGOP16_TL = [4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4, 0]


def me_params_4k_m8(width=3840, height=2160, temporal_layer=4, num_lists=2, ref_poc=(0, 2)):
    """config C: 4K, enc-mode 8, tune 1: single-quadrant HME level 0 (64x32, scaled per layer), 8x7 full-pel area,
    sub-sampled-SAD fractional search gated by su_pel_enable, no 8x8 refinement (cu8x8_mode 1)."""
    P = api.MeParams()
    P.luma_width, P.luma_height = width, height
    P.num_lists, P.temporal_layer_index, P.hierarchical_levels = num_lists, temporal_layer, 4
    P.ref_poc[0], P.ref_poc[1] = ref_poc
    P.enable_hme_flag, P.enable_hme_level0, P.enable_hme_level1, P.enable_hme_level2 = 1, 1, 0, 0
    P.single_hme_quadrant = 1
    P.number_hme_search_region_in_width = P.number_hme_search_region_in_height = 2
    P.hme_level0_total_search_area_width, P.hme_level0_total_search_area_height = 64, 32
    for i in range(2):
        P.hme_level0_search_area_in_width[i], P.hme_level0_search_area_in_height[i] = 32, 16
    P.hme_l0_multiplier_x = P.hme_l0_multiplier_y = HME_L0_MULT_4LEVELS[temporal_layer]
    P.search_area_width, P.search_area_height = 8, 7
    P.fractional_search_method, P.fractional_search_model, P.fractional_search64x64 = 0, 1, 0
    P.cu8x8_mode, P.cu16x16_mode, P.rate_control_mode = 1, 0, 0
    return P


def me_params_1080p_m8(width=1920, height=1080, temporal_layer=4, num_lists=2, ref_poc=(0, 2)):
    """config B: 1080p, enc-mode 8, tune 1: 2x2 quadrants of 32x24 at level 0, 4x4 (widened to 8x4) / 4x2 (widened to 8x2) refinements at levels 1 / 2,
    16x9 full-pel area (read back from the reference like the 4K set)."""
    P = me_params_4k_m8(width, height, temporal_layer, num_lists, ref_poc)
    P.enable_hme_level1 = P.enable_hme_level2 = 1
    P.single_hme_quadrant = 0
    P.hme_level0_total_search_area_width, P.hme_level0_total_search_area_height = 64, 48
    for i in range(2):
        P.hme_level0_search_area_in_width[i], P.hme_level0_search_area_in_height[i] = 32, 24
        P.hme_level1_search_area_in_width[i], P.hme_level1_search_area_in_height[i] = 4, 4
        P.hme_level2_search_area_in_width[i], P.hme_level2_search_area_in_height[i] = 4, 2
    P.search_area_width, P.search_area_height = 16, 9
    return P


# hme_level_0_search_area_multiplier_x/y[hierarchical_levels = 3][temporal layer] (EbDefinitions.h:994, :1003; tune 0 encodes with 3 levels)
HME_L0_MULT_3LEVELS = [200, 140, 100, 70]

# temporal layer of picture n (1..8) inside an 8-picture mini-GOP with 3 hierarchical levels
MINIGOP8_TL = [3, 2, 3, 1, 3, 2, 3, 0]


def me_params_480p_m9(width=640, height=480, temporal_layer=4, num_lists=2, ref_poc=(0, 2)):
    """config A: 640x480, enc-mode 9, tune 1 (read back with oracle/dump_me_params.py 640 480 9 1): 2x2 quadrants of 16x12 at level 0
    (scaled per layer), 4x4 / 4x2 refinements (widened to 8 columns) at levels 1 / 2, 16x7 full-pel area, sub-sampled-SAD
    refinement gated by su_pel_enable, no 8x8 refinement."""
    P = me_params_1080p_m8(width, height, temporal_layer, num_lists, ref_poc)
    P.hme_level0_total_search_area_width, P.hme_level0_total_search_area_height = 32, 24
    for i in range(2):
        P.hme_level0_search_area_in_width[i], P.hme_level0_search_area_in_height[i] = 16, 12
    P.search_area_width, P.search_area_height = 16, 7
    return P


def me_params_4k_m5_tune0(width=3840, height=2160, temporal_layer=3, num_lists=2, ref_poc=(0, 2)):
    """config D: 4K, enc-mode 5, tune 0 (3 hierarchical levels; read back with oracle/dump_me_params.py 3840 2160 5 0): 2x2 quadrants of
    64x32 at level 0 (128x64 in all, scaled 200 / 140 / 100 / 70 % per layer: the wide areas are searched in row stripes), levels 1 / 2
    on, 16x16 full-pel area; 8x8 partitions refined except in the top layer, whose pictures run without sub-pel search
    (use_subpel_flag 0 -> fractional_search_model 2)."""
    P = me_params_1080p_m8(width, height, temporal_layer, num_lists, ref_poc)
    P.hierarchical_levels = 3
    P.hme_level0_total_search_area_width, P.hme_level0_total_search_area_height = 128, 64
    for i in range(2):
        P.hme_level0_search_area_in_width[i], P.hme_level0_search_area_in_height[i] = 64, 32
    P.hme_l0_multiplier_x = P.hme_l0_multiplier_y = HME_L0_MULT_3LEVELS[temporal_layer]
    P.search_area_width, P.search_area_height = 16, 16
    top = temporal_layer == 3
    P.fractional_search_model = 2 if top else 1
    P.cu8x8_mode = 1 if top else 0
    return P
