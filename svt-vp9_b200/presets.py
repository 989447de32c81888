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
