"""Pinned cases for the two implemented ExtraFeatures: multiple rays per pixel and bloom.
Shared by tests/golden/make_golden.py (which renders them with the unmodified reference) and the parity tests."""
import importlib

_cfgs = importlib.import_module("computer-graphics-engine_b200.configs")
AA, BLOOM = _cfgs.FEAT_MULTIPLE_RAYS_PER_PIXEL, _cfgs.FEAT_BLOOM_EFFECT

# key -> (config, width, height, extra feature bits, overrides)
CASES = {
    "c1_aa2": ("c1_cornell", 96, 96, AA, {"rays_per_pixel_side": 2, "seed": 11}),
    "c1_bloom": ("c1_cornell", 96, 96, BLOOM, {"bloom_threshold": 0.2, "bloom_scalar": 0.3}),
    "c2_aa3_bloom": ("c2_cube_textured", 96, 54, AA | BLOOM, {"rays_per_pixel_side": 3, "bloom_threshold": 0.1, "seed": 3}),
    "c3_aa2": ("c3_teapot_soft", 96, 54, AA, {"rays_per_pixel_side": 2}),  # soft shadows: one draw counter across the sub-rays
    "c3_bloom_only": ("c3_teapot_soft", 96, 54, BLOOM, {"bloom_threshold": 0.3, "bloom_debug_option": 1}),
    "c4_aa3_bloom": ("c4_monkey_mirror", 64, 36, AA | BLOOM, {"rays_per_pixel_side": 3, "bloom_scalar": 0.5, "seed": 99}),
    "c1_aa10": ("c1_cornell", 24, 24, AA, {"rays_per_pixel_side": 10, "seed": 1}),  # the GUI maximum: 200 of the 227 draws
    "c1_odd_bloom": ("c1_cornell", 33, 17, BLOOM, {"bloom_threshold": 0.25, "bloom_scalar": 0.6}),  # ragged frame; last column / bottom row untouched
}


def cfg_for(key: str) -> dict:
    name, w, h, extra, over = CASES[key]
    cfg = _cfgs.get(name, w, h)
    cfg["features"] |= extra
    cfg.update(over)
    return cfg
