"""The five BASELINE.json configurations, pinned: scene file, camera, Features, depth, sample counts.

C1/C2 are reference scenes (reference src/scene.cpp:29-39).  C3-C5 are NOT reference scenes:
they are composed from reference meshes by tests/golden/make_scenes.py (C3, C4) or generated procedurally
(C5: ``data/dragon.obj`` is absent from the reference checkout, so a declared stand-in of 868 332 triangles is
used on BOTH sides — see standin.py).  Cameras use the reference's CameraConfig fields (src/config.h:16-21:
field of view in degrees, distance from look-at, look-at point, Euler rotation in degrees).
"""
from __future__ import annotations

from pathlib import Path

FEAT_SHADING = 1 << 0
FEAT_RECURSIVE = 1 << 1
FEAT_HARD_SHADOW = 1 << 2
FEAT_SOFT_SHADOW = 1 << 3
FEAT_NORMAL_INTERP = 1 << 4
FEAT_TEXTURE_MAPPING = 1 << 5
FEAT_ACCEL_STRUCTURE = 1 << 6
# the two implemented ExtraFeatures (bit 16 + position in the reference's ExtraFeatures struct, src/common.h:54-65)
FEAT_BLOOM_EFFECT = 1 << 19
FEAT_MULTIPLE_RAYS_PER_PIXEL = 1 << 22

SCENE_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden" / "scenes"

CONFIGS = {
    # CornellBox-Mirror-Rotated.obj 1024x1024, BVH + Phong + hard shadows + recursion depth 3
    "c1_cornell": {
        "scene": "cornell.cges",
        "width": 1024, "height": 1024,
        "features": FEAT_SHADING | FEAT_RECURSIVE | FEAT_HARD_SHADOW | FEAT_ACCEL_STRUCTURE,
        "ray_depth": 3,
        "camera": {"fov_deg": 50.0, "dist": 1.9, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [10.0, 20.0, 0.0]},
    },
    # cube-textured.obj 1920x1080, texture mapping + barycentric normal interpolation
    "c2_cube_textured": {
        "scene": "cube_textured.cges",
        "width": 1920, "height": 1080,
        "features": FEAT_SHADING | FEAT_TEXTURE_MAPPING | FEAT_NORMAL_INTERP | FEAT_ACCEL_STRUCTURE,
        "ray_depth": 5,
        "camera": {"fov_deg": 50.0, "dist": 3.0, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [20.0, 20.0, 0.0]},
    },
    # teapot.obj 1920x1080, one parallelogram light, 4x4 = 16 shadow samples per hit
    "c3_teapot_soft": {
        "scene": "teapot_area.cges",
        "width": 1920, "height": 1080,
        "features": FEAT_SHADING | FEAT_SOFT_SHADOW | FEAT_ACCEL_STRUCTURE,
        "ray_depth": 0,
        "parallelogram_samples": 4,
        "seed": 20261018,
        "camera": {"fov_deg": 50.0, "dist": 1.6, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [25.0, 30.0, 0.0]},
    },
    # monkey.obj inside the Cornell box with mirror walls, 2560x1440, recursion depth 6
    "c4_monkey_mirror": {
        "scene": "monkey_mirror.cges",
        "width": 2560, "height": 1440,
        "features": FEAT_SHADING | FEAT_RECURSIVE | FEAT_HARD_SHADOW | FEAT_ACCEL_STRUCTURE,
        "ray_depth": 6,
        "camera": {"fov_deg": 50.0, "dist": 1.9, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [10.0, 20.0, 0.0]},
    },
    # dragon.obj stand-in (868 332 triangles) 3840x2160, soft shadows + recursion depth 3
    "c5_dragon": {
        "scene": "standin:dragon",
        "width": 3840, "height": 2160,
        "features": FEAT_SHADING | FEAT_RECURSIVE | FEAT_SOFT_SHADOW | FEAT_ACCEL_STRUCTURE,
        "ray_depth": 3,
        "parallelogram_samples": 4,
        "seed": 20261018,
        "camera": {"fov_deg": 50.0, "dist": 2.6, "look_at": [0.0, -0.1, 0.0], "rotation_deg": [20.0, 25.0, 0.0]},
    },
}

DEFAULTS = {"ray_depth": 5, "segment_samples": 25, "parallelogram_samples": 5, "seed": 0}


def get(name: str, width: int | None = None, height: int | None = None) -> dict:
    """Config dict with defaults filled in; width/height override for reduced-size parity cases (the aspect
    ratio, which enters the camera, then follows the override exactly as the reference Window would)."""
    cfg = dict(DEFAULTS)
    cfg.update(CONFIGS[name])
    cfg["name"] = name
    if width is not None:
        cfg["width"] = width
    if height is not None:
        cfg["height"] = height
    return cfg


def scene_path(cfg: dict) -> Path | None:
    s = cfg["scene"]
    return None if s.startswith("standin:") else SCENE_DIR / s
