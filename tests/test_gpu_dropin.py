"""-m gpu: the C++ layer that mirrors the reference interface (renderRayTracing / getFinalColor / BvhInterface / Trackball /
Screen, computer-graphics-engine_b200/host/cge_engine.hpp) driven from a headless C++ main, against
 (a) the same frame rendered through the Python/ctypes glue (bit-identical: both only call the C ABI), and
 (b) the reference's OWN renderRayTracing (depth literal 5, src/render.cpp:318) run live by oracle/_ref."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import compare_images

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def run_cpp(scene, cfg, out):
    exe = ROOT / "tests" / "cpp" / "drop_in_main"
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    cam = cfg["camera"]
    r = subprocess.run([str(exe), str(scene), str(cfg["width"]), str(cfg["height"]), str(cam["fov_deg"]), str(cam["dist"]),
                        str(cam["rotation_deg"][0]), str(cam["rotation_deg"][1]), hex(cfg["features"]), str(cfg["ray_depth"]),
                        str(cfg["parallelogram_samples"]), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(out, np.float32).reshape(cfg["height"], cfg["width"], 3), r.stdout


@pytest.mark.parametrize("name", ["c1_cornell", "c2_cube_textured"])
def test_cpp_mirror_equals_reference_renderRayTracing(cge, ref, name, tmp_path):
    cfg = cge.configs.get(name, 200, 120)
    cfg["ray_depth"] = 5  # what renderRayTracing hard-codes
    scene = cge.configs.scene_path(cfg)
    rgb_cpp, out = run_cpp(scene, cfg, tmp_path / "o.raw")
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_py, _, _ = sc.render(cfg, want_ids=False)
    assert rgb_cpp.tobytes() == rgb_py.tobytes()
    with ref.RefScene(scene, cfg["features"]) as rs:
        rgb_ref, _, _ = rs.render(cfg, want_ids=False, use_render_ray_tracing=True)
        info = rs.bvh_info()
    err, nan_mm = compare_images(rgb_cpp, rgb_ref)
    assert nan_mm == 0 and err <= 1e-3 * max(1.0, float(np.nan_to_num(rgb_ref, nan=0).max()))
    assert f"levels {info['levels']} leaves {info['leaves']}" in out


def test_cpp_mirror_soft_shadow_config(cge, tmp_path):
    cfg = cge.configs.get("c3_teapot_soft", 160, 90)
    g_rgb, out = run_cpp(cge.configs.scene_path(cfg), dict(cfg, seed=0), tmp_path / "o.raw")
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_py, _, _ = sc.render(dict(cfg, seed=0), want_ids=False)
    assert g_rgb.tobytes() == rgb_py.tobytes()
