"""-m gpu: the headless CLI (computer-graphics-engine_b200/host/cge_cli.cpp) - the reference's command-line
mode (src/main.cpp:478-535) and config reader (src/config.cpp:193-374) on the GPU path.  Every BMP it writes must equal, byte
for byte, the file the REFERENCE's own Screen::writeBitmapToFile (stb BMP writer) produces from the same frame rendered
through the Python glue: that covers the TOML reader, the light override, the per-camera threads, the RGBA8 output stage and
the BMP writer at once."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "tests" / "cpp" / "cge_cli"


def run_cli(config_text, tmp_path):
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    cfg = tmp_path / "config.toml"
    cfg.write_text(config_text)
    return subprocess.run([str(CLI), str(cfg)], capture_output=True, text=True)


def test_cli_two_cameras_point_light(cge, ref, tmp_path):
    out = tmp_path / "out"
    r = run_cli(f"""
# a comment
command_line_rendering = true
window_size = [200, 120]
data_path = "{cge.configs.SCENE_DIR}"
scene = "cornell.cges"
output_dir = "{out}"

[features]
enable_shading = true
enable_recursive = true
enable_hard_shadow = true
enable_accel_structure = true

[render]
ray_depth = 3
timestamp = false

[[cameras]]
field_of_view = 50.0
distance_from_look_at = 1.9
look_at = [0.0, 0.0, 0.0]
rotation = [10.0, 20.0, 0.0]

[[cameras]]
field_of_view = 40.0
distance_from_look_at = 2.5
look_at = [0.0, 0.1, 0.0]
rotation = [25.0, -30.0,
            0.0]   # arrays may span lines

[[lights]]
type = "point"
position = [0.1, 0.6, 0.2]
color = [0.9, 0.8, 0.7]
""", tmp_path)
    assert r.returncode == 0, r.stderr
    assert "2 images rendered" in r.stdout
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / "cornell.cges")
    flat.set_lights([(0, [0.1, 0.6, 0.2, 0.9, 0.8, 0.7])])
    cams = [{"fov_deg": 50.0, "dist": 1.9, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [10.0, 20.0, 0.0]},
            {"fov_deg": 40.0, "dist": 2.5, "look_at": [0.0, 0.1, 0.0], "rotation_deg": [25.0, -30.0, 0.0]}]
    with cge.Scene(flat) as sc:
        for i, cam in enumerate(cams):
            cfg = dict(cge.configs.get("c1_cornell", 200, 120), camera=cam, ray_depth=3)
            rgb, _, _ = sc.render(cfg, want_ids=False)
            want = tmp_path / f"want_{i}.bmp"
            ref.write_bmp(rgb, want)
            got = out / f"cornell_cam_{i}.bmp"
            assert got.exists(), r.stdout
            assert got.read_bytes() == want.read_bytes()


def test_cli_soft_shadows_and_extras(cge, ref, tmp_path):
    out = tmp_path / "out"
    r = run_cli(f"""
window_size = [160, 90]
data_path = "{cge.configs.SCENE_DIR}"
scene = "teapot.cges"
output_dir = "{out}"
[features]
enable_shading = true
enable_soft_shadow = true
enable_accel_structure = true
[features.extra]
enable_bloom_effect = true
enable_multiple_rays_per_pixel = true
[render]
parallelogram_light_samples = 3
rays_per_pixel_side = 2
bloom_threshold = 0.25
seed = 77
timestamp = false
[[cameras]]
field_of_view = 50.0
distance_from_look_at = 1.6
look_at = [0.0, 0.0, 0.0]
rotation = [25.0, 30.0, 0.0]
[[lights]]
type = "parallelogram"
corner = [-0.3, 0.9, -0.3]
edges = [[0.6, 0.0, 0.0], [0.0, 0.0, 0.6]]
colors = [[1.0, 0.9, 0.8], [0.8, 0.9, 1.0], [0.9, 1.0, 0.8], [1.0, 1.0, 1.0]]
""", tmp_path)
    assert r.returncode == 0, r.stderr
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / "teapot.cges")
    flat.set_lights([(2, [-0.3, 0.9, -0.3, 0.6, 0.0, 0.0, 0.0, 0.0, 0.6, 1.0, 0.9, 0.8, 0.8, 0.9, 1.0, 0.9, 1.0, 0.8, 1.0, 1.0, 1.0])])
    C = cge.configs
    cfg = dict(C.get("c3_teapot_soft", 160, 90), ray_depth=5, parallelogram_samples=3, seed=77, rays_per_pixel_side=2, bloom_threshold=0.25,
               features=C.FEAT_SHADING | C.FEAT_SOFT_SHADOW | C.FEAT_ACCEL_STRUCTURE | C.FEAT_BLOOM_EFFECT | C.FEAT_MULTIPLE_RAYS_PER_PIXEL,
               camera={"fov_deg": 50.0, "dist": 1.6, "look_at": [0.0, 0.0, 0.0], "rotation_deg": [25.0, 30.0, 0.0]})
    with cge.Scene(flat) as sc:
        rgb, _, _ = sc.render(cfg, want_ids=False)
    want = tmp_path / "want.bmp"
    ref.write_bmp(rgb, want)
    assert (out / "teapot_cam_0.bmp").read_bytes() == want.read_bytes()


def test_cli_refuses_unsupported_extra_and_bad_config(cge, tmp_path):
    base = f"""
window_size = [32, 32]
data_path = "{cge.configs.SCENE_DIR}"
scene = "cube.cges"
output_dir = "{tmp_path / 'out'}"
[features]
enable_shading = true
[[cameras]]
field_of_view = 50.0
"""
    r = run_cli(base + "[features.extra]\nenable_motion_blur = true\n", tmp_path)
    assert r.returncode == 1 and "enableBloomEffect and enableMultipleRaysPerPixel" in r.stderr
    r = run_cli(base.replace('scene = "cube.cges"', 'scene = "no_such.cges"'), tmp_path)
    assert r.returncode == 1 and "does not exist" in r.stderr
    r = run_cli(base + "window_size = [1, 2\n", tmp_path)
    assert r.returncode == 1 and "config line" in r.stderr


def test_cli_renders_an_obj_scene(cge, ref, tmp_path):
    """The scene named by the config is an OBJ file in data_path, as in the reference (src/config.cpp:216-235, src/main.cpp:491-500):
    host/cge_scene_io.hpp loads OBJ + MTL (quads, several materials, shared vertices), the config's lights are the scene's lights.
    The bitmap must equal the one rendered from the same scene exported as a flat scene file by tests/cpp/scene_export."""
    data = tmp_path / "data"
    data.mkdir()
    (data / "room.mtl").write_text("newmtl floor\nKd 0.7 0.7 0.6\nKs 0.2 0.2 0.2\nNs 30\nnewmtl block\nKd 0.2 0.4 0.9\nNs 5\n")
    (data / "room.obj").write_text(
        "mtllib room.mtl\n"
        "v -2 0 -2\nv 2 0 -2\nv 2 0 2\nv -2 0 2\n"
        "v -0.5 0 -0.5\nv 0.5 0 -0.5\nv 0.5 0 0.5\nv -0.5 0 0.5\nv -0.5 1 -0.5\nv 0.5 1 -0.5\nv 0.5 1 0.5\nv -0.5 1 0.5\n"
        "o floor\nusemtl floor\nf 4 3 2 1\n"
        "o block\nusemtl block\nf 9 10 6 5\nf 10 11 7 6\nf 11 12 8 7\nf 12 9 5 8\nf 12 11 10 9\n")
    out = tmp_path / "out"
    r = run_cli(f"""
window_size = [240, 160]
data_path = "{data}"
scene = "room.obj"
output_dir = "{out}"
[features]
enable_shading = true
enable_hard_shadow = true
enable_recursive = true
enable_accel_structure = true
[render]
ray_depth = 2
timestamp = false
[[cameras]]
field_of_view = 50.0
distance_from_look_at = 5.0
look_at = [0.0, 0.4, 0.0]
rotation = [30.0, 35.0, 0.0]
[[lights]]
type = "point"
position = [1.5, 3.0, -2.0]
color = [1.0, 0.95, 0.9]
""", tmp_path)
    assert r.returncode == 0, r.stderr
    flat_path = tmp_path / "room.cges"
    subprocess.run([str(ROOT / "tests" / "cpp" / "scene_export"), "obj", str(data / "room.obj"), "0", str(flat_path)], check=True)
    flat = cge.scenefile.load(flat_path)
    assert len(flat.meshes) == 2 and flat.n_triangles == 12 and len(flat.vertices) == 4 + 5 * 4
    flat.set_lights([(0, [1.5, 3.0, -2.0, 1.0, 0.95, 0.9])])
    cfg = dict(cge.configs.get("c1_cornell", 240, 160), ray_depth=2,
               camera={"fov_deg": 50.0, "dist": 5.0, "look_at": [0.0, 0.4, 0.0], "rotation_deg": [30.0, 35.0, 0.0]})
    with cge.Scene(flat) as sc:
        rgb, ids, _ = sc.render(cfg)
    assert (ids >= 0).mean() > 0.3
    want = tmp_path / "want.bmp"
    ref.write_bmp(rgb, want)
    assert (out / "room_cam_0.bmp").read_bytes() == want.read_bytes()
