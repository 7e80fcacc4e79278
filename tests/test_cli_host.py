"""CPU test (-m "not gpu"): the CLI reads its config and then fails LOUDLY where the GPU work would start - there is no CPU
rendering path behind it (computer-graphics-engine_b200/host/cge_cli.cpp)."""
import importlib
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pkg = importlib.import_module("computer-graphics-engine_b200")


def test_cli_without_device_reports_cuda_error(tmp_path):
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    cfg = tmp_path / "c.toml"
    cfg.write_text(f"""
window_size = [16, 16]
data_path = "{pkg.configs.SCENE_DIR}"
scene = "cube.cges"
output_dir = "{tmp_path}"
[features]
enable_shading = true
[[cameras]]
field_of_view = 50.0
""")
    r = subprocess.run([str(ROOT / "tests" / "cpp" / "cge_cli"), str(cfg)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    bad = tmp_path / "bad.toml"
    bad.write_text("window_size = [16, 16\n")
    r = subprocess.run([str(ROOT / "tests" / "cpp" / "cge_cli"), str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "config line" in r.stderr


def test_cli_reads_the_reference_config_keys(tmp_path):
    """--print-config: the TOML subset reader and the reference's keys / defaults (src/config.cpp:193-374), no GPU involved."""
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    cfg = tmp_path / "c.toml"
    cfg.write_text("""
# comment line
command_line_rendering = true   # trailing comment
window_size = [640, 360]
scene = "teapot.cges"
output_dir = "somewhere/out"

[features]
enable_shading = true
enable_soft_shadow = true
enable_accel_structure = true

[features.extra]
enable_bloom_effect = true
enable_multiple_rays_per_pixel = true

[render]
ray_depth = 2
parallelogram_light_samples = 4
rays_per_pixel_side = 2
bloom_threshold = 0.25
seed = 77

[[cameras]]
field_of_view = 45.5
look_at = [0.0, 0.5, -1]

[[cameras]]
distance_from_look_at = 2.25
rotation = [ 10.0,
             -20.0,   # a comment inside an array
             0.0 ]

[[lights]]
type = "parallelogram"
corner = [0, 1, 0]
edges = [[1, 0, 0], [0, 0, 1]]
colors = [[1, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1]]

[[lights]]
type = "point"
position = [0, 1, 0]
color = [1, 1, 1]
""")
    r = subprocess.run([str(ROOT / "tests" / "cpp" / "cge_cli"), str(cfg), "--print-config"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    assert out[0] == "window_size 640 360" and out[1] == "scene teapot.cges" and out[3] == "output_dir somewhere/out"
    assert out[4] == "features shading 1 recursive 0 hard_shadow 0 soft_shadow 1 normal_interp 0 texture_mapping 0 accel_structure 1"
    assert out[5] == "extra bits 0x48"  # bloom (bit 3) + multiple rays per pixel (bit 6)
    assert out[6] == "render ray_depth 2 segment_light_samples 25 parallelogram_light_samples 4 rays_per_pixel_side 2 bloom 0.3 0.25 0 seed 77"
    assert out[7] == "camera fov 45.5 dist 3 look_at 0 0.5 -1 rotation 20 20 0"  # defaults of src/config.cpp:325-330
    assert out[8] == "camera fov 50 dist 2.25 look_at 0 0 0 rotation 10 -20 0"
    assert out[9:] == ["light parallelogram", "light point"]


def test_cli_reads_the_other_toml_spellings(tmp_path):
    """The same config written the other way TOML allows: inline tables, dotted keys, literal and multi-line strings, integers with
    underscores / in hex / in binary, floats with exponents and signs, arrays over several lines with a trailing comma."""
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    cfg = tmp_path / "c.toml"
    cfg.write_text("""window_size = [ 1_280, 0x2D0 ]
data_path = 'some\\\\where'   # literal string: the backslashes stay
scene = \"\"\"
teapot.cges\"\"\"
features.enable_shading = true
features.extra = { enable_bloom_effect = true }
cameras = [
  { field_of_view = 4.55e1, look_at = [0.0, 0.5, -1], rotation = [1e1, -2_0.0, +0.0] },
  { distance_from_look_at = 2.25 },
]
lights = [ { type = "point", position = [0, 1, 0], color = [1, 1, 1] },
           { type = "segment", endpoints = [[0, 1, 0], [1, 1, 0]], colors = [[1, 0, 0], [0, 1, 0]] } ]
[render]
seed = 0b1001101
bloom_threshold = 2.5e-1
""")
    r = subprocess.run([str(ROOT / "tests" / "cpp" / "cge_cli"), str(cfg), "--print-config"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    assert out[0] == "window_size 1280 720" and out[1] == "scene teapot.cges" and out[2] == "data_path some\\\\where"
    assert out[4] == "features shading 1 recursive 0 hard_shadow 0 soft_shadow 0 normal_interp 0 texture_mapping 0 accel_structure 0"
    assert out[5] == "extra bits 0x8"
    assert out[6].endswith("bloom 0.3 0.25 0 seed 77")
    assert out[7] == "camera fov 45.5 dist 3 look_at 0 0.5 -1 rotation 10 -20 0"
    assert out[8] == "camera fov 50 dist 2.25 look_at 0 0 0 rotation 20 20 0"
    assert out[9:] == ["light point", "light segment"]
    for bad in ("a = { b = 1", "a = 'unterminated", "a = 1__", "a = 0xZZ"):
        cfg.write_text(bad + "\\n")
        r = subprocess.run([str(ROOT / "tests" / "cpp" / "cge_cli"), str(cfg), "--print-config"], capture_output=True, text=True)
        assert r.returncode == 1 and "config line" in r.stderr, bad
