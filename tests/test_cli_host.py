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
