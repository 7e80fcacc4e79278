"""Kernel ms of the default dispatch for several configs (development aid; run under CGE_LIB=... for A/B builds)."""
import importlib, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("computer-graphics-engine_b200")
import torch
for name in sys.argv[1:] or ["c1_cornell", "c2_cube_textured", "c3_teapot_soft", "c4_monkey_mirror", "c5_dragon"]:
    cfg = pkg.configs.get(name)
    frame = torch.zeros((cfg["height"], cfg["width"], 3), dtype=torch.float32, device="cuda")
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        best = None
        for _ in range(6):
            st = sc.render_device(cfg, frame.data_ptr())
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        print(json.dumps({"lib": os.environ.get("CGE_LIB", "default"), "cfg": name, "kernel_ms": round(best["kernel_ms"], 3),
                          "stages": [round(x, 3) for x in best["stage_ms"]]}), flush=True)
