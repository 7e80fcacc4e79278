"""Render one config a few times in one mode (profiling target).  usage: one_render.py cfg mode[fast-default|fast-wave|fast-thread|reference] [n] [scale]"""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
name, mode = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
full = pkg.configs.get(name)
cfg = pkg.configs.get(name, int(full["width"] * scale), int(full["height"] * scale))
trav, flags = {"fast-default": (1, 0), "fast-wave": (1, pkg.FLAG_WAVEFRONT), "fast-thread": (1, pkg.FLAG_PER_THREAD), "reference": (0, 0)}[mode]
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    for _ in range(n):
        _, _, st = sc.render(cfg, traversal=trav, want_ids=False, flags=flags)
    print(name, mode, "kernel_ms", round(st["kernel_ms"], 3), "gpu_rays", st["gpu_rays"])
