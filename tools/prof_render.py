"""Profiling target: render one config n times as ONE pipeline with the environment's variant.  usage: prof_render.py cfg[:scale] [n] [part_count]"""
import importlib, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CGE_BANDS", "1")
pkg = importlib.import_module("computer-graphics-engine_b200")
name, scale = (sys.argv[1].split(":") + ["1.0"])[:2]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
part = (0, int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1)
full = pkg.configs.get(name)
cfg = pkg.configs.get(name, int(full["width"] * float(scale)), int(full["height"] * float(scale)))
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    for _ in range(n):
        _, _, st = sc.render(cfg, traversal=1, want_ids=False, part=part)
    print(name, "kernel_ms", round(st["kernel_ms"], 3), "stages", [round(x, 3) for x in st["stage_ms"]])
