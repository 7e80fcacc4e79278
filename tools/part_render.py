"""Render one partition k/N of a config on one GPU and print stage times (per-rank cost of the multi-GPU path)."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
name = sys.argv[1]
parts = [int(x) for x in sys.argv[2].split(",")]
flags = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0
cfg = pkg.configs.get(name)
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    for n in parts:
        best = None
        for _ in range(4):
            _, _, st = sc.render(cfg, want_ids=False, part=(0, n), flags=flags)
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        print(name, "part 0 of", n, "kernel_ms", round(best["kernel_ms"], 3), "stages", [round(x, 3) for x in best["stage_ms"]],
              "gpu_rays", best["gpu_rays"], flush=True)
