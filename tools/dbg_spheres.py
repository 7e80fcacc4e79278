"""development aid: fast vs literal traversal on a sphere scene: where do ids / colours differ?"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
C = pkg.configs
scene = sys.argv[1] if len(sys.argv) > 1 else "mixed.cges"
flat = pkg.scenefile.load(C.SCENE_DIR / scene)
print("triangles", len(flat.triangles), "spheres", len(flat.spheres))
base = {"scene": scene, "width": 96, "height": 64, "ray_depth": 2, "segment_samples": 5, "parallelogram_samples": 3, "seed": 5,
        "camera": {"fov_deg": 60.0, "dist": 4.0, "look_at": [0.0, 0.3, 0.0], "rotation_deg": [15.0, 35.0, 0.0]}}
with pkg.Scene(flat) as sc:
    for feats in (C.FEAT_ACCEL_STRUCTURE, C.FEAT_SHADING, C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_HARD_SHADOW):
        cfg = dict(base, features=feats)
        a, ia, _ = sc.render(cfg, traversal=0)
        b, ib, _ = sc.render(cfg, traversal=1, flags=pkg.FLAG_PER_THREAD)
        mm = np.argwhere(ia != ib)
        d = np.abs(np.nan_to_num(a) - np.nan_to_num(b)).max(-1)
        print(f"feats {feats:#x}: id mismatches {len(mm)}, colour mismatches {(d > 1e-3).sum()}")
        for y, x in mm[:8]:
            print("   ", y, x, "literal id", ia[y, x], "fast id", ib[y, x], a[y, x], b[y, x])
        cm = np.argwhere((d > 1e-3) & (ia == ib))
        for y, x in cm[:8]:
            print("   same id", ia[y, x], "at", y, x, a[y, x], b[y, x])
