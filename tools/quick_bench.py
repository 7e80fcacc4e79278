"""Ad-hoc timing of every config (development aid; bench.py is the judged entry)."""
import importlib, sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
names = sys.argv[1:] or list(pkg.configs.CONFIGS)
for name in names:
    scale = 1.0
    if ":" in name:
        name, s = name.split(":"); scale = float(s)
    full = pkg.configs.get(name)
    cfg = pkg.configs.get(name, int(full["width"] * scale), int(full["height"] * scale))
    t0 = time.time(); flat = pkg.load_scene(cfg); t1 = time.time()
    with pkg.Scene(flat) as sc:
        t2 = time.time()
        modes = [(1, 0, "fast-default"), (1, pkg.FLAG_WAVEFRONT, "fast-wave"), (1, pkg.FLAG_PER_THREAD, "fast-thread"), (0, 0, "reference")]
        if name == "c5_dragon" and scale > 0.3:
            modes = modes[:3]
        for trav, flags, label in modes:
            best = None
            for it in range(3):
                rgb, ids, st = sc.render(cfg, traversal=trav, want_ids=False, flags=flags)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            print(json.dumps({"cfg": name, "w": cfg["width"], "h": cfg["height"], "mode": label,
                              "kernel_ms": round(best["kernel_ms"], 3), "total_ms": round(best["total_ms"], 3),
                              "gpu_rays": best["gpu_rays"], "ref_rays": best["reference_rays"],
                              "Mrays_s": round(best["gpu_rays"] / best["kernel_ms"] / 1e3, 1),
                              "ref_Mrays_s": round(best["reference_rays"] / best["kernel_ms"] / 1e3, 1),
                              "load_s": round(t1 - t0, 2), "upload_s": round(t2 - t1, 2),
                              "nan_px": int(np.isnan(rgb).any(-1).sum())}), flush=True)
