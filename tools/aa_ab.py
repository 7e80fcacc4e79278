"""Multi-sample frames with area lights: wavefront pipeline vs per-thread kernel (development aid)."""
import importlib, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("computer-graphics-engine_b200")
import torch
for name, scale in (("c3_teapot_soft", 1.0), ("c5_dragon", 0.5)):
    full = pkg.configs.get(name)
    cfg = pkg.configs.get(name, int(full["width"] * scale), int(full["height"] * scale))
    cfg["features"] |= pkg.configs.FEAT_MULTIPLE_RAYS_PER_PIXEL
    cfg["rays_per_pixel_side"] = 2
    frame = torch.zeros((cfg["height"], cfg["width"], 3), dtype=torch.float32, device="cuda")
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        imgs = {}
        for label, fl in (("wavefront", 0), ("per-thread", pkg.FLAG_PER_THREAD)):
            best = None
            for _ in range(3):
                st = sc.render_device(cfg, frame.data_ptr(), flags=fl)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            imgs[label] = frame.cpu().numpy().tobytes()
            print(json.dumps({"cfg": name, "w": cfg["width"], "aa": 2, "mode": label, "kernel_ms": round(best["kernel_ms"], 3),
                              "stages": [round(x, 3) for x in best["stage_ms"]], "rays": best["gpu_rays"]}), flush=True)
        print("identical:", imgs["wavefront"] == imgs["per-thread"], flush=True)
