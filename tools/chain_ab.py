"""A/B of the chain stage: one lane per pixel chain (default) vs one launch per recursion level (CGE_FLAG_CHAIN_PER_LEVEL)."""
import importlib, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("computer-graphics-engine_b200")
import torch
name = sys.argv[1] if len(sys.argv) > 1 else "c5_dragon"
part = (0, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
cfg = pkg.configs.get(name)
frame = torch.zeros((cfg["height"], cfg["width"], 3), dtype=torch.float32, device="cuda")
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    for label, fl in (("per-pixel", 0), ("per-level", pkg.FLAG_CHAIN_PER_LEVEL), ("per-pixel", 0), ("per-level", pkg.FLAG_CHAIN_PER_LEVEL)):
        best = None
        for _ in range(5):
            st = sc.render_device(cfg, frame.data_ptr(), part=part, flags=fl)
            if best is None or st["stage_ms"][0] < best["stage_ms"][0]:
                best = st
        print(json.dumps({"cfg": name, "part": part[1], "chain": label, "kernel_ms": round(best["kernel_ms"], 3),
                          "stages": [round(x, 3) for x in best["stage_ms"]], "launches": best["kernel_launches"],
                          "bounce": best["bounce_rays"], "ref": best["reference_rays"]}), flush=True)
