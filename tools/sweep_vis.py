"""A/B sweep over the environment-tunable variants of the shadow-ray pass (development aid): kernel ms per variant, every
variant checked bit-identical to the first.  usage: sweep_vis.py cfg[:scale] "K=V,K=V" "K=V" ...   (SWEEP_PART=n: 1/n share)"""
import hashlib, importlib, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
spec = sys.argv[1] if len(sys.argv) > 1 else "c5_dragon"
variants = sys.argv[2:] or [""]
name, scale = (spec.split(":") + ["1.0"])[:2]
full = pkg.configs.get(name)
cfg = pkg.configs.get(name, int(full["width"] * float(scale)), int(full["height"] * float(scale)))
part = (0, int(os.environ.get("SWEEP_PART", "1")))
keys = sorted({kv.split("=")[0] for v in variants for kv in v.split(",") if kv})
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    base = None
    for v in variants:
        for k in keys:
            os.environ.pop(k, None)
        for kv in v.split(","):
            if kv:
                os.environ[kv.split("=")[0]] = kv.split("=")[1]
        best = None
        for _ in range(4):
            rgb, _, st = sc.render(cfg, traversal=1, want_ids=False, part=part)
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        if base is None:
            base = rgb.tobytes()
        print(json.dumps({"cfg": name, "w": cfg["width"], "part": part[1], "variant": v, "kernel_ms": round(best["kernel_ms"], 3),
                          "stages": [round(x, 3) for x in best["stage_ms"]], "shadow_rays": best["shadow_rays"], "culled": best.get("shadow_samples_culled"),
                          "identical": rgb.tobytes() == base,
                          "sha": hashlib.sha256(rgb.tobytes()).hexdigest()[:12]}), flush=True)
