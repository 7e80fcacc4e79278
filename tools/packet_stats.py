"""development aid: counters of the packet shadow pass (needs a library built with -DCGE_PACKET_STATS, path in CGE_LIB)"""
import importlib, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
spec = sys.argv[1]
name, scale = (spec.split(":") + ["1.0"])[:2]
full = pkg.configs.get(name)
cfg = pkg.configs.get(name, int(full["width"] * float(scale)), int(full["height"] * float(scale)))
os.environ["CGE_BANDS"] = "1"
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    os.environ["CGE_PACKET"] = "0"
    _, _, b = sc.render(cfg, traversal=1, want_ids=False)  # the chain kernel's own counts, to subtract
    for v in sys.argv[2:]:
        for kv in v.split(","):
            os.environ[kv.split("=")[0]] = kv.split("=")[1]
        _, _, st = sc.render(cfg, traversal=1, want_ids=False)
        _, _, st = sc.render(cfg, traversal=1, want_ids=False)
        rays = st["shadow_rays"]
        for k in ("primary_rays", "bounce_rays", "reference_rays", "reference_shadow_rays"):
            st[k] -= b[k]
        f = st["reference_shadow_rays"]
        packets, full, phase2 = st["bounce_rays"], 0, f
        print(json.dumps({"variant": v, "vis_ms": round(st["stage_ms"][1], 3), "shadow_rays": rays,
                          "hull_visits_per_ray": round(st["box_tests"] / rays, 3), "ray_visits_per_ray": round(st["tri_tests"] / rays, 3),
                          "packet_tri_tests_per_ray": round(st["primary_rays"] / rays, 3), 
                          "packets": packets, "deferred_per_packet": round(st["reference_rays"] / max(packets, 1), 3),
                          "phase2_frac": round(phase2 / max(packets, 1), 4), "full_walk_frac": round(full / max(packets, 1), 4)}), flush=True)
