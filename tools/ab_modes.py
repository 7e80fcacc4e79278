"""A/B: kernel ms of selected modes for selected configs (development aid). usage: ab_modes.py cfg[:scale] ... """
import importlib, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
W = pkg.FLAG_WAVEFRONT
MODES = {"default": 0, "wave": W, "coupled": W | pkg.FLAG_COUPLED_SHADE, "grouped": W | pkg.FLAG_GROUPED_SHADE, "decoupled": W | pkg.FLAG_DECOUPLED_SHADE, "thread": pkg.FLAG_PER_THREAD}
for spec in sys.argv[1:]:
    name, scale = (spec.split(":") + ["1.0"])[:2]
    full = pkg.configs.get(name)
    cfg = pkg.configs.get(name, int(full["width"] * float(scale)), int(full["height"] * float(scale)))
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        out = {}
        for label, fl in MODES.items():
            best = None
            for _ in range(4):
                _, _, st = sc.render(cfg, traversal=1, want_ids=False, flags=fl)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            out[label] = round(best["kernel_ms"], 3)
            if label in ("wave", "grouped"):
                out[label + "_stages"] = [round(x, 3) for x in best["stage_ms"]]
        print(name, cfg["width"], json.dumps(out), flush=True)
