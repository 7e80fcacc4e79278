"""Small end-to-end run for compute-sanitizer: every kernel variant on reduced frames."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
W = pkg.FLAG_WAVEFRONT
for name, (w, h) in {"c1_cornell": (64, 64), "c2_cube_textured": (64, 36), "c3_teapot_soft": (48, 28), "c4_monkey_mirror": (48, 28)}.items():
    cfg = pkg.configs.get(name, w, h)
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        for trav, fl in [(0, 4), (1, 0), (1, W), (1, W | pkg.FLAG_COUPLED_SHADE), (1, W | pkg.FLAG_DECOUPLED_SHADE), (1, W | pkg.FLAG_AUTO_SHADE),
                         (1, pkg.FLAG_PER_THREAD), (1, pkg.FLAG_COOPERATIVE)]:
            sc.render(cfg, traversal=trav, flags=fl)
            sc.render(cfg, traversal=trav, flags=fl, part=(1, 3))
    print(name, "ok", flush=True)
flat = pkg.standin.make("dragon", n=24)
cfg = pkg.configs.get("c5_dragon", 48, 28)
with pkg.Scene(flat) as sc:
    for fl in (0, W | pkg.FLAG_DECOUPLED_SHADE, pkg.FLAG_PER_THREAD, pkg.FLAG_COOPERATIVE):
        sc.render(cfg, flags=fl)
print("dragon ok")
