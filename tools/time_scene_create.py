"""Wall clock of scene creation (host BVH in reference order + GPU SAH build + uploads).  usage: time_scene_create.py [cfg]
CGE_TIMING=1 prints the GPU builder's phases; CGE_SAH_BUILD=host times the host builder instead."""
import importlib, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("computer-graphics-engine_b200")
cfg = pkg.configs.get(sys.argv[1] if len(sys.argv) > 1 else "c5_dragon")
flat = pkg.load_scene(cfg)
for it in range(3):
    t0 = time.perf_counter()
    sc = pkg.Scene(flat)
    t1 = time.perf_counter()
    sc.close()
    print(f"cge_scene_create #{it}: {1e3 * (t1 - t0):.1f} ms ({flat.n_primitives} primitives)", flush=True)
