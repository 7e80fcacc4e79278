"""Small end-to-end run for compute-sanitizer: every kernel variant on reduced frames, the GPU tree builder, the extras
(multiple rays per pixel, bloom), the RGBA8 output stage and concurrent bands (CGE_BANDS forces them on small frames)."""
import importlib, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
W = pkg.FLAG_WAVEFRONT
for name, (w, h) in {"c1_cornell": (64, 64), "c2_cube_textured": (64, 36), "c3_teapot_soft": (48, 28), "c4_monkey_mirror": (48, 28)}.items():
    cfg = pkg.configs.get(name, w, h)
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        for trav, fl in [(0, 4), (1, 0), (1, W), (1, pkg.FLAG_PER_THREAD)]:
            sc.render(cfg, traversal=trav, flags=fl)
            sc.render(cfg, traversal=trav, flags=fl, part=(1, 3))
            sc.render(cfg, traversal=trav, flags=fl | pkg.FLAG_PARTITION_TILE_ROWS, part=(1, 3))
        C = pkg.configs
        for extra, kw in ((C.FEAT_MULTIPLE_RAYS_PER_PIXEL, {"rays_per_pixel_side": 2}), (C.FEAT_BLOOM_EFFECT, {}),
                          (C.FEAT_BLOOM_EFFECT | C.FEAT_MULTIPLE_RAYS_PER_PIXEL, {"rays_per_pixel_side": 3})):
            sc.render(dict(cfg, features=cfg["features"] | extra, **kw), traversal=1)
        os.environ["CGE_BANDS"] = "3"
        sc.render(cfg, traversal=1)
        sc.render(cfg, traversal=1, flags=W)
        sc.render_rgba8(cfg)
        del os.environ["CGE_BANDS"]
        sc.render_rgba8(cfg)
    t = pkg.build_fast_bvh(pkg.load_scene(cfg), on_gpu=True)
    print(name, "ok", len(t["nodes"]), "fast-tree nodes", flush=True)
flat = pkg.standin.make("dragon", n=24)
cfg = pkg.configs.get("c5_dragon", 48, 28)
with pkg.Scene(flat) as sc:
    for fl in (0, pkg.FLAG_PER_THREAD):
        sc.render(cfg, flags=fl)
print("dragon ok")
