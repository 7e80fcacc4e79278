"""Per-pixel cost map of a config (CGE_DEV_FLAG_DEBUG_CYCLES): where does the frame time go?"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
name = sys.argv[1]
cfg = pkg.configs.get(name)
if len(sys.argv) > 2:
    cfg["features"] = int(sys.argv[2], 0)
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    sc.render(cfg)
    rgb, cost, st = sc.render(cfg, flags=pkg.FLAG_DEBUG_CYCLES | pkg.FLAG_PER_THREAD)
    _, nb, stc = sc.render(cfg, flags=pkg.FLAG_DEBUG_CYCLES | pkg.FLAG_COUNT_TESTS | pkg.FLAG_PER_THREAD)
    _, ids, _ = sc.render(cfg)
print("fast-tree box tests / ray", stc["box_tests"] / stc["gpu_rays"], "tri tests / ray", stc["tri_tests"] / stc["gpu_rays"])
nbt = nb[: nb.shape[0] // 4 * 4, : nb.shape[1] // 8 * 8].reshape(nb.shape[0] // 4, 4, nb.shape[1] // 8, 8)
print("per-pixel box tests: mean", nb.mean(), "max", nb.max(), "p99", np.percentile(nb, 99))
np.save("gpurun_out/nbox_%s.npy" % name, nbt.max(axis=(1, 3)).astype(np.float32))
cost = cost.astype(np.float64) * 16
H, W = cost.shape
tiles = cost[: H // 4 * 4, : W // 8 * 8].reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))  # a tile costs its slowest lane
print("kernel_ms", st["kernel_ms"], "sum tile cycles / (148*32 warps) at 1.9GHz = ms:", tiles.sum() / (148 * 32) / 1.9e6)
q = np.percentile(tiles, [50, 90, 99, 99.9, 100])
print("tile cycles percentiles 50/90/99/99.9/max:", q, "max tile ms:", q[-1] / 1.9e6)
np.save("gpurun_out/cost_%s.npy" % name, tiles.astype(np.float32))
rows = tiles.sum(1)
print("cost by tile-row decile (top of image first):", [round(float(x), 1) for x in (np.add.reduceat(rows, np.linspace(0, len(rows), 11)[:-1].astype(int)) / rows.sum() * 100)])
hit = (ids >= 0)
print("hit frac", hit.mean())
ti = np.argsort(tiles.ravel())[::-1][:8]
for i in ti:
    r, c = np.unravel_index(i, tiles.shape)
    print("hot tile row", r, "col", c, "cycles", tiles[r, c], "max box tests of a pixel in it", nb[r * 4:(r + 1) * 4, c * 8:(c + 1) * 8].max())
