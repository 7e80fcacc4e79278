import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
for name in ("c1_cornell", "c4_monkey_mirror", "c3_teapot_soft"):
    cfg = pkg.configs.get(name, 256, 144)
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        a, ia, sa = sc.render(cfg, flags=pkg.FLAG_PER_THREAD | pkg.FLAG_COUNT_TESTS)   # trace_fast<true,true>
        b, ib, sb = sc.render(cfg, flags=pkg.FLAG_PER_THREAD)                          # trace_shadow
        r, ir, sr = sc.render(cfg, traversal=0)
    d = (np.nan_to_num(a) != np.nan_to_num(b)).any(-1)
    d2 = (np.abs(np.nan_to_num(a) - np.nan_to_num(r)) > 1e-3).any(-1)
    print(name, "pixels differing count-vs-shadow:", int(d.sum()), "count-vs-reference-traversal:", int(d2.sum()), "of", d.size, "kernel_ms", sa["kernel_ms"], sb["kernel_ms"])
    if d.sum():
        ys, xs = np.nonzero(d)
        print("  first diffs at", list(zip(ys[:5], xs[:5])), a[ys[0], xs[0]], b[ys[0], xs[0]])
