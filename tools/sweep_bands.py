"""End-to-end frame time (host pinned output, wall clock) against the number of bands of cge_render (CGE_BANDS), each
variant checked bit-identical to the un-banded frame.  usage: sweep_bands.py [cfg] [bands...]"""
import importlib, json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("computer-graphics-engine_b200")
cfg = pkg.configs.get(sys.argv[1] if len(sys.argv) > 1 else "c5_dragon")
bands = [int(b) for b in sys.argv[2:]] or [1, 2, 4, 8]  # 0 = the library's own choice
W, H = cfg["width"], cfg["height"]
pinned = pkg.PinnedBuffer((H, W, 3), np.float32)
pinned8 = pkg.PinnedBuffer((H, W, 4), np.uint8)
import torch
dev_frame = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
with pkg.Scene(pkg.load_scene(cfg)) as sc:
    base = base8 = None
    for b in bands:
        os.environ.pop("CGE_BANDS", None)
        if b:
            os.environ["CGE_BANDS"] = str(b)
        out = {}
        for label, fn, buf in (("f32", lambda: sc.render(cfg, traversal=1, want_ids=False, rgb_out=pinned.array), pinned),
                               ("rgba8", lambda: sc.render_rgba8(cfg, out=pinned8.array), pinned8),
                               ("device", lambda: (sc.render_device(cfg, dev_frame.data_ptr()),), None)):
            for _ in range(3):
                fn()
            t = []
            for _ in range(8):
                t0 = time.perf_counter()
                r = fn()
                t.append(time.perf_counter() - t0)
            st = r[-1]
            out[label] = {"wall_ms_min": round(1e3 * min(t), 3), "wall_ms_median": round(1e3 * sorted(t)[len(t) // 2], 3),
                          "kernel_ms": round(st["kernel_ms"], 3), "total_ms": round(st["total_ms"], 3)}
        if base is None:
            base, base8 = pinned.array.tobytes(), pinned8.array.tobytes()
        out["identical"] = pinned.array.tobytes() == base and pinned8.array.tobytes() == base8 \
            and dev_frame.cpu().numpy().tobytes() == base
        print(json.dumps({"cfg": cfg["name"], "bands": b, **out}), flush=True)
