"""development aid: SHA-256 of the frames of some configs (to compare library builds bit for bit).  usage: frame_hash.py cfg[:scale] ..."""
import hashlib, importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
out = []
for spec in sys.argv[1:]:
    name, scale = (spec.split(":") + ["1.0"])[:2]
    full = pkg.configs.get(name)
    cfg = pkg.configs.get(name, int(full["width"] * float(scale)), int(full["height"] * float(scale)))
    with pkg.Scene(pkg.load_scene(cfg)) as sc:
        rgb, ids, _ = sc.render(cfg)
    out.append(f"{spec}={hashlib.sha256(rgb.tobytes() + ids.tobytes()).hexdigest()[:12]}")
print("frames:", " ".join(out))
