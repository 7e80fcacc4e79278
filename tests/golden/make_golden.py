"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcge_ref.so): reduced-size renders of
the five configs (float RGB + primary-hit primitive ids + ray/box/triangle counters), KAT vectors for the six
libIntersect functions, and the reference-built BVH of each fixture scene.

    python tests/golden/make_golden.py            (build container; needs oracle/_ref built)
    python tests/golden/make_golden.py extras     (only extras.npz: the ExtraFeatures cases of tests/extras_cases.py)
"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refharness  # noqa: E402
import katgen  # noqa: E402

pkg = importlib.import_module("computer-graphics-engine_b200")
OUT = ROOT / "tests" / "golden"

# reduced sizes: same aspect ratio as the full configs so the camera constants are the full-size ones
SMALL = {
    "c1_cornell": (160, 160),
    "c2_cube_textured": (160, 90),
    "c3_teapot_soft": (128, 72),
    "c4_monkey_mirror": (128, 72),
    "c5_dragon": (96, 54),
}
STANDIN_N = 40  # reduced dragon stand-in for the committed golden (12*40^2+2 = 19 202 triangles)


def scene_file_for(cfg, tmpdir):
    if cfg["scene"].startswith("standin:"):
        p = Path(tmpdir) / "standin.cges"
        pkg.scenefile.save(pkg.standin.make("dragon", n=STANDIN_N), p)
        return p
    return pkg.configs.scene_path(cfg)


def extras():
    """tests/golden/extras.npz: the two implemented ExtraFeatures rendered by the unmodified reference (hash-seeded
    std::mt19937 via the wrapped std::random_device), its Gaussian weights and its getRaySamples for a few pixels."""
    import extras_cases
    out = {"weights_sigma1": refharness.weights_gaussian(1.0)}
    for key in extras_cases.CASES:
        cfg = extras_cases.cfg_for(key)
        with refharness.RefScene(pkg.configs.scene_path(cfg), cfg["features"]) as rs:
            rgb, _, _ = rs.render(cfg, threads=0, want_ids=False)
        out[key + "_rgb"] = rgb
        print(key, cfg["width"], cfg["height"], hex(cfg["features"]), "max", float(np.nanmax(rgb)))
    for n in (1, 2, 3, 10):
        cfg = pkg.configs.get("c1_cornell", 64, 48)
        cfg.update(rays_per_pixel_side=n, seed=7 + n)
        out[f"ray_samples_n{n}"] = np.stack([refharness.ray_samples(cfg, x, y) for x, y in ((0, 0), (10, 20), (63, 47))])
    np.savez_compressed(OUT / "extras.npz", **out)
    print("wrote extras.npz")


def main():
    import tempfile
    tmp = tempfile.mkdtemp()
    for name, (w, h) in SMALL.items():
        cfg = pkg.configs.get(name, w, h)
        path = scene_file_for(cfg, tmp)
        with refharness.RefScene(path, cfg["features"]) as rs:
            rgb, ids, st = rs.render(cfg, threads=0, want_ids=True)
            info = rs.bvh_info()
        np.savez_compressed(OUT / f"{name}_{w}x{h}.npz", rgb=rgb, ids=ids,
                            rays=st["rays"], box_tests=st["box_tests"], tri_tests=st["tri_tests"],
                            bvh_nodes=info["nodes"], bvh_levels=info["levels"], bvh_leaves=info["leaves"])
        print(name, w, h, "rays", st["rays"], "box", st["box_tests"], "tri", st["tri_tests"],
              "nan px", int(np.isnan(rgb).any(-1).sum()), "hit", float((ids >= 0).mean()))

    # reference-built BVH for every fixture scene (pins the host BVH builder)
    bv = {}
    for f in sorted((OUT / "scenes").glob("*.cges")):
        with refharness.RefScene(f, pkg.configs.FEAT_ACCEL_STRUCTURE) as rs:
            outp = Path(tmp) / "with_bvh.cges"
            rs.write_with_bvh(outp)
        s = pkg.scenefile.load(outp)
        bv[f.stem + "_nodes"] = s.bvh_nodes
        bv[f.stem + "_order"] = s.bvh_prim_order
        bv[f.stem + "_root"] = np.uint32(s.bvh_root)
    np.savez_compressed(OUT / "reference_bvh.npz", **bv)

    # KAT vectors for I1-I6 + shading helpers, answered by the prebuilt archive / reference sources
    np.savez_compressed(OUT / "kat_vectors.npz", **katgen.make_golden(refharness, n=4000, seed=7))
    print("wrote goldens")


if __name__ == "__main__":
    if sys.argv[1:] == ["extras"]:
        extras()
    else:
        main()
        extras()
