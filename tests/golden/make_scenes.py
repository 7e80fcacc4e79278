"""Generate tests/golden/scenes/*.cges from the reference checkout (BUILD CONTAINER ONLY).

Runs the reference engine's own loaders (tinyobj, vertex de-duplication, centerAndScaleToUnitMesh, stb_image —
reference framework/src/mesh.cpp:52-176, framework/src/image.cpp:12-35, src/scene.cpp:5-92) through
oracle/_ref/libcge_ref.so and serialises the resulting ``Scene`` so that the GPU box (which has no reference
checkout) sees bit-identical inputs.  C3/C4 are compositions that the reference does not ship; their recipe is
exactly what this script does.

    python tests/golden/make_scenes.py          (requires /root/reference and oracle/_ref built)
"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refharness  # noqa: E402

pkg = importlib.import_module("computer-graphics-engine_b200.scenefile")
OUT = ROOT / "tests" / "golden" / "scenes"
DATA = Path("/root/reference/data")

# SceneType enum values, reference src/scene.h:15-26
SINGLE_TRIANGLE, CUBE, CUBE_TEXTURED, CORNELL, CORNELL_PARALLELOGRAM, MONKEY, TEAPOT, DRAGON, SPHERES, CUSTOM = range(10)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    tmp = OUT / "_tmp.cges"
    data_dir = str(DATA) + "/"

    # --- reference scenes, verbatim -----------------------------------------------------------------
    for name, st in (("triangle", SINGLE_TRIANGLE), ("cube", CUBE), ("cube_textured", CUBE_TEXTURED),
                     ("cornell", CORNELL), ("cornell_parallelogram", CORNELL_PARALLELOGRAM),
                     ("monkey", MONKEY), ("teapot", TEAPOT), ("spheres", SPHERES)):
        refharness.export_prebuilt(st, data_dir, OUT / f"{name}.cges")
        s = pkg.load(OUT / f"{name}.cges")
        print(f"{name}: meshes={len(s.meshes)} verts={len(s.vertices)} tris={len(s.triangles)} "
              f"spheres={len(s.spheres)} lights={len(s.lights)} textures={len(s.textures)}")

    # --- fixtures of tests/test_scene_io.py only (the mirrored loaders): the textured quad of the Custom scene (a material without Kd)
    #     and a quad mesh through loadMesh(file, centerAndNormalize = true) ------------------------------------------------------------
    (OUT.parent / "loader").mkdir(exist_ok=True)
    refharness.export_prebuilt(CUSTOM, data_dir, OUT.parent / "loader" / "custom.cges")
    refharness.export_obj(DATA / "monkey-rotated-quad.obj", True, OUT.parent / "loader" / "monkey_quad.cges")

    # --- C3: teapot + one parallelogram light (area light above / in front of the pot) ---------------
    s = pkg.load(OUT / "teapot.cges")
    s.set_lights([pkg.parallelogram_light(
        v0=(-0.9, 1.3, -1.1), edge01=(0.5, 0.0, 0.0), edge02=(0.0, 0.0, 0.5),
        c0=(1.0, 1.0, 1.0), c1=(1.0, 0.9, 0.8), c2=(0.8, 0.9, 1.0), c3=(1.0, 1.0, 1.0))])
    pkg.save(s, OUT / "teapot_area.cges")

    # --- C4: Cornell box with mirror walls + monkey x0.3 inside ----------------------------------------
    s = pkg.load(OUT / "cornell.cges")
    s.meshes["ks"][:] = np.float32(0.9)        # every wall / block becomes a mirror
    refharness.export_obj(DATA / "monkey.obj", True, tmp)
    monkey = pkg.load(tmp)
    monkey.meshes["ks"][:] = np.float32(0.3)
    s.append_meshes(monkey, scale=0.3, translate=(0.0, -0.05, 0.0))
    pkg.save(s, OUT / "monkey_mirror.cges")
    print(f"monkey_mirror: tris={len(s.triangles)}")

    # --- small mixed scene for unit tests: triangles + spheres + all three light kinds -----------------
    s = pkg.load(OUT / "cube.cges")
    sp = pkg.load(OUT / "spheres.cges")
    sph = sp.spheres.copy()
    sph["center"] = np.array([[1.2, 0.3, 0.4], [-1.3, 0.2, -0.2], [0.1, 1.4, 0.3]], np.float32)
    sph["radius"] = np.array([0.4, 0.5, 0.3], np.float32)
    sph["ks"][0] = np.float32(0.5)
    s.spheres = sph
    s.meshes["transparency"][:] = np.float32(1.0)  # cube.mtl has d 0.452632 -> non-terminating recursion (Q3)
    s.set_lights([
        pkg.point_light((-1.0, 2.0, -1.5), (0.8, 0.8, 0.8)),
        pkg.segment_light((1.5, 1.5, -0.6), (-1.0, 1.5, -0.5), (0.9, 0.2, 0.1), (0.2, 1.0, 0.3)),
        pkg.parallelogram_light((-0.2, 2.0, 0.0), (0.4, 0, 0), (0, 0, 0.4), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)),
    ])
    pkg.save(s, OUT / "mixed.cges")

    tmp.unlink(missing_ok=True)


if __name__ == "__main__":
    main()
