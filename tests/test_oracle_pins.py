"""CPU tests (-m "not gpu"): the oracle (oracle/cge_oracle.cpp, the CPU restatement) pinned against the reference.

Pins: committed goldens produced by the UNMODIFIED reference (tests/golden/*.npz, generator tests/golden/make_golden.py)
and, where oracle/_ref is present, the reference itself run live."""
import importlib

import numpy as np
import pytest

import katgen
import oracleport
from conftest import GOLDEN, compare_images

pkg = importlib.import_module("computer-graphics-engine_b200")

if not oracleport.available():
    pytest.skip("oracle/liboracle.so not built", allow_module_level=True)

SMALL = {"c1_cornell": (160, 160), "c2_cube_textured": (160, 90), "c3_teapot_soft": (128, 72),
         "c4_monkey_mirror": (128, 72), "c5_dragon": (96, 54)}


def check_kats(c, a):
    hit, t = oracleport.kat_triangle(c["tri_v"], c["tri_ray"])
    assert np.array_equal(hit, a["tri_hit"]) and katgen.bits_equal(t, a["tri_t"]).all()
    hit, t = oracleport.kat_aabb(c["box_b"], c["box_ray"])
    assert np.array_equal(hit, a["box_hit"]) and katgen.bits_equal(t, a["box_t"]).all()
    hit, t, n = oracleport.kat_sphere(c["sph_s"], c["sph_ray"])
    assert np.array_equal(hit, a["sph_hit"]) and katgen.bits_equal(t, a["sph_t"]).all()
    assert katgen.bits_equal(n[hit == 1], a["sph_n"][hit == 1]).all()
    hit, t = oracleport.kat_plane(c["pl_p"], c["pl_ray"])
    assert np.array_equal(hit, a["pl_hit"]) and katgen.bits_equal(t, a["pl_t"]).all()
    assert katgen.bits_equal(oracleport.kat_triangle_plane(c["tri_v"]), a["tp"]).all()
    assert np.array_equal(oracleport.kat_point_in_triangle(c["pit_v"], c["pit_n"], c["pit_p"]), a["pit"])


def test_intersect_kats_golden():
    g = np.load(GOLDEN / "kat_vectors.npz")
    check_kats(g, {k[4:]: g[k] for k in g.files if k.startswith("ans_")})
    assert katgen.bits_equal(oracleport.kat_barycentric(g["pit_v"], g["pit_p"]), g["ans_bary"]).all()
    assert katgen.bits_equal(oracleport.kat_reflection(g["refl_in"]), g["ans_refl"]).all()
    # computeShading goes through powf; same libm here, so bit-exact too
    assert katgen.bits_equal(oracleport.kat_shading(g["shade_in"]), g["ans_shade"]).all()


def test_intersect_kats_live_fuzz(ref):
    c = katgen.make_cases(300000, 4242)
    check_kats(c, katgen.answers(ref, c))


def test_camera_constants(ref):
    for name in pkg.configs.CONFIGS:
        cfg = pkg.configs.get(name)
        assert bytes(oracleport.camera(cfg)) == bytes(ref.camera(cfg)), name


def test_bvh_equals_reference_tree():
    g = np.load(GOLDEN / "reference_bvh.npz")
    for f in sorted(pkg.configs.SCENE_DIR.glob("*.cges")):
        flat = pkg.scenefile.load(f)
        with oracleport.OracleScene(f) as sc:
            nodes, order, root = sc.bvh_export(flat.n_primitives)
        assert np.array_equal(order, g[f.stem + "_order"]), f.stem
        assert nodes.tobytes() == g[f.stem + "_nodes"].tobytes(), f.stem
        assert root == int(g[f.stem + "_root"])


@pytest.mark.parametrize("name", list(SMALL))
def test_render_equals_reference_golden(name, tmp_path):
    w, h = SMALL[name]
    cfg = pkg.configs.get(name, w, h)
    g = np.load(GOLDEN / f"{name}_{w}x{h}.npz")
    if cfg["scene"].startswith("standin:"):
        path = tmp_path / "standin.cges"
        pkg.scenefile.save(pkg.standin.make("dragon", n=40), path)
    else:
        path = pkg.configs.scene_path(cfg)
    with oracleport.OracleScene(path) as sc:
        rgb, ids, st = sc.render(cfg)
        info = sc.bvh_info()
    assert np.array_equal(ids, g["ids"])
    err, nan_mm = compare_images(rgb, g["rgb"])
    assert nan_mm == 0 and err <= 1e-6 * max(1.0, float(np.nan_to_num(g["rgb"], nan=0).max()))
    # the restatement performs exactly the reference's intersect calls
    assert (st["rays"], st["box_tests"], st["tri_tests"]) == (int(g["rays"]), int(g["box_tests"]), int(g["tri_tests"]))
    assert (info["nodes"], info["levels"], info["leaves"]) == (int(g["bvh_nodes"]), int(g["bvh_levels"]), int(g["bvh_leaves"]))


def test_render_live_reference_mixed_scene(ref):
    C = pkg.configs
    path = C.SCENE_DIR / "mixed.cges"
    base = {"scene": "mixed.cges", "width": 64, "height": 48, "ray_depth": 2, "segment_samples": 4,
            "parallelogram_samples": 3, "seed": 11,
            "camera": {"fov_deg": 60.0, "dist": 4.0, "look_at": [0.0, 0.3, 0.0], "rotation_deg": [15.0, 35.0, 0.0]}}
    for feats in (C.FEAT_SHADING, C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_HARD_SHADOW | C.FEAT_SOFT_SHADOW | C.FEAT_RECURSIVE,
                  C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_NORMAL_INTERP | C.FEAT_TEXTURE_MAPPING | C.FEAT_RECURSIVE,
                  C.FEAT_ACCEL_STRUCTURE):
        cfg = dict(base, features=feats)
        with ref.RefScene(path, feats) as rs:
            r_rgb, r_ids, r_st = rs.render(cfg)
        with oracleport.OracleScene(path) as sc:
            rgb, ids, st = sc.render(cfg)
        assert np.array_equal(ids, r_ids)
        assert rgb.tobytes() == r_rgb.tobytes() or compare_images(rgb, r_rgb) == (0.0, 0)
        assert (st["rays"], st["box_tests"], st["tri_tests"], st["sphere_tests"]) == (
            r_st["rays"], r_st["box_tests"], r_st["tri_tests"], r_st["sphere_tests"])


def test_reference_harness_loop_equals_renderRayTracing(ref):
    """The harness's own pixel loop (needed because the reference hard-codes depth 5, src/render.cpp:318) is the
    reference's renderRayTracing for depth 5: identical images."""
    cfg = pkg.configs.get("c1_cornell", 96, 96)
    cfg["ray_depth"] = 5
    with ref.RefScene(pkg.configs.scene_path(cfg), cfg["features"]) as rs:
        a, _, _ = rs.render(cfg, want_ids=False)
        b, _, _ = rs.render(cfg, want_ids=False, use_render_ray_tracing=True)
    assert a.tobytes() == b.tobytes()
