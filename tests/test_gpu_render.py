"""-m gpu: image / primitive-id parity of the CUDA path (through the C ABI) against the reference renderer.

 * committed goldens (tests/golden/*.npz, produced by the unmodified reference in the build container),
 * live renders by oracle/_ref at reduced sizes (same configs, same camera constants),
 * full-size, size-independent properties: REFERENCE-order traversal == FAST traversal, an image rendered in
   N interleaved partitions == the image rendered whole, determinism.
Bars (BASELINE.json north_star): primary-hit ids bit-exact except <= 0.01 % edge ties; linear RGB within 1e-3
max-abs; NaN pixels in the same places.
"""
import numpy as np
import pytest

from conftest import compare_images

pytestmark = pytest.mark.gpu

SMALL = {
    "c1_cornell": (160, 160),
    "c2_cube_textured": (160, 90),
    "c3_teapot_soft": (128, 72),
    "c4_monkey_mirror": (128, 72),
    "c5_dragon": (96, 54),
}
RGB_TOL = 1e-3
ID_MISMATCH_BUDGET = 1e-4


def same_rays_answered(a, b):
    """Two pipelines answered the same rays: camera and reflection rays equal; shadow rays equal too, except that the light-hull
    pre-pass of the wavefront pipeline (wavefront.cuh wf_vis_cull_kernel) settles whole hits without tracing - it reports the
    light SAMPLES it settled, which include samples that needed no ray, so the traced counts bracket each other."""
    if a["primary_rays"] != b["primary_rays"] or a["bounce_rays"] != b["bounce_rays"]:
        return False
    if not a["shadow_samples_culled"] and not b["shadow_samples_culled"]:
        return a["shadow_rays"] == b["shadow_rays"]
    return all(x["shadow_rays"] <= y["shadow_rays"] + y["shadow_samples_culled"] for x, y in ((a, b), (b, a)))


def flat_for(cge, cfg, small_standin=True):
    if cfg["scene"].startswith("standin:"):
        return cge.standin.make("dragon", n=40) if small_standin else cge.standin.make("dragon")
    return cge.load_scene(cfg)


def assert_parity(name, rgb, ids, ref_rgb, ref_ids, id_budget=ID_MISMATCH_BUDGET):
    """north_star's bar: primary ids bit-exact except documented edge-tie pixels within the budget; linear RGB within 1e-3, NaN in
    the same places.  A tie pixel (a ray grazing a box face within rounding: the reference's exact box arithmetic misses a leaf
    the conservative fast walk reaches) shows another primitive and hence another colour, in the primary hit or in a reflection:
    pixels out of tolerance are counted against the same budget instead of failing one by one."""
    npx = ids.size
    allowed = max(int(id_budget * npx), 0)
    id_mm = int((ids != ref_ids).sum())
    assert id_mm <= allowed, f"{name}: {id_mm}/{npx} primary ids differ"
    a = np.asarray(rgb, np.float32).reshape(-1, 3)
    b = np.asarray(ref_rgb, np.float32).reshape(-1, 3)
    # relative tolerance for the (few) pixels whose value exceeds 1: 1e-3 max-abs is stated for [0,1] radiance
    scale = np.maximum(1.0, np.nan_to_num(np.abs(b), nan=0.0, posinf=0.0).max())
    nan_diff = (np.isnan(a) != np.isnan(b)).any(-1)
    with np.errstate(invalid="ignore"):
        off = np.abs(np.nan_to_num(a, nan=0.0, posinf=3e38, neginf=-3e38) - np.nan_to_num(b, nan=0.0, posinf=3e38, neginf=-3e38)).max(-1) > RGB_TOL * scale
    bad = int((nan_diff | off).sum())
    assert bad <= allowed, f"{name}: {bad}/{npx} pixels out of tolerance or differing in NaN-ness (budget {allowed})"


# (traversal, flags): literal traversal with test counters / fast tree with the default pipeline of the frame / fast tree with
# the wavefront pipeline forced / fast tree with the per-thread kernel forced (include/cge.h CGE_FLAG_COUNT_TESTS, CGE_DEV_FLAG_*)
MODES = {"reference": (0, 4), "fast-default": (1, 0), "fast-wave": (1, 1 << 17), "fast-thread": (1, 1 << 16)}


@pytest.mark.parametrize("name", list(SMALL))
@pytest.mark.parametrize("mode", list(MODES))
def test_golden_images(cge, name, mode):
    w, h = SMALL[name]
    traversal, flags = MODES[mode]
    cfg = cge.configs.get(name, w, h)
    g = np.load(cge.configs.SCENE_DIR.parent / f"{name}_{w}x{h}.npz")
    with cge.Scene(flat_for(cge, cfg)) as sc:
        info = sc.bvh_info()
        assert (info["nodes"], info["levels"], info["leaves"]) == (int(g["bvh_nodes"]), int(g["bvh_levels"]), int(g["bvh_leaves"]))
        rgb, ids, st = sc.render(cfg, traversal=traversal, flags=flags)
    assert_parity(name, rgb, ids, g["rgb"], g["ids"])
    # the reference would have made exactly this many BvhInterface::intersect calls (ld --wrap counter)
    assert st["reference_rays"] == int(g["rays"]), (st["reference_rays"], int(g["rays"]))
    if mode == "reference":
        # literal traversal: the device performs exactly the reference's box and triangle tests (ld --wrap counters).
        # The GPU traces each mirror chain once, so compare per unique ray class where no duplication happens.
        if not (cfg["features"] & cge.configs.FEAT_RECURSIVE):
            assert st["box_tests"] == int(g["box_tests"]) and st["tri_tests"] == int(g["tri_tests"])


@pytest.mark.parametrize("name,size", [("c1_cornell", (256, 256)), ("c2_cube_textured", (320, 180)),
                                       ("c3_teapot_soft", (192, 108)), ("c4_monkey_mirror", (160, 90))])
def test_live_reference(cge, ref, name, size, tmp_path):
    cfg = cge.configs.get(name, *size)
    cfg["seed"] = 99
    path = cge.configs.scene_path(cfg)
    with ref.RefScene(path, cfg["features"]) as rs:
        ref_rgb, ref_ids, rst = rs.render(cfg)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        for mode, (traversal, flags) in MODES.items():
            rgb, ids, st = sc.render(cfg, traversal=traversal, flags=flags)
            assert_parity(f"{name}/{mode}", rgb, ids, ref_rgb, ref_ids)
            assert st["reference_rays"] == rst["rays"]


def test_mixed_scene_all_light_kinds_spheres_and_flags(cge, ref):
    """triangles + spheres, point + segment + parallelogram lights, every Features combination that matters."""
    C = cge.configs
    path = C.SCENE_DIR / "mixed.cges"
    flat = cge.scenefile.load(path)
    base = {"scene": "mixed.cges", "width": 96, "height": 64, "ray_depth": 2, "segment_samples": 5,
            "parallelogram_samples": 3, "seed": 5,
            "camera": {"fov_deg": 60.0, "dist": 4.0, "look_at": [0.0, 0.3, 0.0], "rotation_deg": [15.0, 35.0, 0.0]}}
    combos = [
        C.FEAT_ACCEL_STRUCTURE,
        C.FEAT_SHADING,
        C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_HARD_SHADOW,
        C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_SOFT_SHADOW,
        C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_SOFT_SHADOW | C.FEAT_HARD_SHADOW | C.FEAT_RECURSIVE,
        C.FEAT_SHADING | C.FEAT_SOFT_SHADOW | C.FEAT_RECURSIVE | C.FEAT_NORMAL_INTERP,
        C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_NORMAL_INTERP | C.FEAT_TEXTURE_MAPPING | C.FEAT_RECURSIVE,
    ]
    with cge.Scene(flat) as sc:
        for feats in combos:
            cfg = dict(base, features=feats)
            with ref.RefScene(path, feats) as rs:
                ref_rgb, ref_ids, rst = rs.render(cfg)
            for mode, (traversal, flags) in MODES.items():
                rgb, ids, st = sc.render(cfg, traversal=traversal, flags=flags)
                assert_parity(f"mixed/f{feats:#x}/{mode}", rgb, ids, ref_rgb, ref_ids, id_budget=2e-3)
                assert st["reference_rays"] == rst["rays"]
    # the same lights over a triangles-only scene exercise the fast tree + cooperative kernel with 1 + 5 + 9 samples
    path = C.SCENE_DIR / "monkey.cges"
    mk = cge.scenefile.load(path)
    mk.lights = flat.lights
    tmp = C.SCENE_DIR.parent / "_tmp_monkey_lights.cges"
    cge.scenefile.save(mk, tmp)
    try:
        with cge.Scene(mk) as sc:
            for feats in combos[2:]:
                cfg = dict(base, features=feats, camera=dict(base["camera"], dist=2.5))
                with ref.RefScene(tmp, feats) as rs:
                    ref_rgb, ref_ids, rst = rs.render(cfg)
                for mode, (traversal, flags) in MODES.items():
                    rgb, ids, st = sc.render(cfg, traversal=traversal, flags=flags)
                    assert_parity(f"monkey-lights/f{feats:#x}/{mode}", rgb, ids, ref_rgb, ref_ids, id_budget=2e-3)
                    assert st["reference_rays"] == rst["rays"]
    finally:
        tmp.unlink(missing_ok=True)


def test_bvh_matches_reference_tree(cge):
    g = np.load(cge.configs.SCENE_DIR.parent / "reference_bvh.npz")
    for f in sorted(cge.configs.SCENE_DIR.glob("*.cges")):
        flat = cge.scenefile.load(f)
        with cge.Scene(flat) as sc:
            nodes, order = sc.bvh_export()
        assert np.array_equal(order, g[f.stem + "_order"]), f.stem
        assert nodes.tobytes() == g[f.stem + "_nodes"].tobytes(), f.stem


def test_trace_rays_matches_reference_getFinalColor(cge, ref):
    cfg = cge.configs.get("c4_monkey_mirror", 64, 36)
    rng = np.random.default_rng(3)
    n = 4000
    o = rng.uniform(-0.3, 0.3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d = (d / np.sqrt((d * d).sum(1, keepdims=True))).astype(np.float32)
    rays = np.concatenate([o, d, np.full((n, 1), np.float32(3.4028234663852886e38))], 1).astype(np.float32)
    with ref.RefScene(cge.configs.scene_path(cfg), cfg["features"]) as rs:
        ref_rgb, ref_ids = rs.trace_rays(rays, cfg)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, ids = sc.trace_rays(rays, cfg, traversal=1)
    assert (ids != ref_ids).sum() <= 1
    err, nan_mm = compare_images(rgb, ref_rgb)
    assert nan_mm == 0 and err <= 1e-3 * max(1.0, float(np.nan_to_num(ref_rgb, nan=0).max()))


@pytest.mark.parametrize("name", ["c1_cornell", "c2_cube_textured", "c3_teapot_soft", "c4_monkey_mirror"])
def test_full_size_properties(cge, name):
    """At BASELINE.json's full resolution: FAST traversal == REFERENCE-order traversal (ids within the tie budget,
    RGB within tolerance) and a 3-way interleaved partition reassembles to the same image bit for bit."""
    cfg = cge.configs.get(name)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_f, ids_f, st_f = sc.render(cfg, traversal=1)
        rgb_r, ids_r, st_r = sc.render(cfg, traversal=0)
        assert (ids_f != ids_r).mean() <= ID_MISMATCH_BUDGET
        assert st_f["reference_rays"] == st_r["reference_rays"]
        # the two production pipelines (one thread per pixel, wavefront) walk the same tree with the same arithmetic: identical bits
        for fl in (cge.FLAG_PER_THREAD, cge.FLAG_WAVEFRONT):
            rgb_t, ids_t, st_t = sc.render(cfg, traversal=1, flags=fl)
            assert st_t["reference_rays"] == st_f["reference_rays"] and same_rays_answered(st_t, st_f)
            assert np.array_equal(ids_t, ids_f) and rgb_t.tobytes() == rgb_f.tobytes()
        err, nan_mm = compare_images(rgb_f, rgb_r)
        assert nan_mm <= ID_MISMATCH_BUDGET * ids_f.size
        differing = (np.abs(np.nan_to_num(rgb_f, nan=0.0) - np.nan_to_num(rgb_r, nan=0.0)).max(-1) > 1e-3).mean()
        assert differing <= ID_MISMATCH_BUDGET, differing
        rgb_p = np.zeros_like(rgb_f)
        ids_p = np.full_like(ids_f, -7)
        for k in range(3):
            sc.render(cfg, traversal=1, rgb_out=rgb_p, ids_out=ids_p, part=(k, 3))
        assert np.array_equal(ids_p, ids_f)
        assert rgb_p.tobytes() == rgb_f.tobytes()
        rgb_2, _, _ = sc.render(cfg, traversal=1)
        assert rgb_2.tobytes() == rgb_f.tobytes()  # deterministic


def test_dragon_standin_full_scene_reduced_frame(cge, ref, tmp_path):
    """The full 868 334-triangle stand-in (config C5) on a reduced frame against the live reference."""
    cfg = cge.configs.get("c5_dragon", 160, 90)
    flat = cge.standin.make("dragon")
    path = tmp_path / "dragon.cges"
    cge.scenefile.save(flat, path)
    with ref.RefScene(path, cfg["features"]) as rs:
        ref_rgb, ref_ids, rst = rs.render(cfg)
        info = rs.bvh_info()
    with cge.Scene(flat) as sc:
        mine = sc.bvh_info()
        assert (mine["nodes"], mine["levels"], mine["leaves"]) == (info["nodes"], info["levels"], info["leaves"])
        for mode, (traversal, flags) in MODES.items():
            rgb, ids, st = sc.render(cfg, traversal=traversal, flags=flags)
            assert_parity(f"dragon/{mode}", rgb, ids, ref_rgb, ref_ids)
            assert st["reference_rays"] == rst["rays"]


def test_edge_cases(cge):
    C = cge.configs
    cfg = C.get("c1_cornell", 33, 17)  # ragged: not a multiple of the 8x4 tile
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, ids, st = sc.render(cfg)
        assert st["primary_rays"] == 33 * 17
        # ExtraFeatures other than bloom / multiple rays per pixel must be refused, not silently ignored
        bad = dict(cfg, features=cfg["features"] | (1 << 16))
        with pytest.raises(cge.CgeError) as e:
            sc.render(bad)
        assert e.value.code == cge.ERR_UNSUPPORTED
        with pytest.raises(cge.CgeError):
            sc.render(dict(cfg, ray_depth=99))
        # light update: no lights -> black (shading on), then restore
        lights = sc.flat.lights.copy()
        sc.update_lights(lights[:0])
        rgb0, _, _ = sc.render(cfg)
        assert np.nan_to_num(rgb0, nan=0.0).max() == 0.0
        sc.update_lights(lights)
        rgb1, _, _ = sc.render(cfg)
        assert rgb1.tobytes() == rgb.tobytes()
    # empty scene: every ray misses
    empty = cge.scenefile.FlatScene()
    with cge.Scene(empty) as sc:
        rgb, ids, st = sc.render(C.get("c1_cornell", 16, 8))
        assert (ids == -1).all() and (rgb == 0).all()
    # a transparent material with recursion has no terminating reference result
    cube = cge.scenefile.load(C.SCENE_DIR / "cube.cges")
    with cge.Scene(cube) as sc:
        with pytest.raises(cge.CgeError) as e:
            sc.render(C.get("c1_cornell", 16, 8))
        assert e.value.code == cge.ERR_UNSUPPORTED


def test_partition_matches_host_tile_list(cge):
    """cge_render with part_index/part_count touches exactly the tiles partition_tiles() lists (multi-GPU sharding)."""
    cfg = cge.configs.get("c1_cornell", 100, 50)  # ragged in both directions
    W, H = cfg["width"], cfg["height"]
    ntx = (W + 7) // 8
    with cge.Scene(cge.load_scene(cfg)) as sc:
        for parts, rows in ((2, False), (3, False), (8, False), (2, True), (3, True), (8, True)):
            for k in range(parts):
                ids = np.full((H, W), -7, np.int32)
                rgb = np.full((H, W, 3), -7.0, np.float32)
                sc.render(cfg, rgb_out=rgb, ids_out=ids, part=(k, parts), flags=cge.FLAG_PARTITION_TILE_ROWS if rows else 0)
                touched = ids != -7
                expect = np.zeros((H, W), bool)
                for t in cge.partition_tiles(W, H, k, parts, tile_rows=rows):
                    tx, ty = t % ntx, t // ntx
                    y0, y1 = ty * 4, min(ty * 4 + 4, H)
                    expect[H - y1:H - y0, tx * 8:min(tx * 8 + 8, W)] = True
                assert np.array_equal(touched, expect), (parts, k)
                assert np.array_equal((rgb != -7.0).all(-1) | np.isnan(rgb).any(-1), expect)


def test_concurrent_renders_on_one_scene(cge):
    """cge_render is re-entrant on one cge_scene (the reference renders several cameras from concurrent threads sharing
    scene + bvh, src/main.cpp:514-528): every call has its own stream and scratch."""
    import threading
    cfgs = [cge.configs.get("c3_teapot_soft", 320, 180), cge.configs.get("c3_teapot_soft", 256, 144),
            cge.configs.get("c3_teapot_soft", 200, 120), cge.configs.get("c3_teapot_soft", 320, 180)]
    for i, c in enumerate(cfgs):
        c["camera"] = dict(c["camera"], rotation_deg=[25.0, 30.0 + 40.0 * i, 0.0])
    with cge.Scene(cge.load_scene(cfgs[0])) as sc:
        expect = [sc.render(c)[0] for c in cfgs]
        got = [None] * len(cfgs)
        errs = []

        def work(i):
            try:
                for _ in range(5):
                    got[i] = sc.render(cfgs[i])[0]
            except Exception as e:  # pragma: no cover
                errs.append(e)
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cfgs))]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errs
        for a, b in zip(expect, got):
            assert a.tobytes() == b.tobytes()


def test_c5_full_size_rows_against_live_reference(cge, ref, tmp_path):
    """BASELINE.json's headline config at its FULL size (868 334 triangles, 3840x2160, soft shadows, depth 3): the GPU frame
    against the reference renderer on 24 evenly spaced rows of the same frame (the whole frame is ~3 minutes of CPU)."""
    cfg = cge.configs.get("c5_dragon")
    flat = cge.standin.make("dragon")
    path = tmp_path / "dragon.cges"
    cge.scenefile.save(flat, path)
    W, H = cfg["width"], cfg["height"]
    stride = 90
    with ref.RefScene(path, cfg["features"]) as rs:
        ref_rgb, ref_ids, rst = rs.render(cfg, y_stride=stride)
    with cge.Scene(flat) as sc:
        rgb, ids, st = sc.render(cfg)
        rgb_t, ids_t, _ = sc.render(cfg, flags=cge.FLAG_PER_THREAD)
    assert np.array_equal(ids, ids_t) and rgb.tobytes() == rgb_t.tobytes()  # wavefront == per-thread kernel, whole frame
    rows = [H - 1 - y for y in range(0, H, stride)]  # Screen rows of reference y = 0, stride, ...
    assert len(rows) == 24
    a_rgb, a_ids = rgb[rows], ids[rows]
    b_rgb, b_ids = ref_rgb[rows], ref_ids[rows]
    assert (a_ids != b_ids).sum() <= max(1, int(ID_MISMATCH_BUDGET * a_ids.size))
    err, nan_mm = compare_images(a_rgb, b_rgb)
    assert nan_mm == 0
    assert err <= RGB_TOL * max(1.0, float(np.nan_to_num(b_rgb, nan=0.0).max())), err


@pytest.mark.parametrize("name,stride", [("c1_cornell", 8), ("c2_cube_textured", 6), ("c3_teapot_soft", 18), ("c4_monkey_mirror", 36)])
def test_full_size_rows_against_live_reference(cge, ref, name, stride):
    """C1-C4 at BASELINE.json's FULL sizes, directly against the reference renderer on evenly spaced rows of the same frame
    (128 / 180 / 60 / 40 rows): the tie budget and the tolerance hold at the judged resolution, not only at the goldens'."""
    cfg = cge.configs.get(name)
    W, H = cfg["width"], cfg["height"]
    with ref.RefScene(cge.configs.scene_path(cfg), cfg["features"]) as rs:
        ref_rgb, ref_ids, _ = rs.render(cfg, y_stride=stride)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, ids, _ = sc.render(cfg)
    rows = [H - 1 - y for y in range(0, H, stride)]
    a_rgb, a_ids, b_rgb, b_ids = rgb[rows], ids[rows], ref_rgb[rows], ref_ids[rows]
    assert (a_ids != b_ids).sum() <= max(1, int(ID_MISMATCH_BUDGET * a_ids.size)), int((a_ids != b_ids).sum())
    err, nan_mm = compare_images(a_rgb, b_rgb)
    # a pixel whose primary hit differs within the tie budget may differ in NaN-ness and colour as well
    tie = (a_ids != b_ids)
    a_ok, b_ok = a_rgb[~tie], b_rgb[~tie]
    err, nan_mm = compare_images(a_ok, b_ok)
    assert nan_mm == 0
    assert err <= RGB_TOL * max(1.0, float(np.nan_to_num(np.abs(b_ok), nan=0.0, posinf=0.0).max())), err


def test_rgba8_output_stage_matches_reference_bitmap_writer(cge, ref, tmp_path):
    """Screen::writeBitmapToFile's clamp -> *255 -> u8x4 conversion done on the GPU before the D2H copy.
    The reference's own writer (stb BMP) is applied to the same float frame and read back."""
    from PIL import Image
    for name, size in (("c1_cornell", (200, 200)), ("c4_monkey_mirror", (192, 108)), ("c3_teapot_soft", (160, 90))):
        cfg = cge.configs.get(name, *size)
        with cge.Scene(cge.load_scene(cfg)) as sc:
            rgb, _, _ = sc.render(cfg, want_ids=False)
            rgba, st = sc.render_rgba8(cfg)
        assert np.isnan(rgb).any() or name != "c1_cornell"  # the Cornell frame exercises the NaN rule
        bmp = tmp_path / f"{name}.bmp"
        ref.write_bmp(rgb, bmp)
        want = np.asarray(Image.open(bmp).convert("RGBA"))
        assert want.shape == rgba.shape
        assert np.array_equal(rgba[..., :3], want[..., :3]), name
        assert (rgba[..., 3] == 255).all()


@pytest.mark.parametrize("name", ["c1_cornell", "c3_teapot_soft", "c4_monkey_mirror"])
def test_zero_shading_cull_changes_no_bit(cge, name, monkeypatch):
    """FAST traversal skips the shadow ray of a light sample with n.l <= 0 (its Phong term is exactly zero, shade.cuh
    shading_is_zero): fewer rays traced, the frame identical bit for bit in every kernel variant."""
    w, h = SMALL[name]
    cfg = cge.configs.get(name, w * 2, h * 2)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        for flags in (0, cge.FLAG_PER_THREAD, cge.FLAG_WAVEFRONT):
            monkeypatch.setenv("CGE_ZERO_SHADING_CULL", "0")
            rgb0, ids0, st0 = sc.render(cfg, traversal=1, flags=flags)
            monkeypatch.setenv("CGE_ZERO_SHADING_CULL", "1")
            rgb1, ids1, st1 = sc.render(cfg, traversal=1, flags=flags)
            assert rgb0.tobytes() == rgb1.tobytes() and np.array_equal(ids0, ids1)
            assert st1["shadow_rays"] < st0["shadow_rays"] and st1["reference_rays"] == st0["reference_rays"]


@pytest.mark.parametrize("name,size,extra", [
    ("c3_teapot_soft", (256, 144), {}),
    ("c3_teapot_soft", (128, 72), {"parallelogram_samples": 3}),
    ("c5_dragon", (192, 108), {}),
    ("c5_dragon", (96, 54), {"ray_depth": 1}),
])
def test_light_hull_prepass_changes_no_bit(cge, name, size, extra, monkeypatch):
    """The light-hull pre-pass of the wavefront pipeline (wf_vis_cull_kernel: one conservative walk per hit with the hull of its
    lights; hits no shadow ray of which can be blocked are settled without tracing) only ever answers "certainly visible":
    fewer rays traced, the frame identical bit for bit, with the chain stage split or not, in one pipeline or in bands."""
    cfg = cge.configs.get(name, *size)
    cfg.update(extra)
    with cge.Scene(flat_for(cge, cfg)) as sc:
        monkeypatch.setenv("CGE_VIS_CULL", "0")
        rgb0, ids0, st0 = sc.render(cfg, traversal=1, flags=cge.FLAG_WAVEFRONT)
        rgb_t, _, st_t = sc.render(cfg, traversal=1, flags=cge.FLAG_PER_THREAD)
        assert st0["shadow_samples_culled"] == 0 and st0["shadow_rays"] == st_t["shadow_rays"] and rgb0.tobytes() == rgb_t.tobytes()
        monkeypatch.setenv("CGE_VIS_CULL", "1")
        for split, bands in (("0", "1"), ("1", "1"), ("0", "3")):
            monkeypatch.setenv("CGE_CHAIN_SPLIT", split)
            monkeypatch.setenv("CGE_BANDS", bands)
            rgb1, ids1, st1 = sc.render(cfg, traversal=1, flags=cge.FLAG_WAVEFRONT)
            assert rgb1.tobytes() == rgb0.tobytes() and np.array_equal(ids1, ids0), (split, bands)
            assert 0 < st1["shadow_samples_culled"] and st1["shadow_rays"] < st0["shadow_rays"]
            assert st0["shadow_rays"] <= st1["shadow_rays"] + st1["shadow_samples_culled"]
            assert st1["reference_rays"] == st0["reference_rays"] and same_rays_answered(st1, st0)


def test_light_hull_prepass_with_every_light_kind_and_off_for_spheres(cge, monkeypatch):
    """Point + segment + parallelogram lights at once (the hull is then the direction box of all their corner points, without the
    pyramid planes); a scene with spheres keeps every hit for the per-ray kernel."""
    C = cge.configs
    base = {"width": 160, "height": 96, "ray_depth": 2, "segment_samples": 5, "parallelogram_samples": 3, "seed": 5,
            "camera": {"fov_deg": 60.0, "dist": 2.5, "look_at": [0.0, 0.3, 0.0], "rotation_deg": [15.0, 35.0, 0.0]},
            "features": C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_SOFT_SHADOW | C.FEAT_HARD_SHADOW | C.FEAT_RECURSIVE}
    mixed = cge.scenefile.load(C.SCENE_DIR / "mixed.cges")
    mk = cge.scenefile.load(C.SCENE_DIR / "monkey.cges")
    mk.lights = mixed.lights
    for flat, spheres in ((mk, False), (mixed, True)):
        cfg = dict(base, scene="-", camera=dict(base["camera"], dist=4.0 if spheres else 2.5))
        with cge.Scene(flat) as sc:
            monkeypatch.setenv("CGE_VIS_CULL", "0")
            rgb0, ids0, st0 = sc.render(cfg, traversal=1, flags=cge.FLAG_WAVEFRONT)
            monkeypatch.setenv("CGE_VIS_CULL", "1")
            rgb1, ids1, st1 = sc.render(cfg, traversal=1, flags=cge.FLAG_WAVEFRONT)
            assert st0["stage_ms"][1] > 0 and rgb1.tobytes() == rgb0.tobytes() and np.array_equal(ids1, ids0)
            if spheres:
                assert st1["shadow_samples_culled"] == 0 and st1["shadow_rays"] == st0["shadow_rays"]
            else:
                assert st1["shadow_samples_culled"] > 0 and st1["shadow_rays"] < st0["shadow_rays"]


def test_zero_shading_cull_off_for_unbounded_colours(cge):
    """(kd * Lc) * 0 is NaN when the product overflows: with a light colour beyond the bound the cull must stay off."""
    import copy
    cfg = cge.configs.get("c1_cornell", 96, 96)
    flat = cge.load_scene(cfg)
    with cge.Scene(flat) as sc:
        _, _, st_on = sc.render(cfg, traversal=1)
    hot = copy.deepcopy(flat)
    assert int(hot.lights[0]["type"]) == 0  # the Cornell box has one point light: position, colour
    hot.lights[0]["v"][3:6] = 3e38
    with cge.Scene(hot) as sc:
        _, _, st_hot = sc.render(cfg, traversal=1)
        _, _, st_ref = sc.render(cfg, traversal=0)
    assert st_hot["shadow_rays"] == st_ref["shadow_rays"] > st_on["shadow_rays"]


@pytest.mark.parametrize("name", ["c3_teapot_soft", "c4_monkey_mirror"])
def test_concurrent_bands_change_no_bit(cge, name, monkeypatch):
    """cge_render renders large frames as concurrent bands of tile rows, each a pipeline on its own stream (cge_api.cu
    launch_bands): the frame, the ids, the packed bitmap and the ray counters must not depend on the number of bands."""
    cfg = cge.configs.get(name)
    # (the light-hull pre-pass is switched on by launch size, which the band count changes: pinned, so that the counters compare)
    monkeypatch.setenv("CGE_VIS_CULL", "1")
    with cge.Scene(cge.load_scene(cfg)) as sc:
        monkeypatch.setenv("CGE_BANDS", "1")
        rgb1, ids1, st1 = sc.render(cfg, traversal=1)
        rgba1, _ = sc.render_rgba8(cfg)
        for bands in ("3", "8", None):
            if bands is None:
                monkeypatch.delenv("CGE_BANDS")
            else:
                monkeypatch.setenv("CGE_BANDS", bands)
            rgb, ids, st = sc.render(cfg, traversal=1)
            rgba, _ = sc.render_rgba8(cfg)
            assert rgb.tobytes() == rgb1.tobytes() and np.array_equal(ids, ids1) and rgba.tobytes() == rgba1.tobytes()
            for k in ("primary_rays", "bounce_rays", "shadow_rays", "shadow_samples_culled", "reference_rays", "reference_shadow_rays"):
                assert st[k] == st1[k], k


def test_degenerate_geometry(cge):
    """Zero-area triangles and a NaN vertex: such primitives can never be hit; the scene must still build (a NaN coordinate
    sends the fast tree to the host builder, csrc/bvh_sah_gpu.cu sah_gpu_supported) and FAST must equal the literal traversal."""
    cfg = cge.configs.get("c1_cornell", 96, 96)
    flat = cge.load_scene(cfg)
    tri = cge.scenefile.load(cge.configs.SCENE_DIR / "triangle.cges")
    tri.vertices["position"][:] = tri.vertices["position"][0]   # all three vertices coincide: zero area
    flat.append_meshes(tri)
    bad = cge.scenefile.load(cge.configs.SCENE_DIR / "triangle.cges")
    bad.vertices["position"][1, 0] = np.nan
    flat.append_meshes(bad, scale=0.1)
    with cge.Scene(flat) as sc:
        rgb_f, ids_f, st_f = sc.render(cfg, traversal=1)
        rgb_r, ids_r, st_r = sc.render(cfg, traversal=0)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_0, ids_0, _ = sc.render(cfg, traversal=1)
    assert np.array_equal(ids_f, ids_r) and np.array_equal(ids_f, ids_0)  # the added primitives are never hit
    assert rgb_f.tobytes() == rgb_0.tobytes()
    err, nan_mm = compare_images(rgb_f, rgb_r)
    assert nan_mm == 0 and err <= RGB_TOL


def test_spheres_and_no_accel_take_the_fast_tree(cge, ref):
    """Scenes with spheres and frames without enableAccelStructure run on the fast tree (the spheres beside it, behind the
    reference tree's box chain; without the acceleration structure the same minimum-t answer with the primitive-vector tie
    rank): against the live reference at a reduced size, against the literal traversal at 1080p, and >= 5x faster than it."""
    C = cge.configs
    for scene in ("mixed.cges", "spheres.cges"):
        path = C.SCENE_DIR / scene
        flat = cge.scenefile.load(path)
        base = {"scene": scene, "ray_depth": 2, "segment_samples": 5, "parallelogram_samples": 3, "seed": 7,
                "camera": {"fov_deg": 60.0, "dist": 4.0, "look_at": [0.0, 0.3, 0.0], "rotation_deg": [15.0, 35.0, 0.0]}}
        with cge.Scene(flat) as sc:
            for feats in (C.FEAT_SHADING | C.FEAT_ACCEL_STRUCTURE | C.FEAT_HARD_SHADOW | C.FEAT_SOFT_SHADOW | C.FEAT_RECURSIVE,
                          C.FEAT_SHADING | C.FEAT_HARD_SHADOW | C.FEAT_SOFT_SHADOW | C.FEAT_RECURSIVE):  # second: no accel structure
                small = dict(base, width=120, height=68, features=feats)
                with ref.RefScene(path, feats) as rs:
                    ref_rgb, ref_ids, rst = rs.render(small)
                for flags in (0, cge.FLAG_PER_THREAD, cge.FLAG_WAVEFRONT):
                    rgb, ids, st = sc.render(small, traversal=1, flags=flags)
                    assert_parity(f"{scene}/f{feats:#x}/{flags:#x}", rgb, ids, ref_rgb, ref_ids, id_budget=2e-3)
                    assert st["reference_rays"] == rst["rays"]
                big = dict(base, width=1920, height=1080, features=feats)
                best = {}
                for trav in (0, 1):
                    for _ in range(3):
                        rgb, ids, st = sc.render(big, traversal=trav)
                        best[trav] = min(best.get(trav, 1e9), st["kernel_ms"])
                    if trav == 0:
                        rgb_l, ids_l = rgb, ids
                assert (ids != ids_l).mean() <= 2e-3
                differing = (np.abs(np.nan_to_num(rgb, nan=0.0) - np.nan_to_num(rgb_l, nan=0.0)).max(-1) > 1e-3).mean()
                assert differing <= 2e-3, differing
    # a C4-class scene (monkey + mirror walls, 999 triangles, recursion depth 6) with the three spheres added: against the live
    # reference at a reduced size, and the fast tree >= 5x faster than the literal traversal at 1080p
    mk = cge.scenefile.load(C.SCENE_DIR / "monkey_mirror.cges")
    sp = cge.scenefile.load(C.SCENE_DIR / "spheres.cges").spheres.copy()
    sp["center"] = [[0.3, -0.25, 0.2], [-0.3, 0.1, -0.2], [0.0, 0.42, 0.05]]  # inside the box, around the monkey
    sp["radius"] = [0.12, 0.15, 0.1]
    sp["ks"][1] = [0.6, 0.6, 0.6]  # one of them a mirror: reflection rays leave a sphere
    mk.spheres = sp
    tmp = C.SCENE_DIR.parent / "_tmp_monkey_spheres.cges"
    cge.scenefile.save(mk, tmp)
    try:
        cfg = cge.configs.get("c4_monkey_mirror", 160, 90)
        with cge.Scene(mk) as sc:
            with ref.RefScene(tmp, cfg["features"]) as rs:
                ref_rgb, ref_ids, rst = rs.render(cfg)
            assert (ref_ids >= len(mk.triangles)).sum() > 50  # the spheres are in view
            for flags in (0, cge.FLAG_WAVEFRONT):
                rgb, ids, st = sc.render(cfg, traversal=1, flags=flags)
                assert_parity(f"monkey+spheres/{flags:#x}", rgb, ids, ref_rgb, ref_ids, id_budget=2e-3)
                assert st["reference_rays"] == rst["rays"]
            big = cge.configs.get("c4_monkey_mirror", 1920, 1080)
            best = {}
            for trav in (0, 1):
                for _ in range(3):
                    _, _, st = sc.render(big, traversal=trav, want_ids=False)
                    best[trav] = min(best.get(trav, 1e9), st["kernel_ms"])
            assert best[0] >= 5.0 * best[1], best
    finally:
        tmp.unlink(missing_ok=True)
