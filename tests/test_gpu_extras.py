"""-m gpu: the two implemented ExtraFeatures through the C ABI against the reference.

 * multiple rays per pixel: against tests/golden/extras.npz (rendered by the unmodified reference with the hash-seeded
   std::mt19937) within the RGB tolerance, FAST and REFERENCE-order traversal, and against the live reference;
 * bloom: (a) bit-exact against the restated renderBloomFilter applied to the GPU's own un-bloomed frame, (b) against the
   reference goldens within tolerance, away from pixels whose brightness sits on the threshold (a 1-ulp difference in the
   un-bloomed value flips the threshold test there: the filter is discontinuous by construction).
"""
import numpy as np
import pytest

import extras_cases
import oracleport
from conftest import GOLDEN, compare_images

pytestmark = pytest.mark.gpu
RGB_TOL = 1e-3
G = np.load(GOLDEN / "extras.npz")


def fragile_mask(base_rgb, threshold):
    """Pixels within reach (3x3) of a pixel whose brightness is within 1e-4 of the bloom threshold, or NaN."""
    b = base_rgb.astype(np.float64)
    br = 0.2126 * b[..., 0] + 0.7152 * b[..., 1] + 0.0722 * b[..., 2]
    near = ~(np.abs(br - threshold) > 1e-4)
    out = near.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            out |= np.roll(np.roll(near, dy, 0), dx, 1)
    return out


@pytest.mark.parametrize("traversal", [0, 1])
@pytest.mark.parametrize("key", list(extras_cases.CASES))
def test_extras_against_reference_golden(cge, key, traversal):
    cfg = extras_cases.cfg_for(key)
    want = G[key + "_rgb"]
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, _, _ = sc.render(cfg, traversal=traversal, want_ids=False)
        keep = np.ones(rgb.shape[:2], bool)
        if cfg["features"] & extras_cases.BLOOM:
            base, _, _ = sc.render(dict(cfg, features=cfg["features"] & ~extras_cases.BLOOM), traversal=traversal, want_ids=False)
            keep = ~fragile_mask(base, cfg.get("bloom_threshold", 0.4))
            assert keep.mean() > 0.9
            # the filter itself: bit-exact against the restatement on the same input
            mine = oracleport.bloom(base, cfg.get("bloom_scalar", 0.3), cfg.get("bloom_threshold", 0.4), cfg.get("bloom_debug_option", 0))
            assert mine.tobytes() == rgb.tobytes()
    err, nan_mm = compare_images(rgb[keep], want[keep])
    scale = max(1.0, float(np.nan_to_num(np.abs(want), nan=0.0, posinf=0.0).max()))
    assert nan_mm == 0 and err <= RGB_TOL * scale, (key, err)


def test_aa_full_size_live_reference(cge, ref):
    """Config C2 at a quarter of its size with 3x3 rays per pixel against the live reference, plus determinism and the
    equality of FAST and literal traversal."""
    cfg = cge.configs.get("c2_cube_textured", 480, 270)
    cfg["features"] |= extras_cases.AA
    cfg.update(rays_per_pixel_side=3, seed=123)
    with ref.RefScene(cge.configs.scene_path(cfg), cfg["features"]) as rs:
        want, _, _ = rs.render(cfg, want_ids=False)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, ids, st = sc.render(cfg, traversal=1)
        rgb2, _, _ = sc.render(cfg, traversal=1)
        rgb_r, ids_r, st_r = sc.render(cfg, traversal=0)
    assert rgb.tobytes() == rgb2.tobytes()
    assert st["primary_rays"] == 9 * 480 * 270 == st_r["primary_rays"]
    err, nan_mm = compare_images(rgb, want)
    assert nan_mm == 0 and err <= RGB_TOL
    err, nan_mm = compare_images(rgb_r, want)
    assert nan_mm == 0 and err <= RGB_TOL
    assert (ids != ids_r).mean() <= 1e-4


def test_aa_soft_shadow_draw_counter(cge, ref):
    """Soft shadows + anti-aliasing: the rand() draw index keeps counting across a pixel's camera rays (the reference's
    call order), so a wrong counter shows up as different jitter, i.e. a different image."""
    cfg = cge.configs.get("c3_teapot_soft", 160, 90)
    cfg["features"] |= extras_cases.AA | cge.configs.FEAT_RECURSIVE
    cfg.update(rays_per_pixel_side=2, ray_depth=2)
    with ref.RefScene(cge.configs.scene_path(cfg), cfg["features"]) as rs:
        want, _, _ = rs.render(cfg, want_ids=False)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb, _, _ = sc.render(cfg, traversal=1, want_ids=False)
    err, nan_mm = compare_images(rgb, want)
    assert nan_mm == 0 and err <= RGB_TOL * max(1.0, float(np.nanmax(want)))


def test_extras_argument_errors(cge):
    cfg = cge.configs.get("c1_cornell", 32, 32)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        bad = dict(cfg, features=cfg["features"] | extras_cases.AA, rays_per_pixel_side=11)
        with pytest.raises(cge.CgeError) as e:
            sc.render(bad)
        assert e.value.code == cge.ERR_INVALID_ARG
        with pytest.raises(cge.CgeError) as e:  # bloom needs the whole frame
            sc.render(dict(cfg, features=cfg["features"] | extras_cases.BLOOM), part=(0, 2))
        assert e.value.code == cge.ERR_UNSUPPORTED
        for bit in (16, 17, 18, 20, 21, 23, 24, 25):  # every other ExtraFeatures flag is refused
            with pytest.raises(cge.CgeError) as e:
                sc.render(dict(cfg, features=cfg["features"] | (1 << bit)))
            assert e.value.code == cge.ERR_UNSUPPORTED


def test_aa_wavefront_equals_per_thread_kernel(cge, monkeypatch):
    """With area lights a multi-sample frame goes through the wavefront pipeline (every camera ray a chain of its own in the
    queues, draw counters offset by the pixel's earlier rays, wf_resolve_kernel adding them in the reference's order): the
    frame must equal the per-thread kernel's bit for bit, whole, in bands and in partitions."""
    cfg = cge.configs.get("c3_teapot_soft", 480, 270)
    cfg["features"] |= extras_cases.AA | cge.configs.FEAT_RECURSIVE
    cfg.update(rays_per_pixel_side=3, ray_depth=2, seed=9)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_t, ids_t, st_t = sc.render(cfg, traversal=1, flags=cge.FLAG_PER_THREAD)
        rgb_w, ids_w, st_w = sc.render(cfg, traversal=1)
        assert st_w["stage_ms"][1] > 0  # the wavefront pipeline ran
        assert rgb_w.tobytes() == rgb_t.tobytes() and np.array_equal(ids_w, ids_t)
        for k in ("primary_rays", "bounce_rays", "shadow_rays", "reference_rays"):
            assert st_w[k] == st_t[k], k
        monkeypatch.setenv("CGE_BANDS", "3")
        rgb_b, _, _ = sc.render(cfg, traversal=1)
        monkeypatch.delenv("CGE_BANDS")
        assert rgb_b.tobytes() == rgb_t.tobytes()
        rgb_p = np.zeros_like(rgb_t)
        for k in range(3):
            sc.render(cfg, traversal=1, rgb_out=rgb_p, want_ids=False, part=(k, 3))
        assert rgb_p.tobytes() == rgb_t.tobytes()


def test_aa_ten_by_ten_in_the_wavefront(cge):
    """The GUI maximum (10 x 10 camera rays per pixel = 200 of the 227 usable MT19937 outputs) through the wavefront pipeline."""
    cfg = cge.configs.get("c3_teapot_soft", 64, 36)
    cfg["features"] |= extras_cases.AA
    cfg.update(rays_per_pixel_side=10, seed=4)
    with cge.Scene(cge.load_scene(cfg)) as sc:
        rgb_t, _, st_t = sc.render(cfg, traversal=1, flags=cge.FLAG_PER_THREAD)
        rgb_w, _, st_w = sc.render(cfg, traversal=1)
        rgb_r, _, _ = sc.render(cfg, traversal=0)
    assert st_w["stage_ms"][1] > 0 and st_w["primary_rays"] == 100 * 64 * 36
    assert rgb_w.tobytes() == rgb_t.tobytes()
    err, nan_mm = compare_images(rgb_w, rgb_r)
    assert nan_mm == 0 and err <= RGB_TOL
