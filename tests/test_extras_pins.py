"""CPU tests (-m "not gpu") for the two implemented ExtraFeatures: multiple rays per pixel
(reference src/render.cpp:211-227,295-303) and the bloom filter (:158-210).

Pins: tests/golden/extras.npz, rendered by the UNMODIFIED reference (oracle/_ref) whose std::random_device value is replaced
by a hash of (seed, pixel) — see oracle/ref/ref_api.cpp; generator tests/golden/make_golden.py extras."""
import importlib

import numpy as np
import pytest

import extras_cases
import oracleport
from conftest import GOLDEN, compare_images

pkg = importlib.import_module("computer-graphics-engine_b200")
G = np.load(GOLDEN / "extras.npz")


@pytest.mark.parametrize("key", list(extras_cases.CASES))
def test_oracle_restatement_equals_reference_golden(key):
    if not oracleport.available():
        pytest.skip("oracle/liboracle.so not built")
    cfg = extras_cases.cfg_for(key)
    with oracleport.OracleScene(pkg.configs.scene_path(cfg), cfg["features"]) as sc:
        rgb, _, _ = sc.render(cfg, want_ids=False)
    assert rgb.tobytes() == G[key + "_rgb"].tobytes()


def test_bloom_weights_equal_reference():
    """The library's host-side weightsGaussian against the reference's own (double exp, float accumulation)."""
    assert pkg.bloom_weights(1.0).tobytes() == G["weights_sigma1"].tobytes()


def test_bloom_weights_live(ref):
    for sigma in (0.5, 1.0, 2.5):
        assert pkg.bloom_weights(sigma).tobytes() == ref.weights_gaussian(sigma).tobytes()


@pytest.mark.parametrize("n", [1, 2, 3, 10])
def test_ray_sample_positions_reproduce_reference_rays(n, ref):
    """The library's MT19937 head + generate_canonical + jitter arithmetic (csrc/sampler.h, host evaluation) yields the NDC
    positions from which the reference's getRaySamples produced its rays: pushing them through the camera restatement must
    give the golden ray directions bit for bit."""
    cfg = pkg.configs.get("c1_cornell", 64, 48)
    cfg.update(rays_per_pixel_side=n, seed=7 + n)
    want = G[f"ray_samples_n{n}"]
    for k, (x, y) in enumerate(((0, 0), (10, 20), (63, 47))):
        ndc = pkg.ray_sample_positions(64, 48, x, y, n, cfg["seed"])
        rays = ref.generate_rays(cfg, ndc)  # the reference's own Trackball::generateRay
        assert rays[:, :6].tobytes() == want[k].tobytes()


def test_mt19937_head_equals_numpy_mt19937():
    """Independent check of the two-recurrence generator: numpy's MT19937 with init_genrand seeding."""
    rng = np.random.MT19937()
    # the library exposes the stream through cge_ray_sample_positions: with W = H = 2 and pixel (0, 0) the pixel corner is -1
    # and the stratum size 1/n, so position = (-1 + i/n) + u/n; the same float32 expression is evaluated here from numpy's
    # MT19937 seeded (init_genrand) with the library's per-pixel seed hash(seed ^ 'RDEV', pixel, 0)
    n = 10
    got = pkg.ray_sample_positions(2, 2, 0, 0, n, 0)
    h = 0 ^ 0x52444556
    def hash_sample(seed, pixel, ctr):
        M = 0xFFFFFFFF
        v = (seed ^ (pixel * 0x9E3779B1 & M)) & M
        v ^= (ctr * 0x85EBCA77) & M
        v ^= v >> 16; v = v * 0x85EBCA6B & M; v ^= v >> 13; v = v * 0xC2B2AE35 & M; v ^= v >> 16
        return v >> 1
    rng._legacy_seeding(hash_sample(int(h), 0, 0))
    raw = rng.random_raw(2 * n * n)
    u = np.minimum(raw.astype(np.float32) / np.float32(4294967296.0), np.float32(0.99999994)).astype(np.float32)
    box = np.float32(np.float32(np.float32(1) / np.float32(2) * np.float32(2)) / np.float32(n))
    k = 0
    for i in range(n):
        for j in range(n):
            jy, jx = np.float32(u[k] * box), np.float32(u[k + 1] * box)
            ex = np.float32(np.float32(np.float32(-1) + np.float32(np.float32(i) * box)) + jx)
            ey = np.float32(np.float32(np.float32(-1) + np.float32(np.float32(j) * box)) + jy)
            assert got[k // 2, 0] == ex and got[k // 2, 1] == ey
            k += 2


def test_extras_validation_needs_no_gpu():
    """Unsupported ExtraFeatures are refused before any CUDA work (status code, not a silent render)."""
    # cge_ray_sample_positions argument checking
    with pytest.raises(pkg.CgeError):
        pkg.ray_sample_positions(8, 8, 0, 0, 11, 0)
    with pytest.raises(pkg.CgeError):
        pkg.ray_sample_positions(8, 8, 8, 0, 2, 0)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_oracle_restatement_equals_live_reference_on_random_extras(ref, seed):
    """Random combinations of the two extra features, their globals, depth and frame size: the restatement must equal the
    unmodified reference bit for bit (same libm, same arithmetic order, same hash-seeded generator)."""
    if not oracleport.available():
        pytest.skip("oracle/liboracle.so not built")
    rng = np.random.default_rng(seed)
    C = pkg.configs
    name = ["c1_cornell", "c2_cube_textured", "c3_teapot_soft", "c4_monkey_mirror"][int(rng.integers(4))]
    w, h = int(rng.integers(17, 49)), int(rng.integers(9, 33))
    cfg = C.get(name, w, h)
    extra = [extras_cases.AA, extras_cases.BLOOM, extras_cases.AA | extras_cases.BLOOM][int(rng.integers(3))]
    cfg["features"] |= extra
    cfg.update(rays_per_pixel_side=int(rng.integers(1, 5)), bloom_scalar=float(rng.uniform(0.05, 0.9)),
               bloom_threshold=float(rng.uniform(0.0, 0.8)), bloom_debug_option=int(rng.integers(3)),
               ray_depth=int(rng.integers(0, 4)), seed=int(rng.integers(1 << 30)))
    path = C.scene_path(cfg)
    with ref.RefScene(path, cfg["features"]) as rs:
        want, _, _ = rs.render(cfg, want_ids=False)
    with oracleport.OracleScene(path, cfg["features"]) as sc:
        got, _, _ = sc.render(cfg, want_ids=False)
    assert got.tobytes() == want.tobytes(), (name, w, h, hex(extra), cfg["rays_per_pixel_side"], cfg["bloom_debug_option"])
