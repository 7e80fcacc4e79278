"""-m gpu: the six libIntersect functions evaluated ON THE DEVICE (through the C ABI) against the reference's
prebuilt archive — committed golden answers, plus fresh fuzz answered live by oracle/_ref when it is present.
Bar: bit-exact return values and ray.t (all NaNs equal)."""
import numpy as np
import pytest

import katgen

pytestmark = pytest.mark.gpu


def check(cge, c, a):
    hit, t = cge.kat_triangle(c["tri_v"], c["tri_ray"])
    assert np.array_equal(hit, a["tri_hit"]) and katgen.bits_equal(t, a["tri_t"]).all()
    hit, t = cge.kat_triangle(c["tri_v"], c["tri_ray"], precomputed=True)
    assert np.array_equal(hit, a["tri_hit"]) and katgen.bits_equal(t, a["tri_t"]).all()
    hit, t = cge.kat_aabb(c["box_b"], c["box_ray"])
    assert np.array_equal(hit, a["box_hit"]) and katgen.bits_equal(t, a["box_t"]).all()
    hit, t, n = cge.kat_sphere(c["sph_s"], c["sph_ray"])
    assert np.array_equal(hit, a["sph_hit"]) and katgen.bits_equal(t, a["sph_t"]).all()
    assert katgen.bits_equal(n[hit == 1], a["sph_n"][hit == 1]).all()
    hit, t = cge.kat_plane(c["pl_p"], c["pl_ray"])
    assert np.array_equal(hit, a["pl_hit"]) and katgen.bits_equal(t, a["pl_t"]).all()
    assert katgen.bits_equal(cge.kat_triangle_plane(c["tri_v"]), a["tp"]).all()
    assert np.array_equal(cge.kat_point_in_triangle(c["pit_v"], c["pit_n"], c["pit_p"]), a["pit"])


def test_golden_kat_vectors(cge):
    g = np.load(cge.configs.SCENE_DIR.parent / "kat_vectors.npz")
    check(cge, g, {k[4:]: g[k] for k in g.files if k.startswith("ans_")})


def test_live_fuzz_against_reference_archive(cge, ref):
    for seed in (101, 202):
        c = katgen.make_cases(200000, seed)
        check(cge, c, katgen.answers(ref, c))
