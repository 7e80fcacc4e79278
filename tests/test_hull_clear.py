"""CPU: the light-hull pre-pass of the shadow stage never calls a hit "certainly unoccluded" wrongly.

The pre-pass (csrc/wavefront.cuh wf_vis_cull_kernel) settles a hit without tracing its shadow rays when every triangle its hull
reaches is CLEAR, i.e. when cull_triangle_clear says that no ray from the hit point to any point of the light can be accepted by
that triangle.  cge_hull_clear_host evaluates that very function on the host.  Here its claim is tested against the REFERENCE's own
intersectRayWithTriangle (oracle/_ref, the prebuilt archive) on rays sampled the way the renderer samples the light
(reference src/light.cpp:30-45: (v0 + hw * e01) + vw * e02 in float, corners and edges included): a triangle called clear must not be
hit by any of them - in the configurations that matter (triangles beside the hit point, grazing the hull's sides, in front of the
light, behind the hit point) as well as in random ones.  The other direction is checked loosely: the test must not be vacuous.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

F = np.float32


@pytest.fixture(scope="module")
def libs():
    cge = importlib.import_module("computer-graphics-engine_b200")
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref is not built")
    return cge, refharness


def light_points(rng, n, k, v0, e01, e02):
    """k sample positions per case, float arithmetic as in sample_light: (v0 + hw * e01) + vw * e02."""
    hw = rng.random((n, k)).astype(F)
    vw = rng.random((n, k)).astype(F)
    special = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, 0], [0, 0.5], [1, 0.5], [0.5, 1]], F)
    hw[:, :8], vw[:, :8] = special[:, 0], special[:, 1]
    a = (v0[:, None, :] + hw[..., None] * e01[:, None, :]).astype(F)
    return (a + vw[..., None] * e02[:, None, :]).astype(F)


def cases(rng, n, kind):
    o = (rng.normal(size=(n, 3)) * 0.7).astype(F)
    v0 = (o + rng.normal(size=(n, 3)) * 0.3 + np.array([0.3, 2.5, -0.4])).astype(F)
    e01 = (rng.normal(size=(n, 3)) * 0.05 + np.array([0.6, 0, 0])).astype(F)
    e02 = (rng.normal(size=(n, 3)) * 0.05 + np.array([0, 0, 0.6])).astype(F)
    centre_l = v0 + 0.5 * e01 + 0.5 * e02
    axis = centre_l - o
    size = (10.0 ** rng.uniform(-3.0, -0.3, size=(n, 1))).astype(F)
    if kind == "between":      # somewhere along the pyramid, up to a few widths off its axis
        t = rng.uniform(0.02, 1.1, size=(n, 1))
        c = o + t * axis + rng.normal(size=(n, 3)) * 0.35 * t
    elif kind == "beside_hit":  # the hit point's own surface: triangles within a few sizes of o
        c = o + rng.normal(size=(n, 3)) * size * 1.5
    elif kind == "grazing":     # close to the pyramid's sides
        t = rng.uniform(0.05, 1.0, size=(n, 1))
        side = rng.choice([-1.0, 1.0], size=(n, 1))
        which = rng.integers(0, 2, size=(n, 1))
        off = np.where(which == 0, e01, e02) * side * 0.5 * t * rng.uniform(0.9, 1.15, size=(n, 1))
        c = o + t * axis + off
    elif kind == "behind":
        c = o - rng.uniform(0.01, 1.0, size=(n, 1)) * axis + rng.normal(size=(n, 3)) * 0.1
    else:                        # around and beyond the light
        c = centre_l + rng.normal(size=(n, 3)) * 0.4
    tri = (c[:, None, :] + rng.normal(size=(n, 3, 3)) * size[:, None, :]).astype(F)
    return o, v0, e01, e02, tri.reshape(n, 9)


@pytest.mark.parametrize("kind", ["between", "beside_hit", "grazing", "behind", "at_light"])
def test_a_clear_triangle_is_hit_by_no_ray_to_the_light(libs, kind):
    cge, ref = libs
    rng = np.random.default_rng({"between": 1, "beside_hit": 2, "grazing": 3, "behind": 4, "at_light": 5}[kind])
    n, k = 150000, 40
    o, v0, e01, e02, tri = cases(rng, n, kind)
    clear = cge.hull_clear_host(o, np.concatenate([v0, e01, e02], 1), tri).astype(bool)
    pos = light_points(rng, n, k, v0, e01, e02)
    d = (pos - o[:, None, :]).astype(F)
    ray = np.concatenate([np.broadcast_to(o[:, None, :], d.shape), d, np.ones((n, k, 1), F)], 2).reshape(n * k, 7)
    hit, _ = ref.kat_triangle(np.repeat(tri, k, axis=0), ray)  # accepted with 0 <= t <= 1: what a shadow ray asks
    hit_any = hit.reshape(n, k).any(1)
    wrong = clear & hit_any
    assert not wrong.any(), f"{kind}: {int(wrong.sum())} triangles called clear are hit, first case {int(np.flatnonzero(wrong)[0])}"
    # not vacuous: most triangles no sampled ray hits are recognised as clear (the rest go to the per-ray kernel: cost, not error)
    power = float((clear & ~hit_any).sum()) / max(int((~hit_any).sum()), 1)
    floor = {"between": 0.5, "beside_hit": 0.3, "grazing": 0.2, "behind": 0.8, "at_light": 0.05}[kind]
    assert power >= floor, f"{kind}: only {power:.2f} of the unhit triangles are called clear"
    print(f"{kind}: hit {hit_any.mean():.3f}, clear {clear.mean():.3f}, clear among unhit {power:.3f}")


def test_degenerate_inputs(libs):
    """NaN / zero-area triangles and a hit point inside the light's plane: never an error, never a wrong claim."""
    cge, ref = libs
    o = np.zeros((4, 3), F)
    light = np.tile(np.array([[-0.3, 2.0, -0.3, 0.6, 0, 0, 0, 0, 0.6]], F), (4, 1))
    tri = np.array([[np.nan] * 9, [0, 1, 0] * 3, [-1, 1, -1, 1, 1, -1, 0, 1, 1], [-5, 1, -5, 5, 1, -5, 0, 1, 5]], F)
    o[2] = (0.1, 2.0, 0.1)  # in the light's plane: the pyramid is flat, nothing may be inferred from its sides
    clear = cge.hull_clear_host(o, light, tri)
    assert clear[0] == 1 and clear[3] == 0  # a NaN triangle is never accepted; a big triangle across the pyramid is not clear
    pos = np.array([[-0.3 + 0.6 * a, 2.0, -0.3 + 0.6 * b] for a in (0, 0.5, 1) for b in (0, 0.5, 1)], F)
    for i in range(4):
        ray = np.concatenate([np.tile(o[i], (9, 1)), pos - o[i], np.ones((9, 1), F)], 1)
        hit, _ = ref.kat_triangle(np.tile(tri[i], (9, 1)), ray)
        assert not (clear[i] and hit.any())


def segment_box_overlap(o, d, lo, hi):
    """Does the segment o + t d, t in [0, 1], meet the closed box [lo, hi]?  Exact slab arithmetic in float64."""
    o, d, lo, hi = (np.asarray(a, np.float64) for a in (o, d, lo, hi))
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
    par = d == 0  # parallel to the slab: inside it or never
    inside = (o >= lo) & (o <= hi)
    tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
    return np.maximum(tn.max(-1), 0.0) <= np.minimum(tf.min(-1), 1.0)


@pytest.mark.parametrize("kind", ["between", "beside_hit", "grazing", "behind", "at_light"])
def test_a_box_the_hull_misses_is_met_by_no_ray_to_the_light(libs, kind):
    """The pre-pass walks the tree with the hull of the light: a subtree is skipped when hull_box says no ray of the hull can pass
    through its box.  Checked against exact segment / box arithmetic on sampled rays, for boxes from leaf size to scene size."""
    cge, _ = libs
    rng = np.random.default_rng({"between": 11, "beside_hit": 12, "grazing": 13, "behind": 14, "at_light": 15}[kind])
    n, k = 150000, 40
    o, v0, e01, e02, tri = cases(rng, n, kind)
    pts = tri.reshape(n, 3, 3)
    grow = (10.0 ** rng.uniform(-4.0, 0.3, size=(n, 1))).astype(F)  # from a sliver around the triangle to a box holding the scene
    lo, hi = (pts.min(1) - grow * rng.random((n, 3))).astype(F), (pts.max(1) + grow * rng.random((n, 3))).astype(F)
    hit = cge.hull_box_host(o, np.concatenate([v0, e01, e02], 1), np.concatenate([lo, hi], 1)).astype(bool)
    pos = light_points(rng, n, k, v0, e01, e02)
    d = (pos - o[:, None, :]).astype(F)
    met = segment_box_overlap(o[:, None, :], d, lo[:, None, :], hi[:, None, :]).any(1)
    wrong = ~hit & met
    assert not wrong.any(), f"{kind}: {int(wrong.sum())} boxes called missed are met by a ray, first case {int(np.flatnonzero(wrong)[0])}"
    rejected = float((~hit & ~met).sum()) / max(int((~met).sum()), 1)
    assert rejected >= 0.25, f"{kind}: only {rejected:.2f} of the boxes no sampled ray meets are rejected"
    print(f"{kind}: met {met.mean():.3f}, hull says hit {hit.mean():.3f}, rejected among unmet {rejected:.3f}")
