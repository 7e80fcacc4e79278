"""Fuzz-case generators for the libIntersect functions (I1-I6) and the shading helpers (S1, S4, S5).

The cases follow SURVEY.md Appendix A's recipe: mixed normalised / un-normalised directions, ray.t in {1, FLT_MAX,
random}, zeroed direction components, origins inside boxes, points exactly on triangle edges and vertices,
degenerate triangles, spheres with tangent rays.  Answers come from the reference (prebuilt archive / sources)
through tests/refharness.py; the CUDA path and the CPU oracle must reproduce them bit for bit.
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def rays(rng, n, origin_scale=2.0):
    o = rng.uniform(-origin_scale, origin_scale, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    kind = rng.integers(0, 4, n)
    norm = np.sqrt((d * d).sum(1, keepdims=True)).astype(np.float32)
    d = np.where((kind != 0)[:, None], d / norm, d * np.float32(3.0)).astype(np.float32)
    # zero out direction components in a fraction of the cases (AABB special constants, Appendix A)
    z = rng.random((n, 3)) < 0.08
    d = np.where(z, np.float32(0.0), d).astype(np.float32)
    tsel = rng.integers(0, 3, n)
    t = np.where(tsel == 0, np.float32(1.0), np.where(tsel == 1, FLT_MAX, rng.uniform(0, 6, n).astype(np.float32))).astype(np.float32)
    return np.concatenate([o, d, t[:, None]], 1).astype(np.float32)


def triangles(rng, n):
    v = rng.uniform(-1.5, 1.5, (n, 9)).astype(np.float32)
    # a few degenerate triangles (two equal vertices / collinear)
    deg = rng.random(n) < 0.01
    v[deg, 3:6] = v[deg, 0:3]
    return v


def triangle_cases(rng, n):
    """Rays aimed at random points of random triangles; a third land exactly on an edge or a vertex (in fp32)."""
    v = triangles(rng, n)
    r = rays(rng, n)
    v0, v1, v2 = v[:, 0:3], v[:, 3:6], v[:, 6:9]
    w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
    mode = rng.integers(0, 6, n)
    w[mode == 1, 2] = 0  # on edge v0-v1
    w[mode == 2, 0] = 0  # on edge v1-v2
    w[mode == 3] = np.array([1, 0, 0], np.float32)  # a vertex
    w = (w / np.maximum(w.sum(1, keepdims=True), np.float32(1e-20))).astype(np.float32)
    target = (w[:, 0:1] * v0 + w[:, 1:2] * v1 + w[:, 2:3] * v2).astype(np.float32)
    aimed = mode != 5  # mode 5 keeps a completely random ray
    d = (target - r[:, 0:3]).astype(np.float32)
    unit = rng.random(n) < 0.6
    nrm = np.sqrt((d * d).sum(1, keepdims=True)).astype(np.float32)
    d = np.where(unit[:, None], d / np.maximum(nrm, np.float32(1e-20)), d).astype(np.float32)
    r[aimed, 3:6] = d[aimed]
    # shadow-ray style: un-normalised direction with t = 1 ends exactly at / before / after the triangle
    return v, r.astype(np.float32)


def aabb_cases(rng, n):
    c = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    h = rng.uniform(0.0, 1.0, (n, 3)).astype(np.float32)
    flat = rng.random((n, 3)) < 0.05
    h = np.where(flat, np.float32(0.0), h).astype(np.float32)
    b = np.concatenate([c - h, c + h], 1).astype(np.float32)
    r = rays(rng, n)
    inside = rng.random(n) < 0.2
    r[inside, 0:3] = (c[inside] + (rng.uniform(-0.9, 0.9, (int(inside.sum()), 3)).astype(np.float32) * h[inside])).astype(np.float32)
    onface = rng.random(n) < 0.1  # origin exactly on a slab plane
    r[onface, 0] = b[onface, 0]
    aim = rng.random(n) < 0.5
    tgt = (c + rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32) * h).astype(np.float32)
    r[aim, 3:6] = (tgt[aim] - r[aim, 0:3]).astype(np.float32)
    z = rng.random((n, 3)) < 0.08
    r[:, 3:6] = np.where(z, np.float32(0.0), r[:, 3:6])
    return b, r.astype(np.float32)


def sphere_cases(rng, n):
    s = np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0.05, 1.5, (n, 1))], 1).astype(np.float32)
    r = rays(rng, n)
    aim = rng.random(n) < 0.6
    off = rng.normal(size=(n, 3)).astype(np.float32)
    off = (off / np.sqrt((off * off).sum(1, keepdims=True))).astype(np.float32)
    scale = np.where(rng.random(n) < 0.3, np.float32(1.0), rng.uniform(0, 1.3, n).astype(np.float32))  # tangent-ish
    tgt = (s[:, 0:3] + off * (s[:, 3:4] * scale[:, None])).astype(np.float32)
    d = (tgt - r[:, 0:3]).astype(np.float32)
    d = (d / np.maximum(np.sqrt((d * d).sum(1, keepdims=True)), np.float32(1e-20))).astype(np.float32)
    r[aim, 3:6] = d[aim]
    inside = rng.random(n) < 0.1
    r[inside, 0:3] = s[inside, 0:3]
    return s, r.astype(np.float32)


def plane_cases(rng, n):
    nn = rng.normal(size=(n, 3)).astype(np.float32)
    nn = (nn / np.sqrt((nn * nn).sum(1, keepdims=True))).astype(np.float32)
    p = np.concatenate([rng.uniform(-2, 2, (n, 1)).astype(np.float32), nn], 1).astype(np.float32)
    return p, rays(rng, n)


def pit_cases(rng, n):
    v = triangles(rng, n)
    v0, v1, v2 = v[:, 0:3], v[:, 3:6], v[:, 6:9]
    nn = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    nn = (nn / np.maximum(np.sqrt((nn * nn).sum(1, keepdims=True)), np.float32(1e-20))).astype(np.float32)
    w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
    mode = rng.integers(0, 5, n)
    w[mode == 1, 2] = 0
    w[mode == 2, 1] = 0
    w[mode == 3] = (w[mode == 3] * np.float32(2.5) - np.float32(0.6)).astype(np.float32)  # outside
    w = (w / np.where(np.abs(w.sum(1, keepdims=True)) < 1e-6, np.float32(1.0), w.sum(1, keepdims=True))).astype(np.float32)
    p = (w[:, 0:1] * v0 + w[:, 1:2] * v1 + w[:, 2:3] * v2).astype(np.float32)
    return v, nn, p


def shading_cases(rng, n):
    """lightPos[3] lightColor[3] rayO[3] rayD[3] t normal[3] kd[3] ks[3] shininess"""
    a = np.zeros((n, 23), np.float32)
    a[:, 0:3] = rng.uniform(-2, 2, (n, 3))
    a[:, 3:6] = rng.uniform(0, 1, (n, 3))
    a[:, 6:9] = rng.uniform(-2, 2, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.sqrt((d * d).sum(1, keepdims=True))
    a[:, 9:12] = d * np.where(rng.random((n, 1)) < 0.5, 1.0, rng.uniform(0.2, 3.0, (n, 1)))
    a[:, 12] = rng.uniform(0.1, 4, n)
    nn = rng.normal(size=(n, 3))
    nn /= np.sqrt((nn * nn).sum(1, keepdims=True))
    a[:, 13:16] = nn * np.where(rng.random((n, 1)) < 0.5, 1.0, rng.uniform(0.2, 3.0, (n, 1)))
    a[:, 16:19] = rng.uniform(0, 1, (n, 3))
    a[:, 19:22] = np.where(rng.random((n, 1)) < 0.3, 0.0, rng.uniform(0, 1, (n, 3)))
    a[:, 22] = np.where(rng.random(n) < 0.5, np.float32(10.000002), rng.choice([1.0, 2.0, 10.0, 32.5], n))
    return a.astype(np.float32)


def make_cases(n: int, seed: int) -> dict:
    rng = np.random.default_rng(seed)
    tv, tr = triangle_cases(rng, n)
    bb, br = aabb_cases(rng, n)
    sp, sr = sphere_cases(rng, n)
    pl, pr = plane_cases(rng, n)
    pv, pn, pp = pit_cases(rng, n)
    sh = shading_cases(rng, n)
    return dict(tri_v=tv, tri_ray=tr, box_b=bb, box_ray=br, sph_s=sp, sph_ray=sr, pl_p=pl, pl_ray=pr,
                pit_v=pv, pit_n=pn, pit_p=pp, shade_in=sh)


def answers(impl, c: dict) -> dict:
    """Evaluate all KATs with an implementation exposing the refharness.kat_* API."""
    out = {}
    out["tri_hit"], out["tri_t"] = impl.kat_triangle(c["tri_v"], c["tri_ray"])
    out["box_hit"], out["box_t"] = impl.kat_aabb(c["box_b"], c["box_ray"])
    out["sph_hit"], out["sph_t"], out["sph_n"] = impl.kat_sphere(c["sph_s"], c["sph_ray"])
    out["pl_hit"], out["pl_t"] = impl.kat_plane(c["pl_p"], c["pl_ray"])
    out["tp"] = impl.kat_triangle_plane(c["tri_v"])
    out["pit"] = impl.kat_point_in_triangle(c["pit_v"], c["pit_n"], c["pit_p"])
    return out


def make_golden(refharness, n: int, seed: int) -> dict:
    c = make_cases(n, seed)
    a = answers(refharness, c)
    a["bary"] = refharness.kat_barycentric(c["pit_v"], c["pit_p"])
    a["shade"] = refharness.kat_shading(c["shade_in"])
    refl_in = np.concatenate([c["shade_in"][:, 6:16], c["shade_in"][:, 19:22]], 1).astype(np.float32)
    a["refl"] = refharness.kat_reflection(refl_in)
    c["refl_in"] = refl_in
    return {**c, **{"ans_" + k: v for k, v in a.items()}}


def bits_equal(a, b) -> np.ndarray:
    """Bitwise float equality, with all NaNs considered equal (payload/sign of NaN is not specified)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
