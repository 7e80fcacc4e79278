"""Declared stand-in for ``data/dragon.obj`` (config C5).

The reference checkout does not contain dragon.obj (``.MISSING_LARGE_BLOBS``; reference src/scene.cpp:68-73 expects
it), so BOTH sides — the CUDA path and the reference renderer used as test oracle — are fed this procedurally generated
mesh instead.  It has 12*269^2 = 868 332 triangles (the Stanford dragon the reference names has ~871 K) plus a
2-triangle mirror ground plane, is scaled into the unit sphere like ``loadMesh(..., normalize=true)`` output, and is
strongly displaced so that it self-shadows and inter-reflects.

Only +, -, *, / and sqrt on float32 arrays are used (all IEEE-exact in numpy), never sin/cos, so every machine
generates bit-identical vertices.
"""
from __future__ import annotations

import numpy as np

from . import scenefile
from .scenefile import FlatScene

_cache: dict = {}

F = np.float32


def _cheb(x, k):
    """Chebyshev T_k(x) by the three-term recurrence (mul/sub only)."""
    t0 = np.ones_like(x)
    t1 = x
    if k == 0:
        return t0
    for _ in range(k - 1):
        t0, t1 = t1, (F(2.0) * x * t1 - t0).astype(F)
    return t1


def _displace(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    f = (F(0.50) * _cheb(x, 3) * _cheb(y, 2)
         + F(0.35) * _cheb(z, 5) * _cheb(x, 4)
         + F(0.25) * _cheb(y, 9) * _cheb(z, 7)
         + F(0.15) * _cheb(x, 17) * _cheb(y, 13) * _cheb(z, 11)).astype(F)
    r = (F(0.62) * (F(1.0) + F(0.28) * f)).astype(F)
    return (p * r[:, None]).astype(F)


def _cube_sphere(n: int):
    g = (np.arange(n + 1, dtype=F) * F(2.0) / F(n) - F(1.0)).astype(F)
    a, b = np.meshgrid(g, g, indexing="ij")
    a, b = a.ravel(), b.ravel()
    one = np.ones_like(a)
    faces = [
        np.stack([one, a, b], 1), np.stack([-one, b, a], 1),
        np.stack([b, one, a], 1), np.stack([a, -one, b], 1),
        np.stack([a, b, one], 1), np.stack([b, a, -one], 1),
    ]
    i, j = np.meshgrid(np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32), indexing="ij")
    i, j = i.ravel(), j.ravel()
    v00 = i * np.uint32(n + 1) + j
    v10 = v00 + np.uint32(n + 1)
    v01 = v00 + np.uint32(1)
    v11 = v10 + np.uint32(1)
    quad = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 0)
    verts, tris = [], []
    for f, cube in enumerate(faces):
        cube = cube.astype(F)
        inv = (F(1.0) / np.sqrt((cube * cube).sum(1, dtype=F))).astype(F)
        verts.append((cube * inv[:, None]).astype(F))
        tris.append(quad + np.uint32(f * (n + 1) * (n + 1)))
    return np.concatenate(verts, 0), np.concatenate(tris, 0).astype(np.uint32)


def make(name: str = "dragon", n: int = 269) -> FlatScene:
    if name != "dragon":
        raise KeyError(name)
    key = (name, n)
    if key in _cache:
        return _cache[key].copy()
    sphere, tris = _cube_sphere(n)
    pos = _displace(sphere)
    s = FlatScene()
    verts = np.zeros(len(pos) + 4, scenefile.VERTEX_DT)
    verts["position"][: len(pos)] = pos
    verts["normal"][: len(pos)] = sphere
    ground_y = F(-0.82)
    e = F(2.5)
    verts["position"][len(pos):] = np.array([[-e, ground_y, -e], [e, ground_y, -e], [e, ground_y, e], [-e, ground_y, e]], F)
    verts["normal"][len(pos):] = np.array([0, 1, 0], F)
    verts["texcoord"][len(pos):] = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F)
    meshes = np.zeros(2, scenefile.MESH_DT)
    meshes[0]["vertex_offset"], meshes[0]["vertex_count"] = 0, len(pos)
    meshes[0]["triangle_offset"], meshes[0]["triangle_count"] = 0, len(tris)
    meshes[0]["kd"], meshes[0]["ks"] = (0.55, 0.70, 0.45), (0.25, 0.25, 0.25)
    meshes[0]["shininess"], meshes[0]["transparency"], meshes[0]["texture_id"] = 20.0, 1.0, -1
    meshes[1]["vertex_offset"], meshes[1]["vertex_count"] = len(pos), 4
    meshes[1]["triangle_offset"], meshes[1]["triangle_count"] = len(tris), 2
    meshes[1]["kd"], meshes[1]["ks"] = (0.60, 0.60, 0.65), (0.35, 0.35, 0.35)
    meshes[1]["shininess"], meshes[1]["transparency"], meshes[1]["texture_id"] = 50.0, 1.0, -1
    s.meshes = meshes
    s.vertices = verts
    s.triangles = np.concatenate([tris, np.array([[0, 2, 1], [0, 3, 2]], np.uint32)], 0)
    s.set_lights([scenefile.parallelogram_light(
        v0=(-0.6, 1.9, -1.6), edge01=(0.6, 0.0, 0.0), edge02=(0.0, 0.0, 0.6),
        c0=(1.0, 1.0, 1.0), c1=(1.0, 0.95, 0.85), c2=(0.85, 0.95, 1.0), c3=(1.0, 1.0, 1.0))])
    _cache[key] = s
    return s.copy()
