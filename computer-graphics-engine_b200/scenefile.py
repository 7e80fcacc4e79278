"""Flat scene files (include/cge_scene_file.h) <-> numpy.

A flat scene is the serialised form of ``cge_scene_desc`` (include/cge.h): the reference engine's ``Scene``
(reference src/scene.h:28-33) after its own loaders ran, flattened into plain arrays.  This module only moves
bytes; it contains no rendering logic.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

MAGIC = b"CGESCN01"

HEADER_DT = np.dtype([
    ("magic", "S8"),
    ("n_meshes", "<u4"), ("n_vertices", "<u4"), ("n_triangles", "<u4"),
    ("n_spheres", "<u4"), ("n_lights", "<u4"), ("n_textures", "<u4"),
    ("n_texels", "<u8"),
    ("n_bvh_nodes", "<u4"), ("bvh_root", "<u4"),
    ("reserved", "<u4", (2,)),
])
MESH_DT = np.dtype([
    ("vertex_offset", "<u4"), ("vertex_count", "<u4"), ("triangle_offset", "<u4"), ("triangle_count", "<u4"),
    ("kd", "<f4", (3,)), ("ks", "<f4", (3,)), ("shininess", "<f4"), ("transparency", "<f4"),
    ("texture_id", "<i4"), ("reserved", "<u4", (3,)),
])
VERTEX_DT = np.dtype([("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("texcoord", "<f4", (2,))])
SPHERE_DT = np.dtype([
    ("center", "<f4", (3,)), ("radius", "<f4"), ("kd", "<f4", (3,)), ("ks", "<f4", (3,)),
    ("shininess", "<f4"), ("transparency", "<f4"), ("texture_id", "<i4"), ("reserved", "<u4"),
])
LIGHT_DT = np.dtype([("type", "<u4"), ("v", "<f4", (21,))])
TEXTURE_DT = np.dtype([("width", "<i4"), ("height", "<i4"), ("texel_offset", "<u8")])
BVH_NODE_DT = np.dtype([
    ("lower", "<f4", (3,)), ("upper", "<f4", (3,)), ("is_leaf", "<u4"), ("depth", "<u4"),
    ("beg", "<u4"), ("end", "<u4"), ("left", "<u4"), ("right", "<u4"),
])

assert HEADER_DT.itemsize == 56 and MESH_DT.itemsize == 64 and VERTEX_DT.itemsize == 32
assert SPHERE_DT.itemsize == 56 and LIGHT_DT.itemsize == 88 and TEXTURE_DT.itemsize == 16
assert BVH_NODE_DT.itemsize == 48

LIGHT_POINT, LIGHT_SEGMENT, LIGHT_PARALLELOGRAM = 0, 1, 2


@dataclass
class FlatScene:
    meshes: np.ndarray = field(default_factory=lambda: np.zeros(0, MESH_DT))
    vertices: np.ndarray = field(default_factory=lambda: np.zeros(0, VERTEX_DT))
    triangles: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), "<u4"))
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, SPHERE_DT))
    lights: np.ndarray = field(default_factory=lambda: np.zeros(0, LIGHT_DT))
    textures: np.ndarray = field(default_factory=lambda: np.zeros(0, TEXTURE_DT))
    texels: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), "<f4"))
    bvh_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, BVH_NODE_DT))
    bvh_prim_order: np.ndarray = field(default_factory=lambda: np.zeros(0, "<u4"))
    bvh_root: int = 0

    @property
    def n_triangles(self) -> int:
        return int(self.triangles.shape[0])

    @property
    def n_primitives(self) -> int:
        return self.n_triangles + int(self.spheres.shape[0])

    def copy(self) -> "FlatScene":
        return copy.deepcopy(self)

    def without_bvh(self) -> "FlatScene":
        s = self.copy()
        s.bvh_nodes = np.zeros(0, BVH_NODE_DT)
        s.bvh_prim_order = np.zeros(0, "<u4")
        s.bvh_root = 0
        return s

    # ---- composition helpers (host-side scene assembly for the non-reference configs C3-C5) ----
    def append_meshes(self, other: "FlatScene", scale: float = 1.0, translate=(0.0, 0.0, 0.0)) -> None:
        """Scene.meshes append (what loadMesh + std::move does in reference src/scene.cpp:12-14), with an
        optional uniform scale + translation applied to the appended vertex positions in fp32."""
        m = other.meshes.copy()
        m["vertex_offset"] += np.uint32(len(self.vertices))
        m["triangle_offset"] += np.uint32(len(self.triangles))
        tex_shift = len(self.textures)
        m["texture_id"] = np.where(m["texture_id"] >= 0, m["texture_id"] + tex_shift, -1)
        v = other.vertices.copy()
        if scale != 1.0 or any(t != 0.0 for t in translate):
            v["position"] = (v["position"] * np.float32(scale) + np.asarray(translate, np.float32)).astype(np.float32)
        t = other.textures.copy()
        t["texel_offset"] += np.uint64(len(self.texels))
        self.meshes = np.concatenate([self.meshes, m])
        self.vertices = np.concatenate([self.vertices, v])
        self.triangles = np.concatenate([self.triangles, other.triangles])
        self.textures = np.concatenate([self.textures, t])
        self.texels = np.concatenate([self.texels, other.texels])
        self.bvh_nodes = np.zeros(0, BVH_NODE_DT)
        self.bvh_prim_order = np.zeros(0, "<u4")

    def set_lights(self, lights) -> None:
        arr = np.zeros(len(lights), LIGHT_DT)
        for i, (kind, values) in enumerate(lights):
            vals = np.asarray(values, np.float32).ravel()
            arr[i]["type"] = kind
            arr[i]["v"][: len(vals)] = vals
        self.lights = arr


def point_light(position, color):
    return (LIGHT_POINT, list(position) + list(color))


def segment_light(e0, e1, c0, c1):
    return (LIGHT_SEGMENT, list(e0) + list(e1) + list(c0) + list(c1))


def parallelogram_light(v0, edge01, edge02, c0, c1, c2, c3):
    return (LIGHT_PARALLELOGRAM, list(v0) + list(edge01) + list(edge02) + list(c0) + list(c1) + list(c2) + list(c3))


def load(path) -> FlatScene:
    buf = np.fromfile(str(path), dtype=np.uint8)
    if len(buf) < HEADER_DT.itemsize:
        raise ValueError(f"{path}: not a flat scene file")
    h = buf[: HEADER_DT.itemsize].view(HEADER_DT)[0]
    if bytes(h["magic"]) != MAGIC:
        raise ValueError(f"{path}: not a flat scene file")
    off = HEADER_DT.itemsize

    def take(dt, count, shape=None):
        nonlocal off
        dt = np.dtype(dt)
        nbytes = dt.itemsize * int(count)
        if off + nbytes > len(buf):
            raise ValueError(f"{path}: truncated scene file")
        a = buf[off: off + nbytes].view(dt).copy()
        off += nbytes
        return a.reshape(shape) if shape is not None else a

    s = FlatScene()
    s.meshes = take(MESH_DT, h["n_meshes"])
    s.vertices = take(VERTEX_DT, h["n_vertices"])
    s.triangles = take("<u4", int(h["n_triangles"]) * 3, (-1, 3))
    s.spheres = take(SPHERE_DT, h["n_spheres"])
    s.lights = take(LIGHT_DT, h["n_lights"])
    s.textures = take(TEXTURE_DT, h["n_textures"])
    s.texels = take("<f4", int(h["n_texels"]) * 3, (-1, 3))
    s.bvh_nodes = take(BVH_NODE_DT, h["n_bvh_nodes"])
    if int(h["n_bvh_nodes"]):
        s.bvh_prim_order = take("<u4", int(h["n_triangles"]) + int(h["n_spheres"]))
    s.bvh_root = int(h["bvh_root"])
    return s


def save(scene: FlatScene, path) -> None:
    h = np.zeros(1, HEADER_DT)
    h["magic"] = MAGIC
    h["n_meshes"] = len(scene.meshes)
    h["n_vertices"] = len(scene.vertices)
    h["n_triangles"] = len(scene.triangles)
    h["n_spheres"] = len(scene.spheres)
    h["n_lights"] = len(scene.lights)
    h["n_textures"] = len(scene.textures)
    h["n_texels"] = len(scene.texels)
    h["n_bvh_nodes"] = len(scene.bvh_nodes)
    h["bvh_root"] = scene.bvh_root
    with open(str(path), "wb") as f:
        f.write(h.tobytes())
        for a, dt in ((scene.meshes, MESH_DT), (scene.vertices, VERTEX_DT), (scene.triangles, "<u4"),
                      (scene.spheres, SPHERE_DT), (scene.lights, LIGHT_DT), (scene.textures, TEXTURE_DT),
                      (scene.texels, "<f4"), (scene.bvh_nodes, BVH_NODE_DT)):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
        if len(scene.bvh_nodes):
            f.write(np.ascontiguousarray(scene.bvh_prim_order, dtype="<u4").tobytes())
