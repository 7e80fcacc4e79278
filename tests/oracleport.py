"""ctypes binding of oracle/liboracle.so — the CPU restatement ("port") of the hot path.  TEST INFRASTRUCTURE ONLY.
Same call shapes as tests/refharness.py so that the two checkers are interchangeable in the tests."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import refharness
from refharness import RefRenderParams, RefRenderStats, CgeCamera, _p, _f32

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "liboracle.so"
_lib = None


def available() -> bool:
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(str(LIB))
        l.oracle_scene_load_flat.restype = C.c_void_p
        l.oracle_scene_load_flat.argtypes = [C.c_char_p]
        l.oracle_scene_free.argtypes = [C.c_void_p]
        l.oracle_bvh_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        l.oracle_bvh_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        l.oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RefRenderParams), C.c_void_p, C.c_void_p,
                                    C.POINTER(RefRenderStats)]
        l.oracle_camera.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(CgeCamera)]
        for name in ("oracle_kat_triangle", "oracle_kat_aabb", "oracle_kat_plane"):
            getattr(l, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.oracle_kat_sphere.argtypes = [C.c_void_p] * 4 + [C.c_uint32]
        l.oracle_kat_triangle_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.oracle_kat_point_in_triangle.argtypes = [C.c_void_p] * 4 + [C.c_uint32]
        l.oracle_kat_barycentric.argtypes = [C.c_void_p] * 3 + [C.c_uint32]
        l.oracle_kat_shading.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.oracle_kat_reflection.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.oracle_bloom.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32]
        _lib = l
    return _lib


def bloom(rgb, scalar: float, threshold: float, debug_option: int = 0) -> np.ndarray:
    """The restated renderBloomFilter (reference src/render.cpp:158-196) applied to a frame in Screen::pixels() order."""
    out = np.ascontiguousarray(rgb, dtype=np.float32).copy()
    h, w = out.shape[:2]
    lib().oracle_bloom(_p(out), w, h, scalar, threshold, debug_option)
    return out


class OracleScene:
    def __init__(self, path, features: int = 0):
        self.l = lib()
        self.scene = self.l.oracle_scene_load_flat(str(path).encode())
        if not self.scene:
            raise RuntimeError(f"oracle could not load {path}")
        self.features = features

    def close(self):
        if self.scene:
            self.l.oracle_scene_free(self.scene)
            self.scene = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def bvh_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.l.oracle_bvh_info(self.scene, C.byref(a), C.byref(b), C.byref(c))
        return {"nodes": a.value, "levels": b.value, "leaves": c.value}

    def bvh_export(self, n_prims: int):
        import importlib
        sf = importlib.import_module("computer-graphics-engine_b200.scenefile")
        nodes = np.zeros(self.bvh_info()["nodes"], sf.BVH_NODE_DT)
        order = np.zeros(n_prims, "<u4")
        root = C.c_uint32()
        self.l.oracle_bvh_export(self.scene, _p(nodes), _p(order), C.byref(root))
        return nodes, order, root.value

    def render(self, cfg: dict, threads: int = 0, want_ids: bool = True, window=None, y_stride: int = 1):
        p = refharness.RefScene.make_params(self, cfg, threads, want_ids, window, False, y_stride)
        W, H = p.width, p.height
        rgb = np.zeros((H, W, 3), np.float32)
        ids = np.full((H, W), -1, np.int32)
        st = RefRenderStats()
        rc = self.l.oracle_render(self.scene, None, C.byref(p), _p(rgb), _p(ids) if want_ids else None, C.byref(st))
        if rc:
            raise RuntimeError(f"oracle_render rc={rc}")
        return rgb, (ids if want_ids else None), {"rays": st.rays, "box_tests": st.box_tests, "tri_tests": st.tri_tests,
                                                  "sphere_tests": st.sphere_tests, "ms": st.ms}


def camera(cfg: dict) -> CgeCamera:
    cam = cfg["camera"]
    out = CgeCamera()
    look = _f32(cam["look_at"])
    rot = _f32([np.radians(np.float32(r)) for r in cam["rotation_deg"]])
    lib().oracle_camera(np.float32(np.radians(np.float32(cam["fov_deg"]))), cfg["width"], cfg["height"], _p(look),
                        cam["dist"], _p(rot), C.byref(out))
    return out


def kat_triangle(v9, ray7):
    v9, ray7 = _f32(v9), _f32(ray7).copy()
    hit = np.zeros(len(v9), np.int32)
    lib().oracle_kat_triangle(_p(v9), _p(ray7), _p(hit), len(v9))
    return hit, ray7[:, 6].copy()


def kat_aabb(b6, ray7):
    b6, ray7 = _f32(b6), _f32(ray7).copy()
    hit = np.zeros(len(b6), np.int32)
    lib().oracle_kat_aabb(_p(b6), _p(ray7), _p(hit), len(b6))
    return hit, ray7[:, 6].copy()


def kat_plane(p4, ray7):
    p4, ray7 = _f32(p4), _f32(ray7).copy()
    hit = np.zeros(len(p4), np.int32)
    lib().oracle_kat_plane(_p(p4), _p(ray7), _p(hit), len(p4))
    return hit, ray7[:, 6].copy()


def kat_sphere(s4, ray7):
    s4, ray7 = _f32(s4), _f32(ray7).copy()
    hit = np.zeros(len(s4), np.int32)
    nrm = np.zeros((len(s4), 3), np.float32)
    lib().oracle_kat_sphere(_p(s4), _p(ray7), _p(nrm), _p(hit), len(s4))
    return hit, ray7[:, 6].copy(), nrm


def kat_triangle_plane(v9):
    v9 = _f32(v9)
    out = np.zeros((len(v9), 4), np.float32)
    lib().oracle_kat_triangle_plane(_p(v9), _p(out), len(v9))
    return out


def kat_point_in_triangle(v9, n3, p3):
    v9, n3, p3 = _f32(v9), _f32(n3), _f32(p3)
    out = np.zeros(len(v9), np.int32)
    lib().oracle_kat_point_in_triangle(_p(v9), _p(n3), _p(p3), _p(out), len(v9))
    return out


def kat_barycentric(v9, p3):
    v9, p3 = _f32(v9), _f32(p3)
    out = np.zeros((len(v9), 3), np.float32)
    lib().oracle_kat_barycentric(_p(v9), _p(p3), _p(out), len(v9))
    return out


def kat_shading(in23):
    in23 = _f32(in23)
    out = np.zeros((len(in23), 3), np.float32)
    lib().oracle_kat_shading(_p(in23), _p(out), len(in23))
    return out


def kat_reflection(in13):
    in13 = _f32(in13)
    out = np.zeros((len(in13), 7), np.float32)
    lib().oracle_kat_reflection(_p(in13), _p(out), len(in13))
    return out
