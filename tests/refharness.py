"""ctypes binding of oracle/_ref/libcge_ref*.so — the UNMODIFIED reference engine built headless.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / "oracle" / "_ref"


class RefRenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("features", C.c_uint32), ("ray_depth", C.c_int32),
        ("segment_samples", C.c_int32), ("parallelogram_samples", C.c_int32), ("sampler", C.c_uint32),
        ("seed", C.c_uint32), ("threads", C.c_int32), ("use_render_ray_tracing", C.c_int32),
        ("want_ids", C.c_int32), ("fovy", C.c_float), ("look_at", C.c_float * 3), ("dist", C.c_float),
        ("rotation", C.c_float * 3), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("y_stride", C.c_int32),
        ("rays_per_pixel_side", C.c_int32), ("bloom_scalar", C.c_float), ("bloom_threshold", C.c_float),
        ("bloom_debug_option", C.c_int32),
    ]


class RefRenderStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("ms", C.c_double)]


class CgeCamera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("quat", C.c_float * 4), ("half_width", C.c_float),
                ("half_height", C.c_float)]


def _lib_name(plain) -> str:
    # plain = "gpu": the reference engine linked with computer-graphics-engine_b200/host/render_gpu.cpp, i.e. its own
    # renderRayTracing(scene, camera, bvh, screen, features) answered by libcge.so (oracle/ref/build_ref.sh)
    return "libcge_ref_gpu.so" if plain == "gpu" else "libcge_ref_plain.so" if plain else "libcge_ref.so"


def available(plain=False) -> bool:
    return (REF_DIR / _lib_name(plain)).exists()


_libs = {}


def lib(plain=False):
    key = plain if plain == "gpu" else bool(plain)
    if key not in _libs:
        path = REF_DIR / _lib_name(plain)
        l = C.CDLL(str(path))
        l.ref_scene_load_flat.restype = C.c_void_p
        l.ref_scene_load_flat.argtypes = [C.c_char_p]
        l.ref_scene_free.argtypes = [C.c_void_p]
        l.ref_bvh_build.restype = C.c_void_p
        l.ref_bvh_build.argtypes = [C.c_void_p, C.c_uint32]
        l.ref_bvh_free.argtypes = [C.c_void_p]
        l.ref_bvh_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        l.ref_scene_export.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_int, C.c_char_p]
        l.ref_scene_write_with_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
        l.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RefRenderParams), C.c_void_p, C.c_void_p,
                                 C.POINTER(RefRenderStats)]
        l.ref_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(RefRenderParams),
                                     C.c_void_p, C.c_void_p]
        l.ref_intersect_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
        l.ref_camera.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(CgeCamera)]
        l.ref_generate_rays.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_uint32]
        for name in ("ref_kat_triangle", "ref_kat_aabb", "ref_kat_plane"):
            getattr(l, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_kat_sphere.argtypes = [C.c_void_p] * 4 + [C.c_uint32]
        l.ref_kat_triangle_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_kat_point_in_triangle.argtypes = [C.c_void_p] * 4 + [C.c_uint32]
        l.ref_kat_barycentric.argtypes = [C.c_void_p] * 3 + [C.c_uint32]
        l.ref_kat_shading.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_kat_reflection.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_weights_gaussian.argtypes = [C.c_float, C.c_void_p]
        l.ref_ray_samples.argtypes = [C.POINTER(RefRenderParams), C.c_int, C.c_int, C.c_void_p]
        _libs[key] = l
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefScene:
    """Reference ``Scene`` + ``BvhInterface`` built from a flat scene file."""

    def __init__(self, path, features: int, plain=False):
        self.l = lib(plain)
        self.scene = self.l.ref_scene_load_flat(str(path).encode())
        if not self.scene:
            raise RuntimeError(f"reference harness could not load {path}")
        self.bvh = self.l.ref_bvh_build(self.scene, features)
        self.features = features

    def close(self):
        if self.bvh:
            self.l.ref_bvh_free(self.bvh)
            self.bvh = None
        if self.scene:
            self.l.ref_scene_free(self.scene)
            self.scene = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def bvh_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.l.ref_bvh_info(self.bvh, C.byref(a), C.byref(b), C.byref(c))
        return {"nodes": a.value, "levels": b.value, "leaves": c.value}

    def write_with_bvh(self, out_path):
        rc = self.l.ref_scene_write_with_bvh(self.scene, self.features, str(out_path).encode())
        if rc:
            raise RuntimeError("ref_scene_write_with_bvh failed")

    def make_params(self, cfg: dict, threads: int = 0, want_ids: bool = True, window=None,
                    use_render_ray_tracing: bool = False, y_stride: int = 1) -> RefRenderParams:
        p = RefRenderParams()
        p.width, p.height = cfg["width"], cfg["height"]
        p.features = cfg["features"]
        p.ray_depth = cfg.get("ray_depth", 5)
        p.segment_samples = cfg.get("segment_samples", 25)
        p.parallelogram_samples = cfg.get("parallelogram_samples", 5)
        p.sampler = 1
        p.seed = cfg.get("seed", 0)
        p.threads = threads
        p.use_render_ray_tracing = int(use_render_ray_tracing)
        p.want_ids = int(want_ids)
        cam = cfg["camera"]
        p.fovy = np.float32(np.radians(np.float32(cam["fov_deg"])))
        p.look_at = (C.c_float * 3)(*cam["look_at"])
        p.dist = cam["dist"]
        p.rotation = (C.c_float * 3)(*[np.float32(np.radians(np.float32(r))) for r in cam["rotation_deg"]])
        if window:
            p.x0, p.y0, p.x1, p.y1 = window
        p.y_stride = y_stride
        # ExtraFeatures globals (reference src/render.cpp:14,19-21 defaults)
        p.rays_per_pixel_side = cfg.get("rays_per_pixel_side", 3)
        p.bloom_scalar = cfg.get("bloom_scalar", 0.3)
        p.bloom_threshold = cfg.get("bloom_threshold", 0.4)
        p.bloom_debug_option = cfg.get("bloom_debug_option", 0)
        return p

    def render(self, cfg: dict, threads: int = 0, want_ids: bool = True, window=None,
               use_render_ray_tracing: bool = False, y_stride: int = 1):
        p = self.make_params(cfg, threads, want_ids, window, use_render_ray_tracing, y_stride)
        W, H = p.width, p.height
        rgb = np.zeros((H, W, 3), np.float32)
        ids = np.full((H, W), -1, np.int32)
        st = RefRenderStats()
        rc = self.l.ref_render(self.scene, self.bvh, C.byref(p), _p(rgb), _p(ids) if want_ids else None, C.byref(st))
        if rc:
            raise RuntimeError(f"ref_render rc={rc}")
        stats = {"rays": st.rays, "box_tests": st.box_tests, "tri_tests": st.tri_tests,
                 "sphere_tests": st.sphere_tests, "ms": st.ms}
        return rgb, (ids if want_ids else None), stats

    def trace_rays(self, rays7, cfg: dict, want_ids: bool = True):
        rays7 = _f32(rays7)
        n = rays7.shape[0]
        p = self.make_params(cfg)
        rgb = np.zeros((n, 3), np.float32)
        ids = np.full(n, -1, np.int32)
        self.l.ref_trace_rays(self.scene, self.bvh, _p(rays7), n, C.byref(p), _p(rgb), _p(ids) if want_ids else None)
        return rgb, ids

    def intersect_rays(self, rays7, features=None):
        rays7 = _f32(rays7)
        n = rays7.shape[0]
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        mat = np.zeros((n, 8), np.float32)
        ids = np.full(n, -1, np.int32)
        self.l.ref_intersect_rays(self.scene, self.bvh, _p(rays7), n, self.features if features is None else features,
                                  _p(hit), _p(t), _p(nrm), _p(mat), _p(ids))
        return hit, t, nrm, mat, ids


def weights_gaussian(sigma: float = 1.0) -> np.ndarray:
    """The reference's weightsGaussian(sigma) (src/render.cpp:198-210) as a 3x3 array indexed [k + 1][j + 1]."""
    out = np.zeros(9, np.float32)
    lib().ref_weights_gaussian(sigma, _p(out))
    return out.reshape(3, 3)


def ray_samples(cfg: dict, x: int, y: int) -> np.ndarray:
    """The reference's getRaySamples (src/render.cpp:211-227) for pixel (x, y) with the hash-seeded std::mt19937:
    (n*n, 6) origin + direction."""
    p = RefScene.make_params(None, cfg)
    n = p.rays_per_pixel_side
    out = np.zeros((n * n, 6), np.float32)
    got = lib().ref_ray_samples(C.byref(p), x, y, _p(out))
    assert got == n * n
    return out


def camera(cfg: dict) -> CgeCamera:
    cam = cfg["camera"]
    out = CgeCamera()
    look = _f32(cam["look_at"])
    rot = _f32([np.radians(np.float32(r)) for r in cam["rotation_deg"]])
    lib().ref_camera(np.float32(np.radians(np.float32(cam["fov_deg"]))), cfg["width"], cfg["height"], _p(look),
                     cam["dist"], _p(rot), C.byref(out))
    return out


def generate_rays(cfg: dict, ndc):
    cam = cfg["camera"]
    ndc = _f32(ndc)
    n = ndc.shape[0]
    look = _f32(cam["look_at"])
    rot = _f32([np.radians(np.float32(r)) for r in cam["rotation_deg"]])
    out = np.zeros((n, 7), np.float32)
    lib().ref_generate_rays(np.float32(np.radians(np.float32(cam["fov_deg"]))), cfg["width"], cfg["height"], _p(look),
                            cam["dist"], _p(rot), _p(ndc), _p(out), n)
    return out


# ---- I1-I6 / S KATs ------------------------------------------------------------------------------------
def kat_triangle(v9, ray7):
    v9, ray7 = _f32(v9), _f32(ray7).copy()
    hit = np.zeros(len(v9), np.int32)
    lib().ref_kat_triangle(_p(v9), _p(ray7), _p(hit), len(v9))
    return hit, ray7[:, 6].copy()


def kat_aabb(b6, ray7):
    b6, ray7 = _f32(b6), _f32(ray7).copy()
    hit = np.zeros(len(b6), np.int32)
    lib().ref_kat_aabb(_p(b6), _p(ray7), _p(hit), len(b6))
    return hit, ray7[:, 6].copy()


def kat_plane(p4, ray7):
    p4, ray7 = _f32(p4), _f32(ray7).copy()
    hit = np.zeros(len(p4), np.int32)
    lib().ref_kat_plane(_p(p4), _p(ray7), _p(hit), len(p4))
    return hit, ray7[:, 6].copy()


def kat_sphere(s4, ray7):
    s4, ray7 = _f32(s4), _f32(ray7).copy()
    hit = np.zeros(len(s4), np.int32)
    nrm = np.zeros((len(s4), 3), np.float32)
    lib().ref_kat_sphere(_p(s4), _p(ray7), _p(nrm), _p(hit), len(s4))
    return hit, ray7[:, 6].copy(), nrm


def kat_triangle_plane(v9):
    v9 = _f32(v9)
    out = np.zeros((len(v9), 4), np.float32)
    lib().ref_kat_triangle_plane(_p(v9), _p(out), len(v9))
    return out


def kat_point_in_triangle(v9, n3, p3):
    v9, n3, p3 = _f32(v9), _f32(n3), _f32(p3)
    out = np.zeros(len(v9), np.int32)
    lib().ref_kat_point_in_triangle(_p(v9), _p(n3), _p(p3), _p(out), len(v9))
    return out


def kat_barycentric(v9, p3):
    v9, p3 = _f32(v9), _f32(p3)
    out = np.zeros((len(v9), 3), np.float32)
    lib().ref_kat_barycentric(_p(v9), _p(p3), _p(out), len(v9))
    return out


def kat_shading(in23):
    in23 = _f32(in23)
    out = np.zeros((len(in23), 3), np.float32)
    lib().ref_kat_shading(_p(in23), _p(out), len(in23))
    return out


def kat_reflection(in13):
    in13 = _f32(in13)
    out = np.zeros((len(in13), 7), np.float32)
    lib().ref_kat_reflection(_p(in13), _p(out), len(in13))
    return out


def write_bmp(rgb, path):
    """The reference's Screen::writeBitmapToFile on a float frame (H, W, 3)."""
    rgb = _f32(rgb)
    l = lib()
    l.ref_write_bmp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    l.ref_write_bmp(_p(rgb), rgb.shape[1], rgb.shape[0], str(path).encode())


def export_prebuilt(scene_type: int, data_dir, out_path, features: int = 0, with_bvh: bool = False):
    rc = lib().ref_scene_export(0, scene_type, str(data_dir).encode(), features, int(with_bvh), str(out_path).encode())
    if rc:
        raise RuntimeError(f"ref_scene_export rc={rc}")


def export_obj(obj_path, normalize: bool, out_path):
    rc = lib().ref_scene_export(1, int(normalize), str(obj_path).encode(), 0, 0, str(out_path).encode())
    if rc:
        raise RuntimeError(f"ref_scene_export rc={rc}")
