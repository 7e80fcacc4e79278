"""computer-graphics-engine_b200 — B200-native ray-tracing hot path of Anton-Kalpakchiev/Computer-Graphics-Engine.

This Python module is glue for tests and bench.py only: it loads the C-ABI shared library ``libcge.so``
(include/cge.h, built by csrc/build.sh for sm_100a) with ctypes and moves numpy arrays in and out of it.  All
rendering happens in the CUDA kernels behind that ABI; there is NO Python or CPU fallback — if the library is
missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import configs, scenefile, standin  # noqa: F401
from .scenefile import FlatScene

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CGE_LIB", _PKG / "libcge.so"))  # CGE_LIB: development A/B builds only

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_NOMEM = range(6)
TRAVERSAL_REFERENCE, TRAVERSAL_FAST = 0, 1
FLAG_WANT_PRIM_IDS, FLAG_RGB_DEVICE_PTR, FLAG_COUNT_TESTS, FLAG_OUTPUT_RGBA8 = 1, 2, 4, 8
FLAG_PARTITION_TILE_ROWS, FLAG_SHARED_HOST_FRAME, FLAG_PEER_FRAME, FLAG_DYNAMIC_TILES = 16, 32, 64, 128
# development switches (include/cge.h CGE_DEV_FLAG_*): force one of the two production pipelines / the per-pixel cost map
FLAG_PER_THREAD, FLAG_WAVEFRONT, FLAG_DEBUG_CYCLES = 1 << 16, 1 << 17, 1 << 18
UNIQUE_ID_BYTES = 128


class CgeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cge error {code}: {msg}")
        self.code = code


class CgeSceneDesc(C.Structure):
    _fields_ = [
        ("n_meshes", C.c_uint32), ("n_vertices", C.c_uint32), ("n_triangles", C.c_uint32),
        ("n_spheres", C.c_uint32), ("n_lights", C.c_uint32), ("n_textures", C.c_uint32),
        ("n_texels", C.c_uint64),
        ("meshes", C.c_void_p), ("vertices", C.c_void_p), ("triangles", C.c_void_p), ("spheres", C.c_void_p),
        ("lights", C.c_void_p), ("textures", C.c_void_p), ("texels", C.c_void_p),
        ("n_bvh_nodes", C.c_uint32), ("bvh_root", C.c_uint32),
        ("bvh_nodes", C.c_void_p), ("bvh_prim_order", C.c_void_p),
    ]


class CgeCamera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("quat", C.c_float * 4), ("half_width", C.c_float),
                ("half_height", C.c_float)]


class CgeParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("features", C.c_uint32), ("ray_depth", C.c_int32),
        ("segment_samples", C.c_int32), ("parallelogram_samples", C.c_int32), ("sampler", C.c_uint32),
        ("seed", C.c_uint32), ("traversal", C.c_uint32), ("flags", C.c_uint32),
        ("part_index", C.c_uint32), ("part_count", C.c_uint32),
        ("rays_per_pixel_side", C.c_int32), ("bloom_scalar", C.c_float), ("bloom_threshold", C.c_float),
        ("bloom_debug_option", C.c_int32),
    ]


class CgeStats(C.Structure):
    _fields_ = [
        ("primary_rays", C.c_uint64), ("bounce_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("reference_rays", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
        ("reference_shadow_rays", C.c_uint64),
        ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("kernel_launches", C.c_uint32),
        ("stage_ms", C.c_float * 4),
        ("shadow_samples_culled", C.c_uint64), ("vis_cull_ms", C.c_float),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "stage_ms"}
        d["stage_ms"] = [float(x) for x in self.stage_ms]
        d["gpu_rays"] = d["primary_rays"] + d["bounce_rays"] + d["shadow_rays"]
        return d


# every symbol include/cge.h declares (checked by tests/test_abi.py without a GPU)
ABI_SYMBOLS = [
    "cge_abi_version", "cge_last_error", "cge_device_count", "cge_camera_from_trackball", "cge_scene_create",
    "cge_scene_update_lights", "cge_scene_destroy", "cge_scene_bvh_info", "cge_scene_bvh_export", "cge_render",
    "cge_bvh_build_reference_order", "cge_bvh_validate", "cge_fast_bvh_build", "cge_ray_sample_positions", "cge_bloom_weights",
    "cge_trace_rays", "cge_kat_triangle", "cge_kat_triangle_precomputed", "cge_kat_aabb", "cge_kat_sphere",
    "cge_kat_plane", "cge_kat_triangle_plane", "cge_kat_point_in_triangle", "cge_comm_unique_id", "cge_comm_create",
    "cge_comm_destroy", "cge_comm_host_frame", "cge_comm_peer_frame", "cge_render_distributed", "cge_hull_clear_host", "cge_hull_box_host", "cge_host_alloc", "cge_host_free",
]

_lib = None


def lib() -> C.CDLL:
    """Load libcge.so.  Raises if it has not been built — there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the ray-tracing path exists only as CUDA code behind this library)")
        l = C.CDLL(str(LIB_PATH))
        l.cge_last_error.restype = C.c_char_p
        l.cge_scene_create.argtypes = [C.POINTER(CgeSceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        l.cge_scene_update_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.cge_scene_destroy.argtypes = [C.c_void_p]
        l.cge_scene_bvh_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4
        l.cge_scene_bvh_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.cge_bvh_build_reference_order.argtypes = [C.POINTER(CgeSceneDesc), C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p] + [C.POINTER(C.c_uint32)] * 3
        if hasattr(l, "cge_bvh_validate"):
            l.cge_bvh_validate.argtypes = [C.POINTER(CgeSceneDesc), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.cge_camera_from_trackball.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p,
                                                C.POINTER(CgeCamera)]
        l.cge_fast_bvh_build.argtypes = [C.POINTER(CgeSceneDesc), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p] \
            + [C.POINTER(C.c_uint32)] * 3 + [C.POINTER(C.c_float)]
        l.cge_ray_sample_positions.argtypes = [C.c_int32] * 5 + [C.c_uint32, C.c_void_p]
        l.cge_bloom_weights.argtypes = [C.c_float, C.c_void_p]
        l.cge_render.argtypes = [C.c_void_p, C.POINTER(CgeCamera), C.POINTER(CgeParams), C.c_void_p, C.c_void_p,
                                 C.POINTER(CgeStats)]
        l.cge_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(CgeParams), C.c_void_p, C.c_void_p]
        for name in ("cge_kat_triangle", "cge_kat_triangle_precomputed", "cge_kat_aabb", "cge_kat_plane"):
            getattr(l, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        l.cge_kat_sphere.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_int]
        l.cge_kat_triangle_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        l.cge_kat_point_in_triangle.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_int]
        l.cge_comm_unique_id.argtypes = [C.c_void_p]
        l.cge_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.cge_comm_destroy.argtypes = [C.c_void_p]
        if hasattr(l, "cge_comm_host_frame"):  # (absent from older development builds loaded through CGE_LIB for A/B runs)
            l.cge_comm_host_frame.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        if hasattr(l, "cge_comm_peer_frame"):
            l.cge_comm_peer_frame.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        l.cge_render_distributed.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CgeCamera), C.POINTER(CgeParams),
                                             C.c_void_p, C.c_void_p, C.POINTER(CgeStats)]
        l.cge_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
        l.cge_host_free.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def _check(rc: int):
    if rc != OK:
        raise CgeError(rc, lib().cge_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return int(lib().cge_device_count())


def camera_from_cfg(cfg: dict) -> CgeCamera:
    """cge_camera for a config dict, computed the way the reference Trackball does (fp32 throughout)."""
    cam = cfg["camera"]
    out = CgeCamera()
    fovy = np.float32(np.radians(np.float32(cam["fov_deg"])))
    aspect = np.float32(cfg["width"]) / np.float32(cfg["height"])
    look = np.asarray(cam["look_at"], np.float32)
    rot = np.asarray([np.float32(np.radians(np.float32(r))) for r in cam["rotation_deg"]], np.float32)
    _check(lib().cge_camera_from_trackball(fovy, aspect, _p(look), np.float32(cam["dist"]), _p(rot), C.byref(out)))
    return out


def ray_sample_positions(width: int, height: int, x: int, y: int, n: int, seed: int) -> np.ndarray:
    """Host-only: the n*n NDC sample positions of the anti-aliasing feature for one pixel (include/cge.h)."""
    out = np.zeros((n * n, 2), np.float32)
    _check(lib().cge_ray_sample_positions(width, height, x, y, n, seed, _p(out)))
    return out


def bloom_weights(sigma: float = 1.0) -> np.ndarray:
    """Host-only: the library's weightsGaussian(sigma) as a 3x3 array."""
    out = np.zeros(9, np.float32)
    _check(lib().cge_bloom_weights(sigma, _p(out)))
    return out.reshape(3, 3)


def params_from_cfg(cfg: dict, traversal: int = TRAVERSAL_FAST, want_ids: bool = True, part=(0, 1),
                    flags: int = 0) -> CgeParams:
    p = CgeParams()
    p.width, p.height = cfg["width"], cfg["height"]
    p.features = cfg["features"]
    p.ray_depth = cfg.get("ray_depth", 5)
    p.segment_samples = cfg.get("segment_samples", 25)
    p.parallelogram_samples = cfg.get("parallelogram_samples", 5)
    p.sampler = 0
    p.seed = cfg.get("seed", 0)
    p.traversal = traversal
    p.flags = (FLAG_WANT_PRIM_IDS if want_ids else 0) | flags
    p.part_index, p.part_count = part
    # globals of the two implemented ExtraFeatures, reference defaults (src/render.cpp:14,19-21)
    p.rays_per_pixel_side = cfg.get("rays_per_pixel_side", 3)
    p.bloom_scalar = cfg.get("bloom_scalar", 0.3)
    p.bloom_threshold = cfg.get("bloom_threshold", 0.4)
    p.bloom_debug_option = cfg.get("bloom_debug_option", 0)
    return p


class PinnedBuffer:
    """Page-locked host array (cudaHostAlloc through the C ABI) for full-speed D2H of the framebuffer."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = C.c_void_p()
        _check(lib().cge_host_alloc(C.byref(self._ptr), max(nbytes, 1)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if self._ptr:
            self.array = None
            lib().cge_host_free(self._ptr)
            self._ptr = None


def scene_desc(flat: FlatScene):
    """(cge_scene_desc, arrays to keep alive) for a flat scene."""
    d = CgeSceneDesc()
    keep = [np.ascontiguousarray(a) for a in (flat.meshes, flat.vertices, flat.triangles, flat.spheres,
                                              flat.lights, flat.textures, flat.texels)]
    d.n_meshes, d.n_vertices, d.n_triangles = len(flat.meshes), len(flat.vertices), len(flat.triangles)
    d.n_spheres, d.n_lights, d.n_textures, d.n_texels = len(flat.spheres), len(flat.lights), len(flat.textures), len(flat.texels)
    (d.meshes, d.vertices, d.triangles, d.spheres, d.lights, d.textures, d.texels) = [
        a.ctypes.data if a.size else None for a in keep]
    return d, keep


def build_reference_bvh_host(flat: FlatScene):
    """Host-only rebuild of the reference's tree (no GPU): (nodes, prim_order, root, levels, leaves)."""
    d, keep = scene_desc(flat)
    n_prims = flat.n_primitives
    nodes = np.zeros(max(2 * n_prims, 1), scenefile.BVH_NODE_DT)
    order = np.zeros(n_prims, "<u4")
    n = C.c_uint32(len(nodes))
    root, levels, leaves = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().cge_bvh_build_reference_order(C.byref(d), _p(nodes), C.byref(n), _p(order) if n_prims else None, C.byref(root),
                                               C.byref(levels), C.byref(leaves)))
    return nodes[: n.value], order, root.value, levels.value, leaves.value


def validate_bvh(flat: FlatScene, nodes: np.ndarray, order: np.ndarray, root: int):
    """Host-only: the check cge_scene_create applies to a caller-supplied tree.  Returns (levels, leaves); raises CgeError."""
    d, keep = scene_desc(flat)
    nodes = np.ascontiguousarray(nodes, scenefile.BVH_NODE_DT)
    order = np.ascontiguousarray(order, "<u4")
    d.n_bvh_nodes, d.bvh_root = len(nodes), root
    d.bvh_nodes, d.bvh_prim_order = nodes.ctypes.data if len(nodes) else None, order.ctypes.data if len(order) else None
    levels, leaves = C.c_uint32(), C.c_uint32()
    _check(lib().cge_bvh_validate(C.byref(d), C.byref(levels), C.byref(leaves)))
    return levels.value, leaves.value


FAST_NODE_DT = np.dtype([("left_lower", "<f4", (3,)), ("left_upper", "<f4", (3,)), ("right_lower", "<f4", (3,)),
                         ("right_upper", "<f4", (3,)), ("left", "<u4"), ("right", "<u4")])


def build_fast_bvh(flat: FlatScene, on_gpu: bool, device: int = 0) -> dict:
    """The FAST traversal tree by the GPU builder (the one scenes use) or the host builder (its checker; no GPU needed)."""
    d, keep = scene_desc(flat)
    n_prims = flat.n_primitives
    nodes = np.zeros(max(n_prims, 1), FAST_NODE_DT)
    order = np.zeros(n_prims, "<u4")
    n = C.c_uint32(len(nodes))
    root, depth, leaves, ms = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_float()
    _check(lib().cge_fast_bvh_build(C.byref(d), int(on_gpu), device, _p(nodes), C.byref(n), _p(order) if n_prims else None,
                                    C.byref(root), C.byref(depth), C.byref(leaves), C.byref(ms)))
    return {"nodes": nodes[: n.value], "order": order, "root": root.value, "depth": depth.value, "leaves": leaves.value,
            "build_ms": ms.value}


class Scene:
    """A flattened scene + reference-order BVH resident in HBM on one GPU (``cge_scene``)."""

    def __init__(self, flat: FlatScene, device: int = 0, use_stored_bvh: bool = False):
        self.flat = flat
        d, self._keep = scene_desc(flat)
        if use_stored_bvh and len(flat.bvh_nodes):
            nodes = np.ascontiguousarray(flat.bvh_nodes)
            order = np.ascontiguousarray(flat.bvh_prim_order)
            self._keep += [nodes, order]
            d.n_bvh_nodes, d.bvh_root = len(nodes), flat.bvh_root
            d.bvh_nodes, d.bvh_prim_order = nodes.ctypes.data, order.ctypes.data
        self.handle = C.c_void_p()
        _check(lib().cge_scene_create(C.byref(d), device, C.byref(self.handle)))
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            lib().cge_scene_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh_info(self) -> dict:
        v = [C.c_uint32() for _ in range(4)]
        _check(lib().cge_scene_bvh_info(self.handle, *[C.byref(x) for x in v]))
        return dict(zip(("nodes", "levels", "leaves", "max_leaf_prims"), (x.value for x in v)))

    def bvh_export(self):
        info = self.bvh_info()
        nodes = np.zeros(info["nodes"], scenefile.BVH_NODE_DT)
        order = np.zeros(self.flat.n_primitives, "<u4")
        _check(lib().cge_scene_bvh_export(self.handle, _p(nodes), _p(order)))
        return nodes, order

    def update_lights(self, lights: np.ndarray):
        lights = np.ascontiguousarray(lights, dtype=scenefile.LIGHT_DT)
        _check(lib().cge_scene_update_lights(self.handle, _p(lights) if len(lights) else None, len(lights)))

    def render(self, cfg: dict, traversal: int = TRAVERSAL_FAST, want_ids: bool = True, rgb_out=None, ids_out=None,
               part=(0, 1), camera: CgeCamera | None = None, flags: int = 0):
        """cge_render into host arrays.  Returns (rgb[H,W,3], ids[H,W] or None, stats dict)."""
        cam = camera or camera_from_cfg(cfg)
        p = params_from_cfg(cfg, traversal, want_ids, part, flags)
        H, W = cfg["height"], cfg["width"]
        rgb = rgb_out if rgb_out is not None else np.zeros((H, W, 3), np.float32)
        ids = (ids_out if ids_out is not None else np.full((H, W), -1, np.int32)) if want_ids else None
        st = CgeStats()
        _check(lib().cge_render(self.handle, C.byref(cam), C.byref(p), _p(rgb), _p(ids), C.byref(st)))
        return rgb, ids, st.as_dict()

    def render_rgba8(self, cfg: dict, out=None, traversal: int = TRAVERSAL_FAST, camera: CgeCamera | None = None):
        """cge_render with CGE_FLAG_OUTPUT_RGBA8: the frame after the reference's bitmap output stage, (H, W, 4) uint8."""
        cam = camera or camera_from_cfg(cfg)
        p = params_from_cfg(cfg, traversal, False, (0, 1), FLAG_OUTPUT_RGBA8)
        H, W = cfg["height"], cfg["width"]
        rgba = out if out is not None else np.zeros((H, W, 4), np.uint8)
        st = CgeStats()
        _check(lib().cge_render(self.handle, C.byref(cam), C.byref(p), _p(rgba), None, C.byref(st)))
        return rgba, st.as_dict()

    def render_device(self, cfg: dict, rgb_ptr: int, ids_ptr: int = 0, traversal: int = TRAVERSAL_FAST,
                      camera: CgeCamera | None = None, part=(0, 1), flags: int = 0) -> dict:
        """cge_render writing to DEVICE pointers (no D2H): the kernel-only / HBM-resident measurement."""
        cam = camera or camera_from_cfg(cfg)
        p = params_from_cfg(cfg, traversal, bool(ids_ptr), part, flags)
        p.flags |= FLAG_RGB_DEVICE_PTR
        st = CgeStats()
        _check(lib().cge_render(self.handle, C.byref(cam), C.byref(p), C.c_void_p(rgb_ptr),
                                C.c_void_p(ids_ptr) if ids_ptr else None, C.byref(st)))
        return st.as_dict()

    def trace_rays(self, rays7, cfg: dict, traversal: int = TRAVERSAL_FAST):
        rays7 = np.ascontiguousarray(rays7, np.float32)
        n = rays7.shape[0]
        p = params_from_cfg(cfg, traversal, True)
        rgb = np.zeros((n, 3), np.float32)
        ids = np.full(n, -1, np.int32)
        _check(lib().cge_trace_rays(self.handle, _p(rays7), n, C.byref(p), _p(rgb), _p(ids)))
        return rgb, ids


class Comm:
    """NCCL communicator for the framebuffer tile gather (one process per GPU)."""

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self.handle = C.c_void_p()
        _check(lib().cge_comm_create(buf, rank, n_ranks, device, C.byref(self.handle)))
        self.rank, self.n_ranks = rank, n_ranks

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
        _check(lib().cge_comm_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.handle:
            lib().cge_comm_destroy(self.handle)
            self.handle = None

    def host_frame(self, shape, dtype=np.float32) -> np.ndarray:
        """Collective: a frame in host memory every rank maps (cge_comm_host_frame), for ``render(..., shared_frame=...)``."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        _check(lib().cge_comm_host_frame(self.handle, nbytes, C.byref(ptr)))
        buf = (C.c_char * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def peer_frame(self, nbytes: int) -> int:
        """Collective: a device frame on rank 0 that every rank's kernels can store into (cge_comm_peer_frame); returns this
        process's device pointer to it, for ``render(..., peer_frame=ptr)``."""
        ptr = C.c_void_p()
        _check(lib().cge_comm_peer_frame(self.handle, int(nbytes), C.byref(ptr)))
        return int(ptr.value)

    def render(self, scene: Scene, cfg: dict, traversal: int = TRAVERSAL_FAST, want_ids: bool = False, rgb_out=None,
               ids_out=None, device_ptrs=None, camera: CgeCamera | None = None, shared_frame=None, shared_ids=None, flags: int = 0,
               peer_frame: int | None = None):
        """cge_render_distributed.  shared_frame: a Comm.host_frame array every rank passes (CGE_FLAG_SHARED_HOST_FRAME);
        peer_frame: this rank's pointer from Comm.peer_frame (CGE_FLAG_PEER_FRAME)."""
        cam = camera or camera_from_cfg(cfg)
        p = params_from_cfg(cfg, traversal, want_ids, (0, 1), flags)
        H, W = cfg["height"], cfg["width"]
        st = CgeStats()
        if peer_frame is not None:
            p.flags |= FLAG_PEER_FRAME
            _check(lib().cge_render_distributed(scene.handle, self.handle, C.byref(cam), C.byref(p), C.c_void_p(peer_frame), None,
                                                C.byref(st)))
            return None, None, st.as_dict()
        if device_ptrs is not None:
            p.flags |= FLAG_RGB_DEVICE_PTR
            rgb_ptr, ids_ptr = device_ptrs
            _check(lib().cge_render_distributed(scene.handle, self.handle, C.byref(cam), C.byref(p), C.c_void_p(rgb_ptr),
                                                C.c_void_p(ids_ptr) if ids_ptr else None, C.byref(st)))
            return None, None, st.as_dict()
        if shared_frame is not None:
            p.flags |= FLAG_SHARED_HOST_FRAME
            _check(lib().cge_render_distributed(scene.handle, self.handle, C.byref(cam), C.byref(p), _p(shared_frame),
                                                _p(shared_ids) if want_ids else None, C.byref(st)))
            return shared_frame, shared_ids, st.as_dict()
        rgb = ids = None
        if self.rank == 0:
            rgb = rgb_out if rgb_out is not None else np.zeros((H, W, 3), np.float32)
            if want_ids:
                ids = ids_out if ids_out is not None else np.full((H, W), -1, np.int32)
        _check(lib().cge_render_distributed(scene.handle, self.handle, C.byref(cam), C.byref(p), _p(rgb), _p(ids), C.byref(st)))
        return rgb, ids, st.as_dict()


def hull_clear_host(o3, light9, tri9) -> np.ndarray:
    """cge_hull_clear_host: the light-hull pre-pass's "no ray of the hull can hit this triangle" test, on the host (tests)."""
    o3, light9, tri9 = (np.ascontiguousarray(a, np.float32) for a in (o3, light9, tri9))
    out = np.zeros(len(o3), np.int32)
    l = lib()
    l.cge_hull_clear_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    _check(l.cge_hull_clear_host(o3.ctypes.data, light9.ctypes.data, tri9.ctypes.data, len(o3), out.ctypes.data))
    return out


def hull_box_host(o3, light9, box6) -> np.ndarray:
    """cge_hull_box_host: the light-hull pre-pass's box test on the host (tests): 0 = no ray of the hull passes through the box."""
    o3, light9, box6 = (np.ascontiguousarray(a, np.float32) for a in (o3, light9, box6))
    out = np.zeros(len(o3), np.int32)
    l = lib()
    l.cge_hull_box_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    _check(l.cge_hull_box_host(o3.ctypes.data, light9.ctypes.data, box6.ctypes.data, len(o3), out.ctypes.data))
    return out


def device_to_host(ptr: int, shape, dtype=np.float32) -> np.ndarray:
    """Copy of device memory the library owns (e.g. rank 0's Comm.peer_frame) through the CUDA runtime (tests only; no torch)."""
    out = np.empty(shape, dtype)
    rt = None
    for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "libcudart.so"):
        try:
            rt = C.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        raise OSError("libcudart not found")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rc = rt.cudaMemcpy(out.ctypes.data, C.c_void_p(int(ptr)), out.nbytes, 2)  # cudaMemcpyDeviceToHost
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy failed: {rc}")
    return out


def device_view(ptr: int, shape, typestr: str = "<f4"):
    """A torch tensor over device memory the library owns (e.g. rank 0's Comm.peer_frame), without a copy (tests / bench only)."""
    import torch

    class _View:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_View(), device="cuda")


# ---- device KATs (libIntersect functions evaluated on the GPU) ---------------------------------------------
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def kat_triangle(v9, ray7, device=0, precomputed=False):
    v9, ray7 = _f32(v9), _f32(ray7).copy()
    hit = np.zeros(len(v9), np.int32)
    fn = lib().cge_kat_triangle_precomputed if precomputed else lib().cge_kat_triangle
    _check(fn(_p(v9), _p(ray7), _p(hit), len(v9), device))
    return hit, ray7[:, 6].copy()


def kat_aabb(b6, ray7, device=0):
    b6, ray7 = _f32(b6), _f32(ray7).copy()
    hit = np.zeros(len(b6), np.int32)
    _check(lib().cge_kat_aabb(_p(b6), _p(ray7), _p(hit), len(b6), device))
    return hit, ray7[:, 6].copy()


def kat_plane(p4, ray7, device=0):
    p4, ray7 = _f32(p4), _f32(ray7).copy()
    hit = np.zeros(len(p4), np.int32)
    _check(lib().cge_kat_plane(_p(p4), _p(ray7), _p(hit), len(p4), device))
    return hit, ray7[:, 6].copy()


def kat_sphere(s4, ray7, device=0):
    s4, ray7 = _f32(s4), _f32(ray7).copy()
    hit = np.zeros(len(s4), np.int32)
    nrm = np.zeros((len(s4), 3), np.float32)
    _check(lib().cge_kat_sphere(_p(s4), _p(ray7), _p(nrm), _p(hit), len(s4), device))
    return hit, ray7[:, 6].copy(), nrm


def kat_triangle_plane(v9, device=0):
    v9 = _f32(v9)
    out = np.zeros((len(v9), 4), np.float32)
    _check(lib().cge_kat_triangle_plane(_p(v9), _p(out), len(v9), device))
    return out


def kat_point_in_triangle(v9, n3, p3, device=0):
    v9, n3, p3 = _f32(v9), _f32(n3), _f32(p3)
    out = np.zeros(len(v9), np.int32)
    _check(lib().cge_kat_point_in_triangle(_p(v9), _p(n3), _p(p3), _p(out), len(v9), device))
    return out


def partition_tiles(width: int, height: int, part_index: int, part_count: int, tile_rows: bool = False) -> list:
    """Tile ids (row-major 8x4 tiles) that cge_render renders for part_index / part_count — the host-side statement of
    the multi-GPU image partition (csrc/dev_scene.h part_tile_of, csrc/cge_api.cu tiles_of): the parts are dealt units of
    consecutive tiles round robin, a unit being one tile, or with tile_rows (CGE_FLAG_PARTITION_TILE_ROWS, what
    cge_render_distributed uses) one whole row of tiles."""
    tiles_x, tiles_y = (width + 7) // 8, (height + 3) // 4
    n_tiles = tiles_x * tiles_y
    if part_count <= 1:
        return list(range(n_tiles))
    unit = tiles_x if tile_rows else 1
    n_units = (n_tiles + unit - 1) // unit
    return [t for u in range(part_index, n_units, part_count) for t in range(u * unit, min((u + 1) * unit, n_tiles))]


def dynamic_tile_plan(width: int, height: int, n_ranks: int, pool_pct: int = 25, chunks_per_owner: int = 2):
    """Host-side statement of CGE_FLAG_DYNAMIC_TILES (csrc/cge_api.cu cge_render_distributed / grant_kernel): the first
    pool_pct % of every rank's tile rows (entries of its tile-row partition list) form the pool, cut into chunks_per_owner
    chunks per rank; chunk c belongs to rank c % n_ranks.  Returns (static tile lists per rank, the tile list of every
    chunk in grant order); a chunk can be empty."""
    tiles_x, tiles_y = (width + 7) // 8, (height + 3) // 4
    rows = [(tiles_y - r + n_ranks - 1) // n_ranks if tiles_y > r else 0 for r in range(n_ranks)]
    pool = [rw * pool_pct // 100 for rw in rows]
    chunk_rows = max(1, -(-pool[0] // chunks_per_owner)) if pool else 1
    lists = [partition_tiles(width, height, r, n_ranks, tile_rows=True) for r in range(n_ranks)]
    if n_ranks == 1:
        lists = [list(range(tiles_x * tiles_y))]
    static = [lists[r][pool[r] * tiles_x:] for r in range(n_ranks)]
    chunks = []
    for c in range(n_ranks * chunks_per_owner):
        owner, j = c % n_ranks, c // n_ranks
        first = j * chunk_rows
        count = min(chunk_rows, pool[owner] - first) if first < pool[owner] else 0
        chunks.append(lists[owner][first * tiles_x:(first + count) * tiles_x])
    return static, chunks


def load_scene(cfg: dict) -> FlatScene:
    """Flat scene for a config: a committed .cges fixture, or the procedurally generated dragon stand-in."""
    if cfg["scene"].startswith("standin:"):
        return standin.make(cfg["scene"].split(":", 1)[1])
    return scenefile.load(configs.scene_path(cfg))
