"""-m gpu: the reference engine with the GPU path behind its OWN seam.

oracle/_ref/libcge_ref_gpu.so is the unmodified reference (Scene loader, Trackball, Window stub, Screen, BvhInterface, the
CPU renderer) linked with computer-graphics-engine_b200/host/render_gpu.cpp — the file INTEGRATION.md tells a maintainer to
add — compiled against the reference's own headers.  In that library

    renderRayTracing(const Scene&, const Trackball&, const BvhInterface&, Screen&, const Features&)      (src/render.h:32)

is the shim: reference objects go in, libcge.so renders, the image lands in Screen::pixels().  The reference's former body
stays linked as renderRayTracingCPU (the shim's fallback).  The same call on libcge_ref.so is the reference's CPU renderer.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import compare_images  # noqa: E402


@pytest.fixture(scope="module")
def gpu_ref(ref):
    if not ref.available("gpu"):
        pytest.skip("oracle/_ref/libcge_ref_gpu.so not built (needs /root/reference and libcge.so at build time)")
    return ref


def _render(ref, cfg, path, which, use_rrt=True):
    with ref.RefScene(path, cfg["features"], plain=which) as rs:
        rgb, _, st = rs.render(cfg, want_ids=False, use_render_ray_tracing=use_rrt)
    return rgb, st


@pytest.mark.parametrize("name,size", [("c1_cornell", (256, 256)), ("c2_cube_textured", (320, 180)), ("c4_monkey_mirror", (256, 144))])
def test_reference_objects_in_gpu_image_out(cge, gpu_ref, name, size):
    """The reference's own renderRayTracing on both sides (depth literal 5, src/render.cpp:318): CPU body vs GPU shim."""
    cfg = cge.configs.get(name, *size)
    path = cge.configs.scene_path(cfg)
    cpu, _ = _render(gpu_ref, cfg, path, False)
    gpu, st = _render(gpu_ref, cfg, path, "gpu")
    err, nan_mm = compare_images(gpu, cpu)
    bad = (np.abs(np.nan_to_num(gpu, nan=0.0) - np.nan_to_num(cpu, nan=0.0)).max(-1) > 1e-3).mean()
    assert nan_mm <= 2 and bad <= 1e-4, (name, err, nan_mm, bad)
    assert np.isfinite(np.nan_to_num(gpu, nan=0.0)).all() and float(np.nan_to_num(gpu, nan=0.0).max()) > 0.05  # a real image


def test_soft_shadows_through_the_shim(cge, gpu_ref):
    """Area light: the reference draws its jitter from rand(); with rand wrapped to the stateless sampler on the CPU side (the
    harness's pixel loop sets the pixel / draw counter, depth 5 as renderRayTracing's literal) the two images agree."""
    cfg = dict(cge.configs.get("c3_teapot_soft", 192, 108), ray_depth=5)
    path = cge.configs.scene_path(cfg)
    cpu, _ = _render(gpu_ref, cfg, path, False, use_rrt=False)
    gpu, _ = _render(gpu_ref, cfg, path, "gpu")
    err, nan_mm = compare_images(gpu, cpu)
    assert nan_mm == 0 and err <= 1e-3, (err, nan_mm)


def test_shim_falls_back_to_the_reference_cpu_body(cge, gpu_ref):
    """An ExtraFeatures flag libcge.so does not implement (enableBvhSahBinning: it only changes how the reference builds its tree)
    makes cge_render answer CGE_ERR_UNSUPPORTED; the shim then calls renderRayTracingCPU - the frame is the reference's."""
    cfg = cge.configs.get("c1_cornell", 96, 96)
    cfg["features"] |= 1 << 17
    path = cge.configs.scene_path(cfg)
    cpu, _ = _render(gpu_ref, cfg, path, False)
    gpu, _ = _render(gpu_ref, cfg, path, "gpu")
    assert gpu.tobytes() == cpu.tobytes()


def test_light_edit_between_frames_is_picked_up(cge, gpu_ref):
    """The shim re-sends the light list only when it changed (the GUI edits lights between frames, src/main.cpp:290-368)."""
    a = cge.configs.get("c1_cornell", 128, 128)
    path = cge.configs.scene_path(a)
    with gpu_ref.RefScene(path, a["features"], plain="gpu") as rs:
        first, _, _ = rs.render(a, want_ids=False, use_render_ray_tracing=True)
        again, _, _ = rs.render(a, want_ids=False, use_render_ray_tracing=True)
    assert first.tobytes() == again.tobytes()
