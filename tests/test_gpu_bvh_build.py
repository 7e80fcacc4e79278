"""-m gpu: the GPU SAH builder (csrc/bvh_sah_gpu.cu) against the host builder (csrc/bvh_sah.cpp): the two
must return the SAME tree - nodes, boxes, child references, primitive order - bit for bit, on every fixture scene, on the
868 334-triangle stand-in, and on degenerate inputs (coincident centroids, a single triangle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def assert_same_tree(a, b):
    assert (a["root"], a["depth"], a["leaves"]) == (b["root"], b["depth"], b["leaves"])
    assert np.array_equal(a["order"], b["order"])
    assert len(a["nodes"]) == len(b["nodes"])
    assert a["nodes"].tobytes() == b["nodes"].tobytes()


def check_tree_is_valid(t, n_prims):
    """Every primitive in exactly one leaf, <= 4 per leaf, pre-order numbering, child boxes inside the parent's."""
    assert sorted(t["order"].tolist()) == list(range(n_prims))
    seen = np.zeros(n_prims, bool)
    stack = [t["root"]]
    visited_inner = 0
    while stack:
        ref = stack.pop()
        if ref & 0x80000000:
            first, count = ref & 0x0FFFFFFF, ((ref >> 28) & 7) + 1
            assert count <= 4 and not seen[first:first + count].any()
            seen[first:first + count] = True
        else:
            visited_inner += 1
            nd = t["nodes"][ref]
            if not int(nd["left"]) & 0x80000000:
                assert int(nd["left"]) == ref + 1  # pre-order: the left subtree follows its parent
            stack += [int(nd["right"]), int(nd["left"])]
    assert seen.all() and visited_inner == len(t["nodes"])


@pytest.mark.parametrize("scene", ["cornell", "cube_textured", "teapot_area", "monkey_mirror", "triangle", "cube", "teapot", "monkey"])
def test_gpu_tree_equals_host_tree(cge, scene):
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / f"{scene}.cges")
    host = cge.build_fast_bvh(flat, on_gpu=False)
    gpu = cge.build_fast_bvh(flat, on_gpu=True)
    assert_same_tree(gpu, host)
    check_tree_is_valid(gpu, flat.n_primitives)


def test_gpu_tree_equals_host_tree_dragon_standin(cge):
    flat = cge.standin.make("dragon")
    host = cge.build_fast_bvh(flat, on_gpu=False)
    gpu = cge.build_fast_bvh(flat, on_gpu=True)
    assert_same_tree(gpu, host)
    assert len(gpu["nodes"]) > 400_000 and gpu["build_ms"] > 0
    print(f"GPU SAH build of {flat.n_primitives} triangles: {gpu['build_ms']:.2f} ms, {len(gpu['nodes'])} inner nodes, depth {gpu['depth']}")


def test_degenerate_inputs(cge):
    """Coincident centroids force the cut-in-the-middle rule; duplicated triangles tie every bin."""
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / "triangle.cges")
    many = cge.scenefile.FlatScene()
    for _ in range(37):
        many.append_meshes(flat)
    many.set_lights([])
    host = cge.build_fast_bvh(many, on_gpu=False)
    gpu = cge.build_fast_bvh(many, on_gpu=True)
    assert_same_tree(gpu, host)
    check_tree_is_valid(gpu, many.n_primitives)
    assert len(gpu["nodes"]) > 0  # 37 coincident triangles cannot sit in one leaf


def test_scene_uses_gpu_built_tree_and_renders_identically(cge, monkeypatch):
    cfg = cge.configs.get("c4_monkey_mirror", 256, 144)
    flat = cge.load_scene(cfg)
    monkeypatch.setenv("CGE_SAH_BUILD", "host")
    with cge.Scene(flat) as sc:
        rgb_h, ids_h, st_h = sc.render(cfg, traversal=1)
    monkeypatch.delenv("CGE_SAH_BUILD")
    with cge.Scene(flat) as sc:
        rgb_g, ids_g, st_g = sc.render(cfg, traversal=1)
    assert rgb_h.tobytes() == rgb_g.tobytes() and np.array_equal(ids_h, ids_g)
    assert st_h["gpu_rays"] == st_g["gpu_rays"]
