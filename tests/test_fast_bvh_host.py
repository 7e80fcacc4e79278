"""CPU tests (-m "not gpu"): the host builder of the FAST traversal tree (csrc/bvh_sah.cpp, the checker of the GPU builder):
structural invariants, determinism, and the spec's degenerate rules (csrc/sah_split.h)."""
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("computer-graphics-engine_b200")


def walk(t):
    """(leaf ranges, inner nodes visited, max depth) by an explicit pre-order walk."""
    leaves, inner, depth = [], 0, 0
    stack = [(t["root"], 1)]
    while stack:
        ref, d = stack.pop()
        depth = max(depth, d)
        if ref & 0x80000000:
            leaves.append((ref & 0x0FFFFFFF, ((ref >> 28) & 7) + 1))
        else:
            inner += 1
            nd = t["nodes"][ref]
            stack += [(int(nd["right"]), d + 1), (int(nd["left"]), d + 1)]
    return leaves, inner, depth


@pytest.mark.parametrize("scene", ["cornell", "cube_textured", "teapot_area", "monkey_mirror", "triangle"])
def test_host_tree_invariants(scene):
    flat = pkg.scenefile.load(pkg.configs.SCENE_DIR / f"{scene}.cges")
    t = pkg.build_fast_bvh(flat, on_gpu=False)
    n = flat.n_primitives
    leaves, inner, depth = walk(t)
    assert inner == len(t["nodes"]) and len(leaves) == t["leaves"] and depth == t["depth"]
    covered = sorted(leaves)
    assert covered[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(covered, covered[1:])) and sum(c for _, c in covered) == n
    assert max(c for _, c in leaves) <= 4
    assert sorted(t["order"].tolist()) == list(range(n))
    # child boxes contain the primitives below them: check the root's children against the vertex cloud
    if inner:
        v = flat.vertices["position"]
        root = t["nodes"][0]
        lo = np.minimum(root["left_lower"], root["right_lower"])
        hi = np.maximum(root["left_upper"], root["right_upper"])
        used = np.unique(np.concatenate([flat.triangles[m["triangle_offset"]:m["triangle_offset"] + m["triangle_count"]].ravel() + m["vertex_offset"]
                                         for m in flat.meshes]))
        assert (v[used] >= lo).all() and (v[used] <= hi).all()
    again = pkg.build_fast_bvh(flat, on_gpu=False)
    assert again["nodes"].tobytes() == t["nodes"].tobytes() and np.array_equal(again["order"], t["order"])


def test_coincident_centroids_cut_in_the_middle():
    flat = pkg.scenefile.load(pkg.configs.SCENE_DIR / "triangle.cges")
    many = pkg.scenefile.FlatScene()
    for _ in range(9):
        many.append_meshes(flat)
    t = pkg.build_fast_bvh(many, on_gpu=False)
    leaves, inner, _ = walk(t)
    # 9 identical triangles: no axis has a centroid extent, so ranges are halved without reordering until <= 4 remain
    assert np.array_equal(t["order"], np.arange(9))
    assert sorted(c for _, c in leaves) == [2, 3, 4] and inner == 2  # 9 -> 4 | 5, 5 -> 2 | 3


def test_gpu_builder_refuses_without_device():
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    flat = pkg.scenefile.load(pkg.configs.SCENE_DIR / "cornell.cges")
    with pytest.raises(pkg.CgeError) as e:
        pkg.build_fast_bvh(flat, on_gpu=True)
    assert e.value.code == pkg.ERR_CUDA
