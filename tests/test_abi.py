"""CPU-only checks of the drop-in boundary: libcge.so loads, exports every symbol include/cge.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "cge.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cge_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cge):
    lib = cge.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libcge.so does not export {n}"
    assert set(names) == set(cge.ABI_SYMBOLS)
    assert lib.cge_abi_version() == 4


def test_struct_sizes_match_header(cge):
    # sizes the C header implies (checked against the ctypes mirrors used by the tests and bench)
    assert C.sizeof(cge.CgeCamera) == 36
    assert C.sizeof(cge.CgeParams) == 64
    assert C.sizeof(cge.CgeStats) == 104  # 7 x u64, 2 x f32, u32, 4 x f32 (= 84, padded to 88), u64, f32 (padded to 104)
    assert C.sizeof(cge.CgeSceneDesc) == 32 + 7 * 8 + 8 + 2 * 8


def test_flag_constants_match_header(cge):
    """The ctypes glue's flag values are the header's (include/cge.h is what a maintainer binds; the glue must not drift)."""
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "cge.h").read_text(), flags=re.S)
    header = {name: 1 << int(shift) for name, shift in re.findall(r"\b(CGE_(?:DEV_)?FLAG_[A-Z0-9_]+)\s*=\s*1u\s*<<\s*(\d+)", text)}
    assert len(header) >= 10
    glue = {"CGE_FLAG_WANT_PRIM_IDS": 1, "CGE_FLAG_RGB_DEVICE_PTR": cge.FLAG_RGB_DEVICE_PTR, "CGE_FLAG_COUNT_TESTS": cge.FLAG_COUNT_TESTS,
            "CGE_FLAG_OUTPUT_RGBA8": cge.FLAG_OUTPUT_RGBA8, "CGE_FLAG_PARTITION_TILE_ROWS": cge.FLAG_PARTITION_TILE_ROWS,
            "CGE_FLAG_SHARED_HOST_FRAME": cge.FLAG_SHARED_HOST_FRAME, "CGE_FLAG_PEER_FRAME": cge.FLAG_PEER_FRAME,
            "CGE_FLAG_DYNAMIC_TILES": cge.FLAG_DYNAMIC_TILES, "CGE_DEV_FLAG_PER_THREAD": cge.FLAG_PER_THREAD,
            "CGE_DEV_FLAG_WAVEFRONT": cge.FLAG_WAVEFRONT, "CGE_DEV_FLAG_DEBUG_CYCLES": cge.FLAG_DEBUG_CYCLES}
    for name, value in glue.items():
        assert header[name] == value, name
    assert len(set(header.values())) == len(header)  # no two flags share a bit


def test_camera_matches_reference_trackball(cge, ref):
    for name in cge.configs.CONFIGS:
        cfg = cge.configs.get(name)
        a, b = cge.camera_from_cfg(cfg), ref.camera(cfg)
        assert bytes(a) == bytes(b), name


def test_no_gpu_means_loud_failure(cge):
    if cge.device_count() > 0:
        pytest.skip("a GPU is present")
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / "triangle.cges")
    with pytest.raises(cge.CgeError) as e:
        cge.Scene(flat)
    assert e.value.code == cge.ERR_CUDA
    with pytest.raises(cge.CgeError):
        cge.kat_aabb(np.zeros((1, 6), np.float32), np.zeros((1, 7), np.float32))


def test_product_never_touches_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = ROOT / "computer-graphics-engine_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.sh")):
        text = f.read_text()
        for line in text.splitlines():
            code = line.split("//")[0].split("#")[0] if f.suffix != ".py" else line.split("#")[0]
            assert "oracle/" not in code and "refharness" not in code and "libcge_ref" not in code, (f, line)


def test_host_bvh_builder_equals_reference_tree(cge):
    """The library's host-side BVH builder (no GPU involved) reproduces the reference's tree node for node and the same
    primitive permutation for every fixture scene (goldens exported from the reference's own BoundingVolumeHierarchy)."""
    g = np.load(ROOT / "tests" / "golden" / "reference_bvh.npz")
    for f in sorted((ROOT / "tests" / "golden" / "scenes").glob("*.cges")):
        flat = cge.scenefile.load(f)
        nodes, order, root, levels, leaves = cge.build_reference_bvh_host(flat)
        assert np.array_equal(order, g[f.stem + "_order"]), f.stem
        assert nodes.tobytes() == g[f.stem + "_nodes"].tobytes(), f.stem
        assert root == int(g[f.stem + "_root"])
        assert leaves == int((nodes["is_leaf"] == 1).sum()) and levels == int(nodes["depth"].max()) + 1
    # empty scene: no nodes, no error
    nodes, order, root, levels, leaves = cge.build_reference_bvh_host(cge.scenefile.FlatScene())
    assert len(nodes) == 0


def test_host_bvh_builder_dragon_standin_against_live_reference(cge, ref, tmp_path):
    flat = cge.standin.make("dragon", n=64)  # 49 154 triangles: exercises the MAX_DEPTH-16 multi-primitive leaves
    path = tmp_path / "s.cges"
    cge.scenefile.save(flat, path)
    with ref.RefScene(path, cge.configs.FEAT_ACCEL_STRUCTURE) as rs:
        out = tmp_path / "with_bvh.cges"
        rs.write_with_bvh(out)
    want = cge.scenefile.load(out)
    nodes, order, root, levels, leaves = cge.build_reference_bvh_host(flat)
    assert np.array_equal(order, want.bvh_prim_order) and nodes.tobytes() == want.bvh_nodes.tobytes() and root == want.bvh_root
    assert levels == 16


def test_supplied_bvh_is_validated_as_untrusted_input(cge):
    """A caller-supplied tree (stored in a flat scene file) is walked from the root before anything trusts it: cycles, shared
    subtrees, overlapping or missing leaves and ranges that do not split are refused, and the depth that sizes the device
    traversal stack is the measured one, whatever the depth fields claim."""
    flat = cge.scenefile.load(cge.configs.SCENE_DIR / "cornell.cges")
    nodes, order, root, levels, leaves = cge.build_reference_bvh_host(flat)
    assert cge.validate_bvh(flat, nodes, order, root) == (levels, leaves)
    lied = nodes.copy()
    lied["depth"] = 0  # understated depth: the walk measures the real one
    assert cge.validate_bvh(flat, lied, order, root) == (levels, leaves)
    inner = [i for i in range(len(nodes)) if not nodes["is_leaf"][i]]
    leaf = [i for i in range(len(nodes)) if nodes["is_leaf"][i]]

    def refused(bad_nodes, bad_order=order, bad_root=root):
        with pytest.raises(cge.CgeError) as e:
            cge.validate_bvh(flat, bad_nodes, bad_order, bad_root)
        assert e.value.code == cge.ERR_INVALID_ARG

    cyc = nodes.copy()  # a cycle: an inner node below the root points back at the root
    child = int(nodes["left"][root])
    assert not nodes["is_leaf"][child]
    cyc["left"][child] = root
    refused(cyc)
    shared = nodes.copy()  # both children are the same subtree
    shared["right"][root] = shared["left"][root]
    refused(shared)
    overlap = nodes.copy()  # a leaf range that overlaps its sibling's
    overlap["end"][leaf[0]] += 1
    refused(overlap)
    gap = nodes.copy()  # the root does not cover every primitive
    gap["end"][root] -= 1
    refused(gap)
    refused(nodes, bad_root=len(nodes))
    dup = order.copy()
    dup[0] = dup[1]
    refused(nodes, bad_order=dup)
    unreachable_leaf_as_root = leaf[0]  # a leaf as root covers one primitive only
    refused(nodes, bad_root=unreachable_leaf_as_root)
    assert inner
