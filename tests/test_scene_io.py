"""CPU: the mirrored scene loaders (host/cge_scene_io.hpp: OBJ + MTL + PNG -> Scene, reference framework/src/mesh.cpp:52-176,
framework/src/image.cpp:12-35, src/scene.cpp:5-103) against the fixtures that the UNMODIFIED reference loaders exported
(tests/golden/scenes/*.cges, tests/golden/make_scenes.py): the flattened scene must be the same file, byte for byte - vertex
de-duplication order, quad splitting, centring arithmetic, material defaults and texels included.  Needs the reference's data
directory (present in the build container only)."""
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
DATA = Path(os.environ.get("CGE_REFERENCE_DIR", "/root/reference")) / "data"
TOOL = ROOT / "tests" / "cpp" / "scene_export"

pytestmark = pytest.mark.skipif(not DATA.is_dir(), reason="the reference's data directory is not available on this machine")


@pytest.fixture(scope="module")
def tool():
    subprocess.run(["bash", str(ROOT / "tests" / "cpp" / "build.sh")], check=True)
    return TOOL


# SceneType (reference src/scene.h:15-26) -> fixture exported by the reference's own loadScenePrebuilt
PREBUILT = [(0, "single_triangle", "triangle"), (1, "cube", "cube"), (2, "cube-textured", "cube_textured"), (3, "CornellBox", "cornell"),
            (4, "cornell_box_parallelogram_light", "cornell_parallelogram"), (5, "Monkey", "monkey"), (6, "teapot", "teapot"),
            (8, "spheres", "spheres"), (9, "custom", "custom")]  # custom: a textured quad whose material has no Kd (loader default 0.6)


@pytest.mark.parametrize("number,name,fixture", PREBUILT)
def test_prebuilt_scene_equals_the_reference_loaders_output(tool, tmp_path, number, name, fixture):
    folder = "loader" if fixture == "custom" else "scenes"  # (fixtures only this test uses live beside the renderer's scenes)
    want = (ROOT / "tests" / "golden" / folder / f"{fixture}.cges").read_bytes()
    for what in (str(number), name):  # by SceneType number and by the names src/config.cpp:404-431 accepts
        out = tmp_path / f"{fixture}_{what}.cges"
        r = subprocess.run([str(tool), "prebuilt", what, str(DATA), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == want, f"{fixture} via {what}"


def test_quad_mesh_equals_the_reference_loaders_output(tool, tmp_path):
    """monkey-rotated-quad.obj: 468 quads + 32 triangles, centred and normalised (fixture: the reference's loadMesh(file, true))."""
    out = tmp_path / "mq.cges"
    r = subprocess.run([str(tool), "obj", str(DATA / "monkey-rotated-quad.obj"), "1", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == (ROOT / "tests" / "golden" / "loader" / "monkey_quad.cges").read_bytes()


def test_loader_details(tool, tmp_path):
    """Quad splitting along the shorter diagonal, negative indices, vertex sharing, a material library that is missing, a face
    whose last triangle changes material (it closes the previous run: framework/src/mesh.cpp:78-85), and what is refused."""
    import importlib
    import sys
    import numpy as np
    sys.path.insert(0, str(ROOT))
    sf = importlib.import_module("computer-graphics-engine_b200.scenefile")
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nNs 12.5\nTr 0.25\nnewmtl blue\nKd 0 0 1\nKs 0.5 0.5 0.5\nd 0.5\nTr 0.9\n")
    (tmp_path / "q.obj").write_text(
        "mtllib missing.mtl m.mtl\n"
        "v 0 0 0\nv 2 0 0\nv 2 1 0\nv 0 1 0\nv 0 0 1\nv 1e0 0 1\nv .5 5.0E-1 1\n"
        "usemtl red\nf 1 2 3 4\n"            # diagonals 0-2 and 1-3 are equally long -> split along 1-3
        "f -3 -2 -1\n"
        "usemtl blue\nf 1 2 5\n")
    out = tmp_path / "q.cges"
    r = subprocess.run([str(tool), "obj", str(tmp_path / "q.obj"), "0", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s = sf.load(out)
    # one shape; the last triangle (blue) closes the run of the red ones: ONE mesh with the red material
    assert len(s.meshes) == 1 and s.meshes["triangle_count"][0] == 4
    assert np.allclose(s.meshes["kd"][0], (1, 0, 0)) and s.meshes["shininess"][0] == np.float32(12.5)
    assert s.meshes["transparency"][0] == np.float32(0.75)  # Tr without d: 1 - Tr
    assert s.triangles[:2].tolist() == [[0, 1, 2], [1, 3, 2]]  # (0,1,3), (1,2,3) in first-use vertex numbering
    assert s.vertices["position"][4:7].tolist() == [[0, 0, 1], [1, 0, 1], [0.5, 0.5, 1]]  # "-3 -2 -1", "1e0", ".5", "5.0E-1"
    assert np.allclose(s.vertices["normal"][:7], (0, 0, 1))  # no vn in the file: the geometric normal
    # a vertex is (position, normal, texcoord): the last face reuses two positions of the quad under another normal -> new vertices
    assert len(s.vertices) == 10 and s.triangles[3].tolist() == [7, 8, 9] and np.allclose(s.vertices["normal"][7:], (0, -1, 0))
    assert s.vertices["position"][7:9].tolist() == [[0, 0, 0], [2, 0, 0]]
    (tmp_path / "p.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0.5 1.5 0\nv 0 1 0\nf 1 2 3 4 5\n")
    r = subprocess.run([str(tool), "obj", str(tmp_path / "p.obj"), "0", str(tmp_path / "p.cges")], capture_output=True, text=True)
    assert r.returncode == 1 and "more than four corners" in r.stderr
    r = subprocess.run([str(tool), "obj", str(tmp_path / "nope.obj"), "0", str(tmp_path / "n.cges")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr


def test_cli_resolves_the_reference_scene_names(tool, tmp_path):
    cli = ROOT / "tests" / "cpp" / "cge_cli"
    for scene, shown in (('"CornellBox"', "cornell_box (built-in)"), ("5", "monkey (built-in)"), ('"cube-textured"', "cube_textured (built-in)"),
                         ('"monkey.obj"', "monkey.obj"), ('"teapot.cges"', "teapot.cges")):
        cfg = tmp_path / "c.toml"
        cfg.write_text(f'window_size = [64, 48]\ndata_path = "{DATA}"\nscene = {scene}\n')
        r = subprocess.run([str(cli), str(cfg), "--print-config"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines()[1] == f"scene {shown}"


def test_flat_scene_files_are_untrusted_input(tool, tmp_path):
    """loadFlatScene (C++ host mirror) and scenefile.load (Python glue): a truncated file, counts larger than the file, offsets and
    ids that point outside their arrays are refused with a clear error instead of being copied or indexed."""
    import importlib
    import sys
    import numpy as np
    sys.path.insert(0, str(ROOT))
    sf = importlib.import_module("computer-graphics-engine_b200.scenefile")
    good = (ROOT / "tests" / "golden" / "scenes" / "cube_textured.cges").read_bytes()

    def check(data):
        f = tmp_path / "x.cges"
        f.write_bytes(data)
        return subprocess.run([str(tool), "check", str(f)], capture_output=True, text=True)

    assert check(good).returncode == 0
    r = check(good[: len(good) // 2])
    assert r.returncode == 1 and "truncated" in r.stderr
    with pytest.raises(ValueError, match="truncated"):
        f = tmp_path / "t.cges"
        f.write_bytes(good[: len(good) // 2])
        sf.load(f)
    with pytest.raises(ValueError, match="not a flat scene file"):
        f.write_bytes(b"short")
        sf.load(f)
    s = sf.load(ROOT / "tests" / "golden" / "scenes" / "cube_textured.cges")
    for field, value, what in (("vertex_offset", 10**6, "mesh vertices"), ("vertex_count", 10**6, "mesh vertices"),
                               ("triangle_count", 10**6, "mesh triangles"), ("texture_id", 7, "mesh texture id")):
        bad = s.copy()
        bad.meshes[field][0] = value
        sf.save(bad, tmp_path / "b.cges")
        r = check((tmp_path / "b.cges").read_bytes())
        assert r.returncode == 1 and what in r.stderr, (field, r.stderr)
    bad = s.copy()
    bad.triangles[0, 1] = 10**6
    sf.save(bad, tmp_path / "b.cges")
    r = check((tmp_path / "b.cges").read_bytes())
    assert r.returncode == 1 and "triangle index" in r.stderr
    bad = s.copy()
    bad.textures["width"][0] = 1 << 20
    sf.save(bad, tmp_path / "b.cges")
    r = check((tmp_path / "b.cges").read_bytes())
    assert r.returncode == 1 and "texture" in r.stderr
    huge = bytearray(good)
    huge[12:16] = np.uint32(0xFFFFFFF0).tobytes()  # n_vertices far beyond the file: refused before anything is allocated
    r = check(bytes(huge))
    assert r.returncode == 1 and "truncated" in r.stderr
