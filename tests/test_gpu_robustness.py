"""GPU tests of the boundary's contracts that are not image parity: a light update racing frames in flight, the distributed
entry on a 1-rank communicator (same frame as cge_render, RGBA8 output included), partial frames to host memory."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_light_update_never_tears_a_frame_in_flight(cge):
    """cge_scene_update_lights while other threads render the scene (the reference's GUI edits lights every frame,
    src/main.cpp:290-368, and its CLI renders cameras from concurrent threads, :514-528).  The light COUNT changes between the two
    lists (1 area light = 16 samples per hit, 2 lights = 17), so a frame that mixed two versions would have wrong queue sizes
    and draw counters; every frame must equal the frame of one of the two lists bit for bit."""
    cfg = cge.configs.get("c3_teapot_soft", 256, 144)
    flat = cge.load_scene(cfg)
    lights_a = flat.lights.copy()
    extra = np.zeros(1, cge.scenefile.LIGHT_DT)
    extra["type"] = 0  # a point light
    extra["v"][0, :6] = [0.3, 0.9, 0.4, 0.5, 0.4, 0.3]
    lights_b = np.concatenate([flat.lights, extra])
    with cge.Scene(flat) as sc:
        want_a = sc.render(cfg, want_ids=False)[0].tobytes()
        sc.update_lights(lights_b)
        want_b = sc.render(cfg, want_ids=False)[0].tobytes()
        assert want_a != want_b
        stop = threading.Event()
        frames, errs = [], []

        def work():
            try:
                while not stop.is_set():
                    frames.append(sc.render(cfg, want_ids=False)[0].tobytes())
            except Exception as e:  # pragma: no cover
                errs.append(e)
        threads = [threading.Thread(target=work) for _ in range(3)]
        [t.start() for t in threads]
        for i in range(200):
            sc.update_lights(lights_a if i % 2 else lights_b)
        stop.set()
        [t.join() for t in threads]
        assert not errs and len(frames) >= 3
        bad = [f for f in frames if f not in (want_a, want_b)]
        assert not bad, f"{len(bad)} of {len(frames)} frames mixed two light lists"
        assert want_a in frames or want_b in frames


def _single_rank_comm(cge):
    return cge.Comm(cge.Comm.unique_id(), 0, 1, 0)


@pytest.mark.parametrize("name,size", [("c3_teapot_soft", (320, 180)), ("c4_monkey_mirror", (256, 144)), ("c1_cornell", (203, 117))])
def test_distributed_entry_on_one_rank_equals_cge_render(cge, name, size):
    cfg = cge.configs.get(name, *size)
    comm = _single_rank_comm(cge)
    try:
        with cge.Scene(cge.load_scene(cfg)) as sc:
            rgb, ids, _ = sc.render(cfg)
            d_rgb, d_ids, st = comm.render(sc, cfg, want_ids=True)
            assert d_rgb.tobytes() == rgb.tobytes() and np.array_equal(ids, d_ids)
            assert st["kernel_launches"] >= 1
            # the output stage of Screen::writeBitmapToFile after the gather: 4 bytes per pixel, not 12
            rgba, _ = sc.render_rgba8(cfg)
            guard = np.full((cfg["height"] * cfg["width"] * 4 + 64,), 0xAB, np.uint8)  # a 12-byte-per-pixel write would run over
            p = cge.params_from_cfg(cfg, cge.TRAVERSAL_FAST, False, (0, 1), cge.FLAG_OUTPUT_RGBA8)
            cam = cge.camera_from_cfg(cfg)
            stc = cge.CgeStats()
            rc = cge.lib().cge_render_distributed(sc.handle, comm.handle, C.byref(cam), C.byref(p), guard.ctypes.data, None, C.byref(stc))
            assert rc == 0, cge.lib().cge_last_error()
            assert (guard[-64:] == 0xAB).all()
            assert np.array_equal(guard[:-64].reshape(cfg["height"], cfg["width"], 4), rgba)
            p.flags |= cge.FLAG_RGB_DEVICE_PTR
            assert cge.lib().cge_render_distributed(sc.handle, comm.handle, C.byref(cam), C.byref(p), guard.ctypes.data, None,
                                                    C.byref(stc)) == cge.ERR_UNSUPPORTED
            # the peer frame (rank 0's device frame every rank's kernels store into; with one rank: the frame itself).  The N-rank
            # case needs N processes and N GPUs: tools/multi_gpu_check.py and bench.py compare its frame with the 1-GPU frame.
            H, W = cfg["height"], cfg["width"]
            pf = comm.peer_frame(H * W * 12)
            comm.render(sc, cfg, peer_frame=pf)
            assert cge.device_to_host(pf, (H, W, 3)).tobytes() == rgb.tobytes()
            with pytest.raises(cge.CgeError) as e:  # a pointer the communicator did not hand out
                comm.render(sc, cfg, peer_frame=pf + 256)
            assert e.value.code == cge.ERR_INVALID_ARG
            pw = cge.params_from_cfg(cfg, cge.TRAVERSAL_FAST, True, (0, 1), cge.FLAG_PEER_FRAME)  # no primitive ids in this mode
            assert cge.lib().cge_render_distributed(sc.handle, comm.handle, C.byref(cam), C.byref(pw), C.c_void_p(pf), None,
                                                    C.byref(stc)) == cge.ERR_UNSUPPORTED
    finally:
        comm.close()


def test_partition_to_host_memory_is_few_large_copies(cge):
    """A partition rendered into a host frame (part_count > 1): the tiles are packed on the device and leave in ONE copy
    (formerly one copy per 8-pixel tile row: 260 K copies for a 4K frame split two ways).  Pixels outside the partition stay
    untouched, the union of the partitions is the full frame, and a 4K-sized frame finishes in well under a second per part."""
    import time
    cfg = cge.configs.get("c4_monkey_mirror")
    H, W = cfg["height"], cfg["width"]
    with cge.Scene(cge.load_scene(cfg)) as sc:
        full, full_ids, _ = sc.render(cfg)
        rgb = np.full((H, W, 3), -7.0, np.float32)
        ids = np.full((H, W), -9, np.int32)
        sc.render(cfg, rgb_out=rgb, ids_out=ids, part=(0, 2))
        assert (rgb == -7.0).all(-1).sum() > 0.4 * H * W  # the other half is untouched
        t0 = time.time()
        sc.render(cfg, rgb_out=rgb, ids_out=ids, part=(1, 2))
        dt = time.time() - t0
        assert rgb.tobytes() == full.tobytes() and np.array_equal(ids, full_ids)
        assert dt < 1.0, dt


def test_compact_row_layout_of_the_distributed_ranks(cge):
    """What a rank of cge_render_distributed does, on one GPU: the tile-row partition (CGE_FLAG_PARTITION_TILE_ROWS) of every
    part reassembles to the full frame bit for bit, ragged sizes included (cge_render scatters a part's packed tiles on the
    host; the ranks' compact device layout is exercised by tools/multi_gpu_check.py and bench.py under torchrun)."""
    for name, size in (("c3_teapot_soft", (203, 117)), ("c4_monkey_mirror", (256, 144)), ("c1_cornell", (64, 61))):
        cfg = cge.configs.get(name, *size)
        H, W = cfg["height"], cfg["width"]
        with cge.Scene(cge.load_scene(cfg)) as sc:
            full, full_ids, _ = sc.render(cfg)
            for parts in (2, 3, 8):
                rgb = np.full((H, W, 3), -7.0, np.float32)
                ids = np.full((H, W), -9, np.int32)
                for k in range(parts):
                    sc.render(cfg, rgb_out=rgb, ids_out=ids, part=(k, parts), flags=cge.FLAG_PARTITION_TILE_ROWS)
                assert rgb.tobytes() == full.tobytes() and np.array_equal(ids, full_ids), (name, parts)
