"""torchrun target: N-GPU tile-partitioned render + NCCL gather must equal the 1-GPU frame bit for bit.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py [cfg ...]"""
import importlib, os, sys, json
from pathlib import Path
import numpy as np
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("computer-graphics-engine_b200")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(pkg.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(pkg.Comm.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm = pkg.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, local)
ok = True
# "+bloom" / "+aa" after a config name switch the implemented ExtraFeatures on (bloom runs on rank 0 after the gather)
for spec in (sys.argv[1:] or ["c1_cornell", "c4_monkey_mirror:0.5", "c3_teapot_soft:0.5", "c5_dragon:0.25", "c1_cornell+bloom+aa:0.5",
                              "c4_monkey_mirror+ragged:0.37"]):
    name, scale = (spec.split(":") + ["1.0"])[:2]
    name, *extras = name.split("+")
    full = pkg.configs.get(name)
    cfg = pkg.configs.get(name, int(full["width"] * float(scale)) + (3 if "ragged" in extras else 0),
                          int(full["height"] * float(scale)) + (1 if "ragged" in extras else 0))
    if "bloom" in extras:
        cfg["features"] |= pkg.configs.FEAT_BLOOM_EFFECT
    if "aa" in extras:
        cfg["features"] |= pkg.configs.FEAT_MULTIPLE_RAYS_PER_PIXEL
        cfg["rays_per_pixel_side"] = 2
    H, W = cfg["height"], cfg["width"]
    shared_rgb, shared_ids = comm.host_frame((H, W, 3), np.float32), comm.host_frame((H, W), np.int32)
    with pkg.Scene(pkg.load_scene(cfg), device=local) as sc:
        rgb, ids, st = comm.render(sc, cfg, want_ids=True)
        dist.barrier()
        # the same frame with every rank writing its own rows into the shared host frame (no gather)
        shared_rgb[:] = -7.0
        shared_ids[:] = -7
        dist.barrier()
        _, _, sts = comm.render(sc, cfg, want_ids=True, shared_frame=shared_rgb, shared_ids=shared_ids)
        # ... and with every rank's kernels storing straight into rank 0's device frame over NVLink (no gather either)
        pf = comm.peer_frame(H * W * 12)
        dist.barrier()
        _, _, stp = comm.render(sc, cfg, peer_frame=pf)
        _, _, stp = comm.render(sc, cfg, peer_frame=pf)
        peer_rgb = pkg.device_view(pf, (H, W, 3)).cpu().numpy() if rank == 0 else None
        # ... and with part of the tile rows dealt dynamically between the GPUs (a counter beside rank 0's frame)
        if rank == 0:
            pkg.device_view(pf, (H, W, 3)).fill_(-3.0)
        dist.barrier()
        _, _, std_ = comm.render(sc, cfg, peer_frame=pf, flags=pkg.FLAG_DYNAMIC_TILES)
        _, _, std_ = comm.render(sc, cfg, peer_frame=pf, flags=pkg.FLAG_DYNAMIC_TILES)
        dyn_rgb = pkg.device_view(pf, (H, W, 3)).cpu().numpy() if rank == 0 else None
        rgba = np.zeros((H, W, 4), np.uint8)
        import ctypes as C
        p8 = pkg.params_from_cfg(cfg, pkg.TRAVERSAL_FAST, False, (0, 1), pkg.FLAG_OUTPUT_RGBA8)
        cam8, st8 = pkg.camera_from_cfg(cfg), pkg.CgeStats()
        rc8 = pkg.lib().cge_render_distributed(sc.handle, comm.handle, C.byref(cam8), C.byref(p8), rgba.ctypes.data if rank == 0 else None,
                                               None, C.byref(st8))
        if rank == 0:
            rgb1, ids1, st1 = sc.render(cfg, want_ids=True)
            rgba1, _ = sc.render_rgba8(cfg)
            same = rgb.tobytes() == rgb1.tobytes() and np.array_equal(ids, ids1)
            same_shared = shared_rgb.tobytes() == rgb1.tobytes() and np.array_equal(shared_ids, ids1)
            same8 = rc8 == 0 and np.array_equal(rgba, rgba1)
            same_peer = peer_rgb.tobytes() == rgb1.tobytes()
            same_dyn = dyn_rgb.tobytes() == rgb1.tobytes()
            ok &= same and same_shared and same8 and same_peer and same_dyn
            print(json.dumps({"cfg": spec, "w": cfg["width"], "h": cfg["height"], "ranks": world, "bit_identical_to_1gpu": bool(same),
                              "shared_host_frame_identical": bool(same_shared), "rgba8_identical": bool(same8), "peer_frame_identical": bool(same_peer),
                              "peer_total_ms": round(stp["total_ms"], 3),
                              "dynamic_tiles_identical": bool(same_dyn), "dynamic_total_ms": round(std_["total_ms"], 3),
                              "dynamic_launches_rank0": std_["kernel_launches"],
                              "shared_total_ms": round(sts["total_ms"], 3),
                              "dist_total_ms": round(st["total_ms"], 3), "dist_kernel_ms_rank0": round(st["kernel_ms"], 3),
                              "single_kernel_ms": round(st1["kernel_ms"], 3), "launches_rank0": st["kernel_launches"]}), flush=True)
        dist.barrier()
comm.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
