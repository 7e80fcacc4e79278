"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path.

The frame is partitioned into units of 8x4 tiles dealt round robin (cge_render's part_index / part_count): single tiles
(rank r renders tiles r, r+N, ...) or, as cge_render_distributed does, whole tile rows (rank r renders tile rows r, r+N, ...).
Each rank derives its tile list independently; here two gloo ranks exchange them and check that the lists are disjoint, cover
every tile exactly once, are balanced to within one unit, and that bench.py's reference arm obeys the 'rank 0 prints, the
others exit 0 without work' rule under torch.distributed.run."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import importlib
pkg = importlib.import_module("computer-graphics-engine_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for (w, h) in [(3840, 2160), (1024, 1024), (33, 17), (8, 4), (7, 3)]:
    for rows in (False, True):
        mine = pkg.partition_tiles(w, h, rank, world, tile_rows=rows)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        allt = sorted(t for g in gathered for t in g)
        tiles_x = (w + 7) // 8
        n_tiles = tiles_x * ((h + 3) // 4)
        unit = tiles_x if rows else 1
        ok &= allt == list(range(n_tiles))
        ok &= max(len(g) for g in gathered) - min(len(g) for g in gathered) <= unit
        ok &= all((t // unit) % world == r for r, g in enumerate(gathered) for t in g)
    # dynamic dealing (CGE_FLAG_DYNAMIC_TILES): every rank derives the same plan; its static rows plus the pool's chunks cover every
    # tile exactly once, chunk c lies in rank c % world's rows, and whoever takes the chunks - here: rank (c * 7) % world - the union holds
    for pct, chunks in ((25, 2), (10, 1), (90, 8), (1, 3)):
        static, pool = pkg.dynamic_tile_plan(w, h, world, pct, chunks)
        taken = [t for c, ch in enumerate(pool) if (c * 7) % world == rank for t in ch]
        gathered = [None] * world
        dist.all_gather_object(gathered, static[rank] + taken)
        tiles_x = (w + 7) // 8
        ok &= sorted(t for g in gathered for t in g) == list(range(tiles_x * ((h + 3) // 4)))
        ok &= all((t // tiles_x) % world == c % world for c, ch in enumerate(pool) for t in ch)
        ok &= len(pool) == world * chunks
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), "world": world}))
dist.destroy_process_group()
'''


def run_torchrun(args, port):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args,
                          capture_output=True, text=True, env=env, timeout=600)


def test_tile_partition_two_ranks(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = run_torchrun([str(w), str(ROOT)], 29631)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line) == {"ok": True, "world": 2}


def test_reference_arm_prints_once_under_torchrun():
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    r = run_torchrun([str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                      "--config", "c1_cornell", "--cpu-sample-rows", "4"], 29632)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
