"""World-size-2 gloo test (CPU) of the multi-GPU host logic: frames shard over ranks with no
data-path collective; the only collective is the MAX-reduce of the timing scalar."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import bench
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bench.N_TRIS = 2000                                   # small frames: the oracle renders them in ms
    sc = bench.frame_scene(pkg, world, rank)              # rank r renders C5 frame r
    rgba, z, tm, rc = orc.render_scene(sc)
    assert rc == 0
    digest = np.frombuffer(hashlib.sha256(rgba.tobytes()).digest(), dtype=np.uint8).copy()
    gathered = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(digest))
    # the timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([bytes(g.numpy().tobytes()) for g in gathered], float(t.item()), sc.name))
    dist.barrier()
    dist.destroy_process_group()


def test_frames_shard_over_ranks_gloo():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    entry.build_oracle()
    from oracle import oracle as orc
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    digests, tmax, name0 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0                                     # MAX over ranks
    assert name0.startswith("c5_frame0")
    # each rank rendered its own, distinct frame: C5 frame r (seed 0xB3200500 + r)
    want = []
    for r in range(world):
        sc = pkg.scenes.scene_c5(r, n_tris=2000)
        rgba, z, tm, rc = orc.render_scene(sc)
        want.append(hashlib.sha256(rgba.tobytes()).digest())
    assert digests == want
    assert digests[0] != digests[1]


def test_single_gpu_workload_is_config4():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import bench
    assert bench.N_TRIS == 100_000
    cfg = bench.workload_config(1)
    assert "configs[3]" in cfg["workload"] and "100k" in cfg["workload"]
    assert "configs[4]" in bench.workload_config(8)["workload"]
