"""World-size-2 test of the one-process-per-GPU host logic on CPU (gloo): every rank renders
the rows its shard owns into ONE shared host framebuffer (/dev/shm), no data-path collective;
rank 0 checks the assembled image against a single-process render.  The renderer here is the
CPU oracle (this is a test: on the GPU box the same rows come from rtclj_render with shard_*);
what is under test is the shard arithmetic, the disjoint-row gather and the reductions bench.py
performs on its timings and segment counts.  The library's own (C++) download plan for a shard is checked
against the same arithmetic in tests/test_host.py::test_shard_plan_covers_exactly_the_shards_rows; the GPU
side of N > 1 runs in tests/test_gpu_host_paths.py (one process, several devices) and in bench.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, shm_path, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    import raytracing_clj_b200 as R
    from raytracing_clj_b200 import render

    cam = R.camera.main_camera(48)
    soa = R.scenes.to_soa(R.scenes.main_hittables())
    H, W = cam.height, cam.width
    if rank == 0:
        with open(shm_path, "wb") as f:
            f.truncate(H * W * 3 * 8)
    dist.barrier()
    fb = np.memmap(shm_path, dtype=np.float64, mode="r+", shape=(H, W, 3))
    mine = render.shard_rows(H, rank, world, 4)
    segs = 0
    local = np.zeros((H, W, 3), dtype=np.float64)   # this rank's "device image": full-size layout, own rows filled
    for j in mine:
        lin, _, st = O.render(soa, cam, 6, 50, seed=3, flags=O.FLAGS_MAIN, rows=(j, j + 1), want_rgb8=False)
        local[j] = lin[j]
        segs += st.segments
    # the gather: the copies the LIBRARY plans for this shard (rtclj_shard_plan, the C++ behind rtclj_render's
    # download), applied to the shared framebuffer -- each rank writes only its own rows
    src, dst = local.reshape(-1).view(np.uint8), np.asarray(fb).reshape(-1).view(np.uint8)
    for off, pitch, width, height in render.shard_plan(H, W * 24, rank, world, 4, 4096):
        for r in range(height):
            dst[off + r * pitch: off + r * pitch + width] = src[off + r * pitch: off + r * pitch + width]
    fb.flush()
    t = torch.tensor([float(segs)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)           # whole-job segment count
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # slowest rank defines the step time
    dist.barrier()
    if rank == 0:
        whole, _, st = O.render(soa, cam, 6, 50, seed=3, flags=O.FLAGS_MAIN, want_rgb8=False)
        ok = bool(np.array_equal(np.asarray(fb), whole)) and int(t.item()) == st.segments and ms.item() == 10.0 + world - 1
        with open(result_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_assemble_one_image(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    shm = f"/dev/shm/rtclj_test_{os.getpid()}"
    result = str(tmp_path / "result.txt")
    try:
        mp.spawn(_worker, args=(2, port, shm, result), nprocs=2, join=True)
        assert open(result).read() == "ok"
    finally:
        if os.path.exists(shm):
            os.unlink(shm)
