"""Two-rank NCCL run of the sharded render + gather to rank 0 (needs >= 2 GPUs; skipped on a single-GPU box, where
tests/test_gpu_parity.py::test_sharded_handles_reassemble_the_frame and tests/test_sharding_gloo.py cover the logic)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, result_path, fused=False, fence="kernel"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import raytracer_rs_b200 as rt
    from raytracer_rs_b200.multi_gpu import FrameGather

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    w, h = 640, 360
    scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
    cfg = dict(recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH)
    t = rt.RayTracer.from_scene(scene, rt.Config(w, h, device=rank, shard_index=rank, shard_count=world, band_rows=8, **cfg))
    stream = torch.cuda.Stream(device=dev)
    t.set_stream(stream.cuda_stream)
    g = FrameGather(t, rank, world, dev, stream, mode=mode, fused_signal=fused, fence=fence)
    host = torch.empty(w * h, dtype=torch.int32).pin_memory()
    ok = True
    for frame in range(5):  # alternates between the two peer buffers
        g.begin_frame()
        t.trace_rows(0, h, 1, want_shadow=False)
        g.device_gather()
        if rank == 0:
            if frame % 2 == 0:
                g.read_frame_into(host)
            else:  # pipelined form
                g.read_frame_async(host)
                g.wait_frame()
            if frame == 0:
                full = rt.RayTracer.from_scene(scene, rt.Config(w, h, device=0, **cfg))
                _, n_shadow = full.trace_rows(0, h, 1)
                ref = full.get_tonemapped_pixels()
                full.close()
            ok = ok and bool(np.array_equal(host.numpy().view(np.uint32), ref))
        if mode != "nccl":
            shadow = g.global_counters()[0]  # collective
            if rank == 0:
                ok = ok and shadow == n_shadow
        dist.barrier()
    ok = ok and t.sync_timeouts() == 0
    if rank == 0:
        open(result_path, "w").write(str(ok))
    g.close()
    t.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,fused,fence", [("peer", False, "kernel"), ("peer", True, "kernel"), ("peer", False, "memops"), ("peer_allreduce", False, "kernel"),
                                              ("nccl", False, "kernel")])
def test_two_rank_gather(tmp_path, mode, fused, fence):
    """mode "peer" three times: the frame-done signal as a launch of its own, published by the trace kernel's last warp out
    (rt_set_done_signal), and the whole fence made of stream memory operations (no kernel launch)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    result = tmp_path / "r.txt"
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, mode, str(result), fused, fence), nprocs=2, join=True)
    assert result.read_text() == "True"


def _host_worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import raytracer_rs_b200 as rt
    from raytracer_rs_b200.multi_gpu import HostFrameGather

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    w, h = 640, 364  # 45 full bands of 8 rows and a partial one
    scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
    cfg = dict(recursions=0, jitter_mode=rt.JITTER_HASHED, seed=3, accel=rt.ACCEL_BVH)
    t = rt.RayTracer.from_scene(scene, rt.Config(w, h, device=rank, shard_index=rank, shard_count=world, band_rows=8, **cfg))
    stream = torch.cuda.Stream(device=dev)
    t.set_stream(stream.cuda_stream)
    g = HostFrameGather(t, rank, world, dev, stream, "rtb200_test_%d" % port)
    ok, refs = True, {}
    if rank == 0:
        full = rt.RayTracer.from_scene(scene, rt.Config(w, h, device=0, **cfg))
    for frame in range(6):  # pipelined: the host takes frame k-1 while frame k is on its way
        g.begin_frame()
        t.trace_rows(0, h, 2, want_shadow=False)
        g.publish()
        if rank == 0:
            g.wait_frame(keep=1)
            full.trace_rows(0, h, 2, want_shadow=False)
            refs[frame] = full.get_tonemapped_pixels().copy()
            if frame >= 1:
                ok = ok and bool(np.array_equal(g.frame(frame - 1), refs[frame - 1]))
    if rank == 0:
        g.wait_frame()
        ok = ok and bool(np.array_equal(g.frame(5), refs[5]))
        full.close()
    dist.barrier()
    if rank == 0:
        open(result_path, "w").write(str(ok))
    g.close()
    t.close()
    dist.destroy_process_group()


def test_two_rank_host_frame_gather(tmp_path):
    """Every rank copies the rows it owns over its own PCIe link into one frame in shared page-locked host memory (HostFrameGather): the
    assembled frames equal rank 0's unsharded render, frame after frame, with the copies pipelined behind the next trace."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    result = tmp_path / "r.txt"
    mp.spawn(_host_worker, args=(2, 29700 + os.getpid() % 1000, str(result)), nprocs=2, join=True)
    assert result.read_text() == "True"
