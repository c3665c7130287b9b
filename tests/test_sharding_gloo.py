"""N > 1 host-side logic on CPU: world_size-2 gloo run of the band sharding + gather + assembly
(raytracer_rs_b200/multi_gpu.py). The pixels come from the CPU oracle here; the GPU path is covered by -m gpu tests
and bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

from raytracer_rs_b200.multi_gpu import assemble_frame, band_partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_partition_covers_every_row_once():
    for h, world, band in [(1080, 8, 8), (1080, 2, 8), (768, 4, 8), (37, 3, 5), (2160, 8, 8)]:
        parts = band_partition(h, world, band)
        allrows = np.sort(np.concatenate(parts))
        assert np.array_equal(allrows, np.arange(h))
        for r, rows in enumerate(parts):
            assert ((rows // band) % world == r).all()
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= band


def test_interleaving_balances_the_empty_bottom_of_the_frame():
    """Q1 leaves the lower ~38 % of a 16:9 frame empty; contiguous bands would idle half the GPUs, interleaved bands
    give every rank the same share of rows from the top (hit-heavy) 60 %."""
    parts = band_partition(1080, 8, 8)
    top = [int((p < 648).sum()) for p in parts]
    assert max(top) - min(top) <= 8


def _worker(rank, world, port, w, h, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import raytracer_rs_b200 as rt
    from oracle_lib import JITTER_FIXED, Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = rt.load_scene(os.path.join(ROOT, "data", "ico2.dae"))
    parts = band_partition(h, world, 8)
    mine = parts[rank]
    orc = Oracle(scene, w, h)
    orc.configure(recursions=0, jitter=JITTER_FIXED)
    for b0 in range(0, len(mine), 8):  # one band at a time, exactly the rows this rank owns
        band = mine[b0:b0 + 8]
        orc.trace_rows(int(band[0]), len(band), 1, threads=1)
    ldr = orc.get_tonemapped_pixels().reshape(h, w)
    max_rows = max(len(p) for p in parts)
    compact = torch.zeros(max_rows * w, dtype=torch.int32)
    compact[: len(mine) * w] = torch.from_numpy(ldr[mine].reshape(-1).view(np.int32))
    recv = [torch.zeros_like(compact) for _ in range(world)] if rank == 0 else None
    dist.gather(compact, recv, dst=0)
    counts = torch.tensor([orc.counters()["rays"]["primary"], orc.counters()["rays"]["shadow"]], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        frame = assemble_frame([r.numpy().view(np.uint32) for r in recv], parts, w, h)
        full = Oracle(scene, w, h)
        full.configure(recursions=0, jitter=JITTER_FIXED)
        full.trace_rows(0, h, 1, threads=1)
        ok = bool(np.array_equal(frame, full.get_tonemapped_pixels()))
        c = full.counters()
        ok_counts = counts.tolist() == [c["rays"]["primary"], c["rays"]["shadow"]]
        with open(result_path, "w") as f:
            f.write(f"{ok} {ok_counts}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_gathers_to_rank0(tmp_path, world):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000) + world
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, port, 96, 70, str(result)), nprocs=world, join=True)
    assert result.read_text() == "True True"


def test_pipelined_readback_keeps_one_copy_in_flight():
    """FrameGather.read_frame_async / wait_frame(keep) with stand-ins for the CUDA objects (no GPU here): the copy of
    frame k is enqueued before the host waits for the copy of frame k-1, events alternate so that a pending copy's
    event is never re-recorded, and every frame publishes its "has been read" flag on the copy stream."""
    import contextlib
    import types

    from raytracer_rs_b200.multi_gpu import FrameGather

    log = []

    class Event:
        n = 0

        def __init__(self):
            Event.n += 1
            self.id, self.recorded, self.synced = Event.n, 0, 0

        def record(self, stream):
            self.recorded += 1
            log.append(("record", self.id, stream.name))

        def synchronize(self):
            self.synced += 1
            log.append(("sync", self.id))

    class Stream:
        def __init__(self, device=None, name="copy"):
            self.name = name

        def wait_event(self, ev):
            log.append(("wait_event", self.name, ev.id))

    class Flags:
        def __getitem__(self, sl):
            return self

        def fill_(self, v):
            log.append(("read_flag", v))

    class Host:
        def copy_(self, src, non_blocking=False):
            log.append(("copy", src))

    cuda = types.SimpleNamespace(Stream=Stream, Event=Event, stream=lambda s: contextlib.nullcontext())
    g = FrameGather.__new__(FrameGather)
    g.torch = types.SimpleNamespace(cuda=cuda, as_tensor=lambda *a, **k: types.SimpleNamespace(view=lambda dt: Flags()), int32=None)
    g.mode, g.device, g.stream, g.world = "peer", None, Stream(name="render"), 4
    g.copy_stream, g.copy_events, g.copy_pending, g.copy_seq, g.flags_view = None, None, [], 0, None
    g.local_bufs, g.frames, g.consumed_signalled = [0, 0, 0], ["buffer0", "buffer1"], 0
    g.buffer_copy_event = [None, None]
    hosts = [Host(), Host()]
    for k in range(5):
        g.ready, g.frame_no = k & 1, k + 1  # what device_gather leaves behind for frame k
        g.read_frame_async(hosts[k & 1])
        assert len(g.copy_pending) == min(k + 1, 2)
        g.wait_frame(keep=1)
        assert len(g.copy_pending) == 1
    g.wait_frame()
    assert g.copy_pending == []
    copies = [e for e in log if e[0] == "copy"]
    assert copies == [("copy", "buffer%d" % (k & 1)) for k in range(5)]
    assert [e[1] for e in log if e[0] == "read_flag"] == [1, 2, 3, 4, 5]
    syncs = [e[1] for e in log if e[0] == "sync"]
    assert len(syncs) == 5 and len(set(syncs)) == 2  # two alternating completion events, each copy waited for once
    # the host never waits for copy k before copy k+1 has been enqueued (except for the last one)
    order = [e for e in log if e[0] in ("copy", "sync")]
    assert [e[0] for e in order] == ["copy", "copy", "sync", "copy", "sync", "copy", "sync", "copy", "sync", "sync"]


def test_peer_fence_protocol_calls():
    """FrameGather.device_gather in mode "peer" with a recording stand-in for the tracer, both forms of the fence.
    Separate signal (fused_signal=False): every rank issues exactly one fence launch per frame — rank r != 0 signals "frame k
    done" as flags[r] = k + 1 and, from frame 1 on, waits in the same launch for flags[world] >= k ("frame k-1 has been read":
    frame k+1 reuses its buffer); rank 0 signals its own slot, waits for all slots and, when the frame stays on the device,
    publishes "read" in that launch too. Fused signal (optional): begin_frame arms the trace call to publish flags[r] = k + 1
    from its last warp out (rt_set_done_signal), the fence launch only waits (rank r != 0: nothing at all for frame 0).
    Fence "memops": the same protocol as stream memory operations (write_value / wait_value), no kernel launch.
    The store target alternates between the two frame buffers; rank 0's render stream waits for the host copy that last
    read a buffer before its own kernel stores into it again."""
    import contextlib
    import types

    from raytracer_rs_b200.multi_gpu import FrameGather

    def gather(rank, world, fused, fence="kernel"):
        calls = []
        tracer = types.SimpleNamespace(
            stream_write_value=lambda p, v: calls.append(("write_value", p, v)),
            stream_wait_value=lambda p, v: calls.append(("wait_value", p, v)),
            signal_flag=lambda p, v: calls.append(("signal", p, v)),
            signal_then_wait=lambda p, v, q, t: calls.append(("signal_then_wait", p, v, q, t)),
            wait_flags=lambda p, n, t, s=-1, r=-1: calls.append(("wait", p, n, t, s, r)),
            set_done_signal=lambda p, v: calls.append(("arm", p, v)),
            set_ldr_target=lambda p: calls.append(("target", p)))
        g = FrameGather.__new__(FrameGather)
        g.torch = types.SimpleNamespace(cuda=types.SimpleNamespace(stream=lambda s: contextlib.nullcontext()))
        g.dist, g.tracer, g.rank, g.world, g.mode = None, tracer, rank, world, "peer"
        g.stream = types.SimpleNamespace(wait_event=lambda ev: calls.append(("render_waits_for_copy", ev)))
        g.flags, g.targets, g.frame_no, g.kernels, g.consumed_signalled = 1000, [0xA000, 0xB000], 0, 0, 0
        g.fused_signal, g.buffer_copy_event, g.fence = fused, [None, None], fence
        return g, calls

    g, calls = gather(rank=2, world=4, fused=False)
    for _ in range(4):
        g.begin_frame()  # enqueues nothing in this form
        g.device_gather()
    assert calls == [("signal", 1008, 1), ("target", 0xB000),
                     ("signal_then_wait", 1008, 2, 1016, 1), ("target", 0xA000),
                     ("signal_then_wait", 1008, 3, 1016, 2), ("target", 0xB000),
                     ("signal_then_wait", 1008, 4, 1016, 3), ("target", 0xA000)]
    assert g.kernels == 4 and g.frame_no == 4

    g, calls = gather(rank=0, world=4, fused=False)
    g.device_gather(release=True)
    g.device_gather()
    assert calls == [("wait", 1000, 4, 1, 0, 4), ("target", 0xB000), ("wait", 1000, 4, 2, 0, -1), ("target", 0xA000)]
    assert g.consumed_signalled == 1 and g.ready == 1 and g.kernels == 2

    g, calls = gather(rank=2, world=4, fused=True)
    for _ in range(3):
        g.begin_frame()
        g.device_gather()
    assert calls == [("arm", 1008, 1), ("target", 0xB000),
                     ("arm", 1008, 2), ("wait", 1016, 1, 1, -1, -1), ("target", 0xA000),
                     ("arm", 1008, 3), ("wait", 1016, 1, 2, -1, -1), ("target", 0xB000)]
    assert g.kernels == 2 and g.frame_no == 3  # the first frame needs no fence launch at all on this rank

    g, calls = gather(rank=0, world=4, fused=True)
    g.begin_frame()
    g.device_gather(release=True)
    g.buffer_copy_event[0] = "copy-of-buffer-0"  # what read_frame_async leaves behind
    g.begin_frame()
    g.device_gather()
    assert calls == [("arm", 1000, 1), ("wait", 1000, 4, 1, -1, 4), ("target", 0xB000),
                     ("arm", 1000, 2), ("wait", 1000, 4, 2, -1, -1), ("target", 0xA000), ("render_waits_for_copy", "copy-of-buffer-0")]
    assert g.buffer_copy_event == [None, None] and g.kernels == 2

    # the same fence made of stream memory operations: no kernel at all
    g, calls = gather(rank=2, world=4, fused=False, fence="memops")
    for _ in range(3):
        g.begin_frame()
        g.device_gather()
    assert calls == [("write_value", 1008, 1), ("target", 0xB000),
                     ("write_value", 1008, 2), ("wait_value", 1016, 1), ("target", 0xA000),
                     ("write_value", 1008, 3), ("wait_value", 1016, 2), ("target", 0xB000)]
    assert g.kernels == 0
    g, calls = gather(rank=0, world=3, fused=False, fence="memops")
    g.device_gather(release=True)
    g.device_gather()
    assert calls == [("wait_value", 1004, 1), ("wait_value", 1008, 1), ("write_value", 1012, 1), ("target", 0xB000),
                     ("wait_value", 1004, 2), ("wait_value", 1008, 2), ("target", 0xA000)]
    assert g.consumed_signalled == 1 and g.kernels == 0


def _host_gather_worker(rank, world, port, w, h, result_path):
    """HostFrameGather between real processes over real POSIX shared memory; the device side (staging frames, strided copy, flag store) is
    a numpy stand-in that applies the same band arithmetic as rt_copy_owned_rows, and CUDA streams / events are inert objects."""
    sys.path.insert(0, ROOT)
    import ctypes
    import types

    import torch
    import torch.distributed as dist

    from raytracer_rs_b200.multi_gpu import HostFrameGather

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = band_partition(h, world, 8)[rank]
    keep = []

    def view(ptr, n):
        return np.ctypeslib.as_array((ctypes.c_uint32 * n).from_address(ptr))

    class Tracer:
        width, height = w, h
        target = None

        def host_register(self, ptr, nbytes):
            return ptr

        def host_unregister(self, ptr):
            pass

        def device_alloc(self, nbytes):
            keep.append(np.zeros(nbytes // 4, np.uint32))
            return keep[-1].ctypes.data

        def device_free(self, p):
            pass

        def set_ldr_target(self, p):
            self.target = p

        def copy_owned_rows(self, src, dst, stream=0):
            s, d = view(src, w * h).reshape(h, w), view(dst, w * h).reshape(h, w)
            d[mine] = s[mine]

        def signal_flag_on_stream(self, flag, value, stream=0):
            view(flag, 1)[0] = value

    class Stream:
        cuda_stream = 0

        def __init__(self, device=None):
            pass

        def wait_event(self, ev):
            pass

        def synchronize(self):
            pass

    class Event:
        def record(self, stream=None):
            pass

    torch.cuda.Stream, torch.cuda.Event, torch.cuda.synchronize = Stream, Event, lambda d=None: None
    t = Tracer()
    g = HostFrameGather(t, rank, world, None, Stream(), "rtb200_gloo_%d" % port)
    ok = True
    rows = np.arange(h, dtype=np.uint32)[:, None]
    for k in range(7):
        g.begin_frame()
        stage = view(t.target, w * h).reshape(h, w)
        stage[:] = 0xDEAD  # rows this rank does not own must never reach the shared frame
        stage[mine] = (1000 * k + rows + np.arange(w, dtype=np.uint32)[None, :])[mine]
        g.publish()
        if rank == 0:
            g.wait_frame(keep=1)
            if k >= 1:  # frame k-1 is complete and stays valid until rank 0 publishes again
                ok = ok and bool(np.array_equal(g.frame(k - 1).reshape(h, w), 1000 * (k - 1) + rows + np.arange(w, dtype=np.uint32)[None, :]))
    if rank == 0:
        g.wait_frame()
        ok = ok and bool(np.array_equal(g.frame(6).reshape(h, w), 6000 + rows + np.arange(w, dtype=np.uint32)[None, :]))
        ok = ok and g.taken == 7
        with open(result_path, "w") as f:
            f.write(str(ok))
    dist.barrier()
    g.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_host_frame_gather_over_shared_memory(tmp_path, world):
    """The N > 1 end-to-end path of bench.py (every rank delivers the rows it owns into one shared host frame, rank 0's host waits for the
    arrival flags, two alternating buffers with a consumed counter) between real processes; 37 rows: a partial band at the bottom."""
    import torch.multiprocessing as mp

    port = 31000 + (os.getpid() % 2000) + world
    result = tmp_path / "result.txt"
    mp.spawn(_host_gather_worker, args=(world, port, 24, 37, str(result)), nprocs=world, join=True)
    assert result.read_text() == "True"
