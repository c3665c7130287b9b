"""One-process-per-GPU plumbing for a sharded frame.

The image is partitioned into interleaved bands of `band_rows` rows (band b belongs to rank b % world); scene, tree
and camera are replicated; every rank owns the film entries of its bands. Rendering needs no exchange. What has to
cross NVLink is the packed LDR frame, once per displayed frame, to rank 0:

  mode "peer" (fused, default): rank 0 owns two frame buffers and a small flag array (plain cudaMalloc, exported over
      CUDA IPC); every rank maps them, and the trace kernel's epilogue (or the sample-plane accumulation) stores each
      packed pixel straight into rank 0's buffer (P2P store over NVLink, raytracer_rs_b200/csrc/kernels.cu `ldr_remote`).
      The fence is device side and stream ordered, without NCCL or the host: after its kernels of frame k every rank
      stores k+1 into its slot flags[rank]; rank 0's stream waits until all slots reached k+1 before it touches the
      frame, then publishes flags[world] = k+1 ("frame k has been read"). Frames alternate between the two buffers and a
      rank starts storing frame k only once frame k-2 (the previous user of that buffer) has been read, so stores never
      race with rank 0's readback however far the other ranks run ahead; that wait sits in the same launch as the
      rank's "frame k-1 is done" signal (rt_stream_signal_then_wait), so a frame costs every rank one fence launch. Waits are bounded (a lost peer becomes an
      error count, rt_sync_timeouts, not a hung GPU).
  mode "peer_allreduce": same stores, but one small NCCL all-reduce of the ray counters per frame is both the global
      ray count and the completion fence (the previous design; kept for comparison).
  mode "nccl" (baseline): every rank compacts its rows (rt_get_owned_ldr_rows_device) and rank 0 gathers them with
      torch.distributed.gather, then scatters the rows into place.

torch is plumbing here (process group, streams, the receive buffer); no render arithmetic happens in Python.
`band_partition` is pure host logic and is what the CPU (gloo) tests exercise.
"""
from __future__ import annotations

import numpy as np


def band_partition(height: int, world: int, band_rows: int = 8):
    """rows owned by every rank: band b (rows b*band_rows ...) belongs to rank b % world. Mirrors
    rt_raytracer::owns_row (raytracer_rs_b200/csrc/raytracer.cu)."""
    rows = np.arange(height, dtype=np.int64)
    owner = (rows // band_rows) % world
    return [rows[owner == r] for r in range(world)]


def assemble_frame(parts, partition, width: int, height: int) -> np.ndarray:
    """Inverse of the sharding: parts[r] holds rank r's owned rows compacted in ascending row order."""
    frame = np.empty((height, width), dtype=np.uint32)
    for part, rows in zip(parts, partition):
        frame[rows] = np.asarray(part, dtype=np.uint32).reshape(-1, width)[: len(rows)]
    return frame.reshape(-1)


class _DevPtr:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class HostFrameGather:
    """Delivers full frames to HOST memory with every rank using its own PCIe link: the frame lives in POSIX shared memory that
    every process page-locks (rt_host_register); a rank stores its pixels into one of two device staging frames (rt_set_ldr_target,
    alternating), copies the bands it owns from there into the shared frame on a copy stream (one strided 2-D copy,
    rt_copy_owned_rows) while its next frame traces, and publishes "my rows of frame k have arrived" in a shared flag array in stream
    order. Rank 0's host waits for all flags. No NVLink, no NCCL, no funnel through one link: at 8 GPUs each link carries 1 MB of an
    8.3 MB frame. Call order per frame: begin_frame() -> tracer.trace_rows(...) -> publish(); rank 0: wait_frame(keep) -> frame(k).

    Reuse protocol: frame k uses staging / host buffer k & 1. A rank waits (on the device) for its own copy of frame k-2 before
    frame k stores into the same staging buffer, and (on the host) for `consumed >= k - 1` — rank 0 publishes how many frames its
    caller is done with — before it overwrites the host buffer of frame k-2."""

    def __init__(self, tracer, rank: int, world: int, device, stream, name: str):
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.torch, self.tracer, self.rank, self.world, self.device, self.stream = torch, tracer, rank, world, device, stream
        self.W, self.H = tracer.width, tracer.height
        self.nbytes = self.W * self.H * 4
        total = 2 * self.nbytes + 4096  # two frames + one page of flags (flags[0..world): arrived, flags[world]: consumed)
        if rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(name=name, create=True, size=total)  # a fresh segment is zero-filled: all flags start at 0
            except FileExistsError:  # left behind by a run that died: take it over
                stale = shared_memory.SharedMemory(name=name, create=False)
                stale.close()
                stale.unlink()
                self.shm = shared_memory.SharedMemory(name=name, create=True, size=total)
        dist.barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name, create=False)
            try:  # only the creator may unlink the segment: keep this process's resource tracker out of it (Python < 3.13 has no track=False)
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        # views by ADDRESS, not through the buffer protocol: a view a caller still holds must not keep the segment from being closed
        import ctypes

        anchor = ctypes.c_char.from_buffer(self.shm.buf)
        self.host_base = ctypes.addressof(anchor)
        del anchor
        self.frames = [np.ctypeslib.as_array((ctypes.c_uint32 * (self.W * self.H)).from_address(self.host_base + k * self.nbytes)) for k in (0, 1)]
        self.flags = np.ctypeslib.as_array((ctypes.c_uint32 * 1024).from_address(self.host_base + 2 * self.nbytes))
        self.dev_base = tracer.host_register(self.host_base, total)  # device address of the same range
        self.staging = [tracer.device_alloc(self.nbytes), tracer.device_alloc(self.nbytes)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copy_done = [None, None]
        self.frame_no, self.taken, self.kernels = 0, 0, 0
        torch.cuda.synchronize(device)
        dist.barrier()

    def begin_frame(self):
        k = self.frame_no
        if self.copy_done[k & 1] is not None:  # the copy of frame k-2 has left this staging buffer
            self.stream.wait_event(self.copy_done[k & 1])
        self.tracer.set_ldr_target(self.staging[k & 1])

    def publish(self):
        """enqueue: copy of this rank's rows of the frame just traced into the shared host frame, then the arrival flag"""
        torch, k = self.torch, self.frame_no
        self._mark_consumed()
        if k >= 2:  # the consumer may still hold the host buffer of frame k-2
            self._spin(lambda: int(self.flags[self.world]) >= k - 1, "rank 0 to release the host buffer of frame %d" % (k - 2))
        traced = torch.cuda.Event()
        traced.record(self.stream)
        self.copy_stream.wait_event(traced)
        cs = self.copy_stream.cuda_stream
        self.tracer.copy_owned_rows(self.staging[k & 1], self.host_base + (k & 1) * self.nbytes, cs)
        self.tracer.signal_flag_on_stream(self.dev_base + 2 * self.nbytes + 4 * self.rank, k + 1, cs)
        self.kernels += 1
        done = torch.cuda.Event()
        done.record(self.copy_stream)
        self.copy_done[k & 1] = done
        self.frame_no += 1

    def _spin(self, done, what: str, timeout_s: float = 30.0):
        """busy-waits on a condition over the shared flags; a peer that died turns into an error instead of a hang"""
        import time

        n, t0 = 0, None
        while not done():
            n += 1
            if n & 0x3FF == 0:
                t0 = t0 or time.perf_counter()
                if time.perf_counter() - t0 > timeout_s:
                    raise RuntimeError("HostFrameGather (rank %d): timed out after %.0f s waiting for %s" % (self.rank, timeout_s, what))

    def _mark_consumed(self):
        # rank 0: the frames handed out by the previous wait_frame are done with (the caller came back for more)
        if self.rank == 0:
            self.flags[self.world] = self.taken

    def wait_frame(self, keep: int = 0):
        """rank 0: blocks until all but the last `keep` published frames have arrived from every rank. The frames this call hands out
        (frame(k) for k < frame_no - keep) stay valid until rank 0's next publish() or wait_frame()."""
        self._mark_consumed()
        target = self.frame_no - keep
        if target <= self.taken:
            return
        self._spin(lambda: int(self.flags[: self.world].min()) >= target, "the rows of frame %d from every rank" % (target - 1))
        self.taken = target

    def frame(self, k: int) -> np.ndarray:
        """host view of frame k (valid between the wait_frame that covers k and rank 0's next publish() / wait_frame(); dangling after
        close(): copy what must outlive the gather)"""
        return self.frames[k & 1]

    def close(self):
        torch = self.torch
        self.tracer.set_ldr_target(None)
        torch.cuda.synchronize(self.device)
        self.copy_stream.synchronize()
        self.tracer.host_unregister(self.host_base)
        for p in self.staging:
            self.tracer.device_free(p)
        self.frames, self.flags = None, None
        import torch.distributed as dist

        dist.barrier()
        self.shm.close()
        if self.rank == 0:
            self.shm.unlink()


class FrameGather:
    """Delivers full frames to rank 0. Call order per frame: begin_frame() -> tracer.trace_rows(0, H, spp) ->
    device_gather() -> (rank 0) read_frame_into(pinned_host_tensor), or the pipelined read_frame_async(pinned) ...
    wait_frame(), or device_gather(release=True) when the frame stays on the device."""

    def __init__(self, tracer, rank: int, world: int, device, stream, mode: str = "peer", band_rows: int = 8, fused_signal: bool = False, fence: str = "kernel"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.tracer, self.rank, self.world, self.device, self.stream, self.mode = tracer, rank, world, device, stream, mode
        self.W, self.H = tracer.width, tracer.height
        self.partition = band_partition(self.H, world, band_rows)
        self.kernels = 0  # kernels of OURS launched by this object (row compaction, flag fence)
        # pipelined readback: copies enqueued on the copy stream and not yet waited for (their completion events, oldest first)
        self.copy_stream, self.copy_events, self.copy_pending, self.copy_seq, self.flags_view = None, None, [], 0, None
        self.frame_no = 0
        nbytes = self.W * self.H * 4
        # mode "peer", fused_signal: the trace kernel's last warp out publishes "my stores of frame k are done" itself
        # (rt_set_done_signal) instead of a signal launch behind the kernel. Measured at 2 GPUs: 0.1661 ms per step against 0.1609 with
        # the separate launch — every warp has to fence its peer stores at system scope before it checks out, which costs more than
        # the launch it saves — so the separate launch is the default.
        self.fused_signal = fused_signal and mode == "peer"
        # mode "peer", fence "memops": the per-frame fence is made of stream memory operations (rt_stream_write_value /
        # rt_stream_wait_value: the stream's front end stores / polls the flag, no kernel launch) instead of the flag kernels; no timeout
        # (measured, 100 steps: 2 GPUs 0.1601 ms per step against 0.1629 with the flag kernels — what the memory operations save is
        # rank 0's wait launch —, 4 GPUs 0.1575 vs 0.1570, 8 GPUs 0.1603 vs 0.1544: seven wait operations in a row cost more than one
        # kernel polling eight flags. The flag kernels stay the default; their waits are also bounded.)
        self.fence = fence if mode == "peer" else "kernel"
        assert self.fence in ("kernel", "memops")
        self.buffer_copy_event = [None, None]  # rank 0: completion event of the last host copy that read each frame buffer
        if mode in ("peer", "peer_allreduce"):
            handles = [None, None, None]
            self.local_bufs = []
            flag_bytes = 4 * (world + 1)
            if rank == 0:
                self.local_bufs = [tracer.device_alloc(nbytes), tracer.device_alloc(nbytes), tracer.device_alloc(256)]
                torch.as_tensor(_DevPtr(self.local_bufs[2], 256), device=device).zero_()  # flags start at frame 0
                torch.cuda.synchronize(device)
                handles = [tracer.ipc_export(p) for p in self.local_bufs]
            box = [handles]
            dist.broadcast_object_list(box, src=0)
            handles = box[0]
            if rank == 0:
                mapped = list(self.local_bufs)
            else:
                mapped = [tracer.ipc_open(h) for h in handles]
            self.mapped = mapped
            self.targets, self.flags = mapped[:2], mapped[2]
            assert flag_bytes <= 256
            self.frames = None
            if rank == 0:
                self.frames = [torch.as_tensor(_DevPtr(p, nbytes), device=device).view(torch.int32) for p in self.local_bufs[:2]]
            tracer.set_ldr_target(self.targets[0])
            self.consumed_signalled = 0
        elif mode == "nccl":
            self.max_rows = max(len(r) for r in self.partition)
            self.compact = torch.zeros(self.max_rows * self.W, dtype=torch.int32, device=device)
            self.recv = [torch.zeros_like(self.compact) for _ in range(world)] if rank == 0 else None
            self.frame = torch.zeros(self.H, self.W, dtype=torch.int32, device=device) if rank == 0 else None
            self.row_index = [torch.as_tensor(r, device=device) for r in self.partition] if rank == 0 else None
        else:
            raise ValueError("gather mode must be 'peer' or 'nccl'")
        torch.cuda.synchronize(device)
        dist.barrier()

    def _counters(self):
        """zero copy view of the library's ray counters of the LAST trace call (two sets alternate, so ask per call)"""
        return self.torch.as_tensor(_DevPtr(self.tracer.counters_device_ptr(), 32), device=self.device).view(self.torch.int64)

    def begin_frame(self):
        """Call before tracing a frame. In mode "peer" the wait for the buffer the frame will be stored into ("frame k-2 has
        been read") sits in the fence launch at the end of frame k-1 (device_gather); with fused_signal the trace call is
        armed to publish this rank's "frame k is done" flag from its last warp out."""
        if self.fused_signal:
            self.tracer.set_done_signal(self.flags + 4 * self.rank, self.frame_no + 1)

    def device_gather(self, release: bool = False):
        """Enqueue (on the tracer's stream) whatever makes the frame just traced complete on rank 0. release=True
        (rank 0, mode "peer"): the frame is not going to be read, its buffer may be reused at once."""
        torch, dist = self.torch, self.dist
        with torch.cuda.stream(self.stream):
            if self.mode == "peer":
                k = self.frame_no
                fused = self.fused_signal  # flags[rank] = k + 1 was published by the trace kernel itself
                if self.fence == "memops":
                    t = self.tracer
                    if self.rank == 0:
                        for r in range(1, self.world):  # (rank 0's own stores are ordered by its stream)
                            t.stream_wait_value(self.flags + 4 * r, k + 1)
                        if release:
                            t.stream_write_value(self.flags + 4 * self.world, k + 1)
                            self.consumed_signalled = k + 1
                    else:
                        if not fused:
                            t.stream_write_value(self.flags + 4 * self.rank, k + 1)
                        if k >= 1:
                            t.stream_wait_value(self.flags + 4 * self.world, k)
                elif self.rank == 0:
                    # one launch: (my stores of frame k are done ->) wait for everybody's -> (optionally) frame k is read
                    self.tracer.wait_flags(self.flags, self.world, k + 1, -1 if fused else 0, self.world if release else -1)
                    self.kernels += 1
                    if release:
                        self.consumed_signalled = k + 1
                elif k >= 1:
                    # one launch: (my stores of frame k are done ->) frame k+1 (same buffer as frame k-1) may be stored once
                    # rank 0 has read frame k-1, which it publishes as flags[world] = k
                    if fused:
                        self.tracer.wait_flags(self.flags + 4 * self.world, 1, k)
                    else:
                        self.tracer.signal_then_wait(self.flags + 4 * self.rank, k + 1, self.flags + 4 * self.world, k)
                    self.kernels += 1
                elif not fused:
                    self.tracer.signal_flag(self.flags + 4 * self.rank, k + 1)  # my stores of frame 0 are done
                    self.kernels += 1
                self.ready = k & 1
                self.frame_no += 1
                nxt = self.frame_no & 1
                self.tracer.set_ldr_target(self.targets[nxt])
                if self.rank == 0 and self.buffer_copy_event[nxt] is not None:
                    # rank 0's own trace kernel is not held by the flag protocol: before it stores the next frame into this
                    # buffer, the pipelined host copy that last read the buffer must have finished
                    self.stream.wait_event(self.buffer_copy_event[nxt])
                    self.buffer_copy_event[nxt] = None
            elif self.mode == "peer_allreduce":
                # global ray counters + completion fence: once this all-reduce has finished on rank 0, every rank's
                # trace kernel (earlier on its stream) has completed, so its peer stores are visible
                dist.all_reduce(self._counters())
                self.ready = self.frame_no & 1
                self.frame_no += 1
                self.tracer.set_ldr_target(self.targets[self.frame_no & 1])
            else:
                self.tracer.owned_ldr_rows_to(self.compact.data_ptr())
                self.kernels += 1
                dist.gather(self.compact, self.recv, dst=0)
                if self.rank == 0:
                    for r in range(self.world):
                        n = len(self.partition[r])
                        self.frame.index_copy_(0, self.row_index[r], self.recv[r][: n * self.W].view(n, self.W))

    def rearm(self):
        """after something else (HostFrameGather) used the tracer's LDR target: point it at this gather's next frame buffer again"""
        if self.mode in ("peer", "peer_allreduce"):
            self.tracer.set_ldr_target(self.targets[self.frame_no & 1])

    def release_frame(self):
        """rank 0, mode "peer": the completed frame is no longer needed on the device (it was copied out, or nobody wants
        it): lets the other ranks overwrite its buffer. Stream ordered; idempotent per frame."""
        if self.mode == "peer" and self.rank == 0 and self.consumed_signalled < self.frame_no:
            self.tracer.signal_flag(self.flags + 4 * self.world, self.frame_no)
            self.consumed_signalled = self.frame_no
            self.kernels += 1

    def read_frame_async(self, host_tensor):
        """rank 0: pipelined readback. The device -> (pinned) host copy of the completed frame runs on a copy stream
        while the next frame is traced; in mode "peer" the "frame has been read" flag is published on that stream when
        the copy is done, so the other ranks can never overwrite a buffer that is still being read. The pixels are valid
        after wait_frame(). Enqueue the copy of frame k BEFORE waiting for the copy of frame k-1 (wait_frame(keep=1)
        with two alternating host buffers): the PCIe link then never idles while the host gets around to the next call."""
        torch = self.torch
        if self.mode != "peer":  # only the flag protocol protects a buffer that is still being copied
            return self.read_frame_into(host_tensor)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.copy_events = [torch.cuda.Event(), torch.cuda.Event()]
            self.flags_view = torch.as_tensor(_DevPtr(self.local_bufs[2], 256), device=self.device).view(torch.int32)
        ready = torch.cuda.Event()
        ready.record(self.stream)  # after the fence of the frame just gathered
        self.copy_stream.wait_event(ready)
        done = self.copy_events[self.copy_seq & 1]
        self.copy_seq += 1
        with torch.cuda.stream(self.copy_stream):
            host_tensor.copy_(self.frames[self.ready], non_blocking=True)
            if self.consumed_signalled < self.frame_no:
                self.flags_view[self.world:self.world + 1].fill_(self.frame_no)  # frame k has been read (k + 1)
                self.consumed_signalled = self.frame_no
            done.record(self.copy_stream)
        self.copy_pending.append(done)
        self.buffer_copy_event[self.ready] = done

    def wait_frame(self, keep: int = 0):
        """Blocks until at most `keep` of the copies enqueued by read_frame_async are still in flight (oldest first)."""
        while len(self.copy_pending) > keep:
            self.copy_pending.pop(0).synchronize()

    def read_frame_into(self, host_tensor):
        """rank 0: device -> (pinned) host copy of the completed frame, synchronous."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            src = self.frames[self.ready] if self.mode in ("peer", "peer_allreduce") else self.frame.view(-1)
            host_tensor.copy_(src, non_blocking=True)
            self.release_frame()
        self.stream.synchronize()

    def global_counters(self):
        """[shadow rays, primary hits, bounce rays, blocked] of the last frame summed over ranks (collective call)."""
        with self.torch.cuda.stream(self.stream):
            c = self._counters().clone()
            if self.mode != "peer_allreduce":
                self.dist.all_reduce(c)
        self.stream.synchronize()
        return [int(x) for x in c.cpu()]

    def close(self):
        if self.mode in ("peer", "peer_allreduce"):
            self.tracer.set_ldr_target(None)
            self.torch.cuda.synchronize(self.device)
            self.dist.barrier()
            if self.rank == 0:
                for p in self.local_bufs:
                    self.tracer.device_free(p)
            else:
                for p in self.mapped:
                    self.tracer.ipc_close(p)
