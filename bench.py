#!/usr/bin/env python
"""bench.py — throughput of the per-pixel render loop on B200 (BASELINE.json metric).

A step = one full frame (1 sample per pixel over all H rows) of data/thai2.dae at 1920x1080, primary + shadow
rays (recursions = 0, fixed sub-pixel offset 0.5: the pinned parity mode), accumulated into the device film.
  value    : Mrays/s (primary + shadow), scene/film resident in HBM, timed with CUDA events on the launching stream
  e2e      : same metric through the public API with HOST buffers: camera state in, 8.3 MB LDR frame out (pinned),
             every step, wall clock
  roofline : algorithmic bytes of the reference algorithm (24 B per cube test + 36 B per triangle test + 4 B per
             pixel) / measured duration of the trace kernel, against the measured HBM copy peak
  cpu_baseline / --impl reference : the CPU oracle (C++ restatement of the reference, oracle/) on the host cores

N > 1 (torchrun, one process per GPU): the frame is sharded by interleaved 8-row bands, the scene is replicated,
rank 0 receives the packed frame over NVLink (see --gather), no other exchange.
  --scaling weak (default): a step renders N samples per pixel of the frame (hashed jitter), so every rank traces
      H/N rows x N samples = as many camera rays as the single GPU does in its step; value = all rays of all ranks / time
  --scaling strong: a step is one sample per pixel whatever N is (each rank traces H/N rows); a 0.23 ms frame cut in
      N pieces is bounded by launch + fence latency and by the slowest tile, see DESIGN.md section 7
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+shadow) on thai2.dae 1920x1080"
UNIT = "Mrays/s"
WORKLOADS = {
    # name: (file, width, height, spp)
    "thai2_1080p": ("thai2.dae", 1920, 1080, 1),
    "ico2_1024x768": ("ico2.dae", 1024, 768, 1),
    "4boxes_1080p": ("4boxes.dae", 1920, 1080, 1),
    "ico3_tex_1080p": ("ico3_tex.dae", 1920, 1080, 1),
    "thai2_4k_16spp": ("thai2.dae", 3840, 2160, 16),
}
# Exact work of the REFERENCE algorithm (octree, triangles_per_leaf = 70) for one pinned-mode frame, counted by
# the oracle (tools/count_work.py; checked again against the live oracle in the cpu_baseline leg):
# (primary rays, shadow rays, cube tests, triangle tests)
REFERENCE_WORK = {
    "thai2_1080p": (2073600, 525594, 64718736 + 21902544, 83666141 + 36577192),
    "ico2_1024x768": (786432, 313374, 17397368 + 6640944, 28736151 + 16316966),
    "4boxes_1080p": (2073600, 311140, 0, 99532800 + 14934720),
    "ico3_tex_1080p": (2073600, 619720, 38552600 + 13133280, 56826791 + 32284213),
}


def algorithmic_bytes(workload: str):
    if workload not in REFERENCE_WORK:
        return None
    prim, _sh, cubes, tris = REFERENCE_WORK[workload]
    return 24 * cubes + 36 * tris + 4 * prim


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.max_mhz = [], None  # (time, sm MHz, reason names)
        self._halt = threading.Event()
        self.ready = threading.Event()
        self.err = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
                nv.nvmlClocksThrottleReasonSyncBoost: "sync_boost",
                nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting",
            }
            while not self._halt.is_set():
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), mhz, tuple(name for bit, name in names.items() if r & bit)))
                self.ready.set()
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report it, never fake numbers
            self.err = repr(e)
            self.ready.set()

    def window(self, t0, t1):
        """median SM clock and the union of throttle reasons of the samples taken inside [t0, t1]"""
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        mhz = sorted(x[1] for x in inside)
        reasons = sorted({r for x in inside for r in x[2]})
        return {"sm_mhz": (mhz[len(mhz) // 2] if mhz else None), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), **({"error": self.err} if self.err else {})}

    def stop(self):
        self._halt.set()
        self.join(timeout=2)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the Rust original cannot be
    built here) on all host threads. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import raytracer_rs_b200 as rt  # product loader only flattens the scene file; the render below is the oracle's
    from oracle_lib import JITTER_FIXED, Oracle, lib as orc_lib

    fname, w, h, spp = WORKLOADS[args.workload]
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    orc = Oracle(scene, w, h, rt.DEFAULT_TRIANGLES_PER_LEAF)
    orc.configure(recursions=0, jitter=JITTER_FIXED)
    threads = orc_lib().orc_max_threads()
    # Bounded sample: a step is the whole frame unless steps + warmup whole frames would take longer than --reference-budget
    # seconds on this host; then a step is one of `parts` equal row ranges, rotating over the frame from step to step, so
    # that any `parts` consecutive steps cover the frame once (rays are counted exactly either way).
    t1 = time.perf_counter()
    orc.trace_rows(0, h, spp, threads=threads)
    orc.get_tonemapped_pixels()
    t_frame = time.perf_counter() - t1
    parts = max(1, min(h // 8, int(-(-t_frame * (args.steps + args.warmup) // args.reference_budget))))
    rows = -(-h // parts)

    def step(i):
        first = (i % parts) * rows
        orc.trace_rows(first, min(rows, h - first), spp, threads=threads)
        orc.get_tonemapped_pixels()

    for i in range(args.warmup):
        step(i)
    orc.counters(reset=True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    c = orc.counters()
    rays = c["rays"]["primary"] + c["rays"]["shadow"]
    value = rays / dt / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "reference scene fixture data/%s (the upstream repository's own scene), pinned camera" % fname,
        "config": {"workload": args.workload, "width": w, "height": h, "spp": spp, "recursions": 0, "jitter": "fixed 0.5",
                   "accel": "reference octree, triangles_per_leaf 70",
                   "step": ("one full frame" if parts == 1 else "%d rows (1/%d of the frame, rotating)" % (rows, parts)) + " + get_tonemapped_pixels"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": ("%d full frames" % args.steps if parts == 1 else
                                    "%d steps of %d rows each (1/%d of the frame, rotating over it)" % (args.steps, rows, parts))
                                   + " (%d rays) of the same workload, OpenMP over rows" % rays},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, rt, scene):
    """Oracle timed on the GPU box's host cores: all threads (value) and one thread (what the reference ships)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import JITTER_FIXED, Oracle, lib as orc_lib

    _f, w, h, spp = WORKLOADS[workload]
    spp = min(spp, 1)  # bounded sample: one sample per pixel
    orc = Oracle(scene, w, h, rt.DEFAULT_TRIANGLES_PER_LEAF)
    orc.configure(recursions=0, jitter=JITTER_FIXED)
    threads = orc_lib().orc_max_threads()
    orc.trace_rows(0, h, 1, threads=threads)  # warm
    orc.counters(reset=True)
    frames = 0
    t0 = time.perf_counter()
    while True:
        orc.trace_rows(0, h, 1, threads=threads)
        orc.get_tonemapped_pixels()
        frames += 1
        if time.perf_counter() - t0 > 4.0 or frames >= 20:
            break
    dt = time.perf_counter() - t0
    c = orc.counters(reset=True)
    rays = c["rays"]["primary"] + c["rays"]["shadow"]
    work_ok = None
    if workload in REFERENCE_WORK:
        p, s, cu, tr = REFERENCE_WORK[workload]
        got = (c["rays"]["primary"] // frames, c["rays"]["shadow"] // frames,
               (c["cube_tests"]["primary"] + c["cube_tests"]["shadow"]) // frames,
               (c["tri_tests"]["primary"] + c["tri_tests"]["shadow"]) // frames)
        work_ok = got == (p, s, cu, tr)
    # single thread on a quarter of the rows (bounded), scaled by its own ray count
    rows = max(1, h // 4)
    t1 = time.perf_counter()
    orc.trace_rows(0, rows, 1, threads=1)
    dt1 = time.perf_counter() - t1
    c1 = orc.counters(reset=True)
    rays1 = c1["rays"]["primary"] + c1["rays"]["shadow"]
    return {
        "value": rays / dt / 1e6,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": "%d full frames (%d rays) all threads; single-thread leg: rows 0..%d (%d rays)" % (frames, rays, rows, rays1),
        "single_thread_value": rays1 / dt1 / 1e6,
        "reference_work_matches_constants": work_ok,
    }


class _DevPtr:
    """Wraps a raw device pointer for torch.as_tensor (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="thai2_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--accel", default="bvh", choices=["bvh", "octree", "cwbvh", "bvh4", "lbvh"])
    ap.add_argument("--gather", default="peer", choices=["peer", "peer_allreduce", "nccl"], help="N>1: how rank 0 receives the frame")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N>1: samples per pixel per step = N (weak) or 1 (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-budget", type=float, default=150.0,
                    help="--impl reference: seconds the whole run may take; longer runs trace a rotating part of the frame per step")
    ap.add_argument("--no-launch-timing", action="store_true",
                    help="developer: RT_TUNE_TIME_LAUNCHES = 0 in the device-timed and end-to-end legs (the library then skips the two CUDA "
                         "events behind launch_stats().trace_kernel_ms); the roofline leg, which needs that figure, switches them back on")
    ap.add_argument("--tune", default="", help="developer: comma separated key=value pairs passed to rt_set_tuning")
    ap.add_argument("--sync-readback", action="store_true",
                    help="e2e leg (N=1): blocking rt_get_tonemapped_pixels after every trace call instead of the pipelined "
                         "rt_get_tonemapped_pixels_async (device->host copy of frame k overlapping the trace of frame k+1)")
    ap.add_argument("--zero-copy", action="store_true",
                    help="e2e leg: let the kernel store packed pixels straight into the pinned host frame (rt_set_host_frame) instead of "
                         "copying the frame after the kernel; measured slower on PCIe (32-byte writes), kept as an option")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import raytracer_rs_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (one process per GPU)" % (args.gpus, args.gpus))
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: raytracer_rs_b200 has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    fname, W, H, spp = WORKLOADS[args.workload]
    base_spp = spp
    if world > 1 and args.scaling == "weak":
        spp = spp * world  # per-GPU work stays what one GPU does at N = 1
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    accel = {"bvh": rt.ACCEL_BVH, "octree": rt.ACCEL_OCTREE, "cwbvh": rt.ACCEL_CWBVH, "bvh4": rt.ACCEL_BVH4, "lbvh": rt.ACCEL_LBVH}[args.accel]
    cfg = rt.Config(W, H, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF if spp == 1 else rt.JITTER_HASHED, accel=accel,
                    device=local_rank, shard_index=rank, shard_count=world, band_rows=8)
    sampler = ClockSampler(local_rank)
    sampler.start()
    tracer = rt.RayTracer.from_scene(scene, cfg)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        tracer.set_tuning(int(k), int(v))
    if args.no_launch_timing:
        tracer.set_tuning(10, 0)
    stream = torch.cuda.Stream(device=dev)
    tracer.set_stream(stream.cuda_stream)

    # ---- multi-GPU plumbing -------------------------------------------------------------------------
    from raytracer_rs_b200 import multi_gpu

    gather = multi_gpu.FrameGather(tracer, rank, world, dev, stream, mode=args.gather) if world > 1 else None

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_frame = torch.empty(W * H, dtype=torch.int32).pin_memory()
    host_frames = [host_frame, torch.empty(W * H, dtype=torch.int32).pin_memory()]
    pipelined = world == 1 and not args.sync_readback and not args.zero_copy

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        if gather is not None:
            gather.begin_frame()
        tracer.trace_rows(0, H, spp, want_shadow=False)
        if gather is not None:
            gather.device_gather(release=True)  # device-timed leg: the frame stays on rank 0

    def global_ray_totals():
        t = tracer.ray_totals()  # exact device counters since the handle was created (synchronises this rank's stream)
        v = torch.tensor([t["primary"], t["shadow"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(v)
        return int(v[0]), int(v[1])

    # ---- device-timed leg ------------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            device_step()
    barrier()
    rays0 = global_ray_totals()
    launches0 = tracer.kernels_launched() + (gather.kernels if gather else 0)
    sampler.ready.wait(timeout=10)
    t_region0 = time.perf_counter()
    events = []
    kernel_ms = []
    with torch.cuda.stream(stream):
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (not inside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            device_step()
            e1.record(stream)
            events.append((e0, e1))
    barrier()
    clocks = sampler.window(t_region0, time.perf_counter())
    launches = tracer.kernels_launched() + (gather.kernels if gather else 0) - launches0
    rays1 = global_ray_totals()
    # rays of the timed steps, counted on the device (with hashed jitter every step draws new samples, so the shadow-ray
    # count differs slightly from step to step)
    n_primary_total, n_shadow_total = (rays1[0] - rays0[0]) / args.steps, (rays1[1] - rays0[1]) / args.steps
    rays_per_step = n_primary_total + n_shadow_total
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- trace-kernel duration for the roofline (library's own CUDA events around the kernel, same stream) ----
    tracer.set_tuning(10, 1)
    with torch.cuda.stream(stream):
        for _ in range(20):
            flush.zero_()
            tracer.trace_rows(0, H, spp, want_shadow=False)
            kernel_ms.append(tracer.launch_stats()["trace_kernel_ms"])
    if args.no_launch_timing:
        tracer.set_tuning(10, 0)
    kms = float(np.mean(kernel_ms))
    kms_t = torch.tensor([kms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kms_t, op=dist.ReduceOp.MAX)
    kms = float(kms_t)

    # ---- end-to-end leg: public API, host buffers, every step ---------------------------------------------------
    def e2e_step(i):
        # per-step input: camera state from the host (travels to the device as the kernel's launch parameters)
        tracer.camera.set_state(0.0, 0.0, (0.0, 0.0, 0.0))
        if gather is not None:
            gather.begin_frame()
        tracer.trace_rows(0, H, spp, want_shadow=False)
        if gather is not None:
            gather.device_gather()
        if pipelined:
            # frame i-1 (copied on the copy stream while frame i traces) is now complete in host memory; then hand
            # frame i to the copy stream
            tracer.wait_pixels()
            tracer.get_tonemapped_pixels_async(host_frames[i & 1].data_ptr())
        elif gather is not None and not args.sync_readback:
            if rank == 0:  # same pipelining on rank 0 of a multi-GPU run: hand frame i to the copy stream, then make sure
                # frame i-1 (the other host buffer) has arrived
                gather.read_frame_async(host_frames[i & 1])
                gather.wait_frame(keep=1)
        elif rank == 0:
            tracer_or_gather_readback()

    def tracer_or_gather_readback():
        if gather is not None:
            gather.read_frame_into(host_frame)
        else:
            tracer.get_tonemapped_pixels_into(host_frame.data_ptr())

    if gather is None and args.zero_copy:
        tracer.set_host_frame(host_frame.data_ptr())  # the kernel stores packed pixels straight into the pinned frame
    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    rays0 = global_ray_totals()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    if pipelined:
        tracer.wait_pixels()  # the last frame is delivered inside the timed region too
    elif gather is not None and rank == 0:
        gather.wait_frame()
    barrier()
    e2e_s = time.perf_counter() - t0
    rays1 = global_ray_totals()
    e2e_rays = (rays1[0] - rays0[0]) + (rays1[1] - rays0[1])
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t)
    e2e_value = e2e_rays / e2e_s / 1e6
    sampler.stop()
    e2e_frame_ok = None
    if rank == 0 and gather is None:
        # the frame delivered by the last e2e step must equal the device frame
        check = np.empty(W * H, np.uint32)
        tracer.set_host_frame(None)
        tracer.get_tonemapped_pixels(check)
        last = host_frames[(args.steps - 1) & 1] if pipelined else host_frame
        e2e_frame_ok = bool(np.array_equal(check, last.numpy().view(np.uint32)))

    # ---- the reference's own call pattern: 50-row bands, full-frame readback after every band (main.rs:200-201) ----
    band = None
    if world == 1 and spp == 1:
        tracer.set_rows_per_call(50)
        calls = (H + 49) // 50
        for _ in range(2 * calls):
            tracer.trace_frame_additive()
            tracer.get_tonemapped_pixels_into(host_frame.data_ptr())
        torch.cuda.synchronize(dev)
        reps = max(1, min(args.steps, 40))
        t0 = time.perf_counter()
        for _ in range(reps * calls):
            tracer.trace_frame_additive()
            tracer.get_tonemapped_pixels_into(host_frame.data_ptr())
        dtb = time.perf_counter() - t0
        band_rays = rays_per_step * (calls * 50 / H) * reps
        band = {"value": band_rays / dtb / 1e6, "unit": UNIT, "calls_per_frame": calls,
                "d2h_bytes_per_call": W * H * 4, "note": "trace_frame_additive (50 rows) + get_tonemapped_pixels per call"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    alg = algorithmic_bytes(args.workload)
    traffic, warp_inst = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            prof = json.load(f)
        traffic = prof.get("%s/%s" % (args.workload, args.accel))
        warp_inst = prof.get("warp_instructions", {}).get("%s/%s" % (args.workload, args.accel))
    except Exception:
        pass
    roofline = None
    if alg is not None:
        note = "traffic-equivalent of the reference algorithm's reads; the <= 2 MB scene is L1/L2 resident, so frac may exceed 1"
        if world == 1 and spp == 1:
            per_launch = float(alg)  # exact oracle counters of the pinned frame
        else:
            # jittered / sharded launches: the pinned frame's average bytes per ray x the rays one rank traces per launch
            p0, s0 = REFERENCE_WORK[args.workload][:2]
            per_launch = alg / (p0 + s0) * rays_per_step / world
            note += "; per-launch bytes estimated from the pinned frame's average bytes per ray"
        achieved = per_launch / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic if world == 1 and spp == 1 else None,
                    "peak_source": peak_src, "kernel": "trace_shade_persistent_kernel<%s>" % args.accel, "kernel_ms": kms,
                    "algorithmic_bytes_per_launch": per_launch, "note": note}
        if roofline["traffic"]:
            # what actually crosses the HBM interface (ncu dram__bytes of one launch, profiles/traffic.json): the film
            roofline["dram_achieved"] = roofline["traffic"] / (kms * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_achieved"] / peak
            roofline["bound_in_practice"] = ("instruction issue at partly filled warps + latency of dependent node loads "
                                             "(ncu: issue slots 65 % busy, 22 of 32 lanes active; profiles/README.md)")
        if warp_inst and world == 1 and spp == 1 and clocks.get("sm_mhz"):
            # the resource that actually binds: warp instructions issued (ncu smsp__inst_executed.sum of one launch,
            # profiles/traffic.json) / measured kernel time, against 4 schedulers x SMs x the SM clock measured during the run
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
            roofline["issue_slots"] = {"warp_instructions_per_launch": warp_inst, "achieved_ginst_s": warp_inst / (kms * 1e-3) / 1e9,
                                       "peak_ginst_s": peak_issue / 1e9, "frac": warp_inst / (kms * 1e-3) / peak_issue,
                                       "note": "one warp instruction per scheduler and cycle; 22 of 32 lanes active on average"}
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps,
        "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "reference scene fixture data/%s (the upstream repository's own scene), pinned camera" % fname,
        "config": {"workload": args.workload, "width": W, "height": H, "spp": spp, "recursions": 0,
                   "jitter": "fixed 0.5" if spp == 1 else "hashed seed 0", "accel": args.accel, "l2_flush_between_steps": True,
                   "step": "one full frame: trace %d rows x %d spp%s" % (H, spp, "" if world == 1 else " (%d spp per GPU-count unit: %s scaling), rows sharded by interleaved 8-row bands, every rank traces its rows x all samples in one launch, packed frame stored into rank 0's buffer over NVLink (%s)" % (base_spp, args.scaling, args.gather)),
                   "primary_rays_per_step": n_primary_total, "shadow_rays_per_step": n_shadow_total},
        "frames_per_s": args.steps / (total_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rt.lib().rt_launch_param_bytes()), "d2h_bytes_per_step": W * H * 4 + 32,
                "ms_per_step": e2e_s / args.steps * 1e3, "readback": ("gather to rank 0 (%s) + %s copy to pinned host memory" % (args.gather, "blocking" if args.sync_readback else "pipelined (copy stream, overlaps the next frame)")) if world > 1 else ("zero-copy stores from the trace kernel into the pinned frame" if args.zero_copy else ("pipelined: device snapshot + copy stream, the copy of frame k overlaps the trace of frame k+1, every frame delivered inside the timed region" if pipelined else "blocking cudaMemcpyAsync after the kernel")),
                "frame_matches_device": e2e_frame_ok,
                "note": "camera state in (launch parameters), packed LDR frame out to pinned host memory, wall clock"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if band:
        line["e2e_reference_call_pattern"] = band
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.workload, rt, scene)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
