#!/usr/bin/env python
"""bench.py — throughput of the per-pixel render loop on B200 (BASELINE.json metric).

A step = one full frame (1 sample per pixel over all H rows) of data/thai2.dae at 1920x1080, primary + shadow
rays (recursions = 0, fixed sub-pixel offset 0.5: the pinned parity mode), accumulated into the device film.
  value    : Mrays/s (primary + shadow), scene/film resident in HBM, timed with CUDA events on the launching stream
  e2e      : same metric through the public API with HOST buffers: camera state in, 8.3 MB LDR frame out (pinned),
             every step, wall clock
  roofline : frac = the resource that binds the trace kernel — warp-instruction issue slots (ncu count of one launch /
             measured kernel time, against 4 schedulers x SMs x the SM clock sampled during the run); achieved / peak /
             traffic_equivalent = algorithmic bytes of the reference algorithm (24 B per cube test + 36 B per triangle
             test + 4 B per pixel) / measured kernel duration against the measured HBM copy peak
  cpu_baseline / --impl reference : the CPU oracle (C++ restatement of the reference, oracle/) on the host cores
Extra blocks on the same line (N = 1): value_long (>= 0.5 s of GPU time), e2e_reference_call_pattern (the reference's own
50-row band loop), first_frame_after_move_ms, configs (the other BASELINE configurations, each checked against an oracle band),
recursions2 (the reference's default RECURSIONS = 2 mode).

N > 1 (torchrun, one process per GPU): the frame is sharded by interleaved 8-row bands, the scene is replicated,
rank 0 receives the packed frame over NVLink (see --gather), no other exchange.
  --scaling weak (default): a step renders N samples per pixel of the frame (hashed jitter), so every rank traces
      H/N rows x N samples = as many camera rays as the single GPU does in its step; value = all rays of all ranks / time
  --scaling strong: a step is one sample per pixel whatever N is (each rank traces H/N rows)
  The weak line also carries a `strong` block (the fixed 1-spp frame over the N GPUs), `configs.thai2_4k_16spp` sharded over the
  N GPUs, and e2e.frame_matches_device: the gathered host frame compared bit for bit with rank 0's own unsharded render.
"""
from __future__ import annotations

import argparse
import hashlib
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
# Exact work of the REFERENCE algorithm (octree, triangles_per_leaf = 70) for one pinned-mode frame, counted by the oracle's
# per-ray counters (oracle/rt_oracle.cpp `Counters`; checked against the live oracle in the cpu_baseline leg of every run and
# by tests/test_host_logic.py): (primary rays, shadow rays, cube tests, triangle tests)
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


def host_threads() -> int:
    """Host threads this process may use. torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, so the OpenMP default
    is not a usable answer under torchrun; the affinity mask is."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def kernel_source_hash() -> str:
    """sha256 (16 hex digits) of the kernel sources the loaded library was built from is compiled into the library
    (rt_kernels_hash); this is the same digest computed from the sources in the tree."""
    h = hashlib.sha256()
    for name in ("kernels.cu", "device_types.h"):
        with open(os.path.join(ROOT, "raytracer_rs_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


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


def load_scene_for_oracle(fname):
    """The scene for the CPU arm comes from the independent numpy Collada reader (tests/collada_ref.py), so that neither
    `--impl reference` nor the oracle checks of the configs block depend on the product's loader (librt_b200.so is not even
    mapped into the reference arm)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from collada_ref import load_collada

    return load_collada(os.path.join(ROOT, "data", fname))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the Rust original cannot be built
    here) on all host threads, on the configuration the b200 arm runs with the same --gpus / --scaling. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)  # before liboracle.so (and with it libgomp) is loaded
    os.environ.setdefault("OMP_DYNAMIC", "FALSE")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import JITTER_FIXED, JITTER_HASHED, Oracle

    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    fname, w, h, spp = WORKLOADS[args.workload]
    if world > 1 and args.scaling == "weak":
        spp *= world  # the b200 arm's step at this N
    scene = load_scene_for_oracle(fname)
    orc = Oracle(scene, w, h, 70)  # DEFAULT_TRIANGLES_PER_LEAF (oct_tree_intersector.rs:12)
    orc.configure(recursions=0, jitter=JITTER_FIXED if spp == 1 else JITTER_HASHED, seed=0)
    # Bounded sample: a step is the whole frame unless steps + warmup whole frames would take longer than --reference-budget
    # seconds on this host; then a step is one of `parts` equal row ranges, rotating over the frame from step to step, so
    # that any `parts` consecutive steps cover the frame once (rays are counted exactly either way).
    t1 = time.perf_counter()
    probe_rows = max(8, h // 16)
    orc.trace_rows(h // 4, probe_rows, spp, threads=threads)
    t_frame = (time.perf_counter() - t1) * h / probe_rows
    parts = max(1, min(h // 8, int(-(-t_frame * (args.steps + args.warmup) // args.reference_budget))))
    rows = -(-h // parts)
    orc.film_clear()

    def step(i, n_threads):
        first = (i % parts) * rows
        orc.trace_rows(first, min(rows, h - first), spp, threads=n_threads)
        orc.get_tonemapped_pixels()

    for i in range(args.warmup):
        step(i, threads)
    orc.counters(reset=True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, threads)
    dt = time.perf_counter() - t0
    c = orc.counters(reset=True)
    rays = c["rays"]["primary"] + c["rays"]["shadow"]
    value = rays / dt / 1e6
    # the single thread the reference actually ships (one render thread, raytracer/src/main.rs:194), on one part of the frame
    t0 = time.perf_counter()
    orc.trace_rows(h // 4, max(8, min(rows, h // 8)), spp, threads=1)
    dt1 = time.perf_counter() - t0
    c1 = orc.counters(reset=True)
    single = (c1["rays"]["primary"] + c1["rays"]["shadow"]) / dt1 / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "reference scene fixture data/%s (the upstream repository's own scene), pinned camera" % fname,
        "config": {"workload": args.workload, "width": w, "height": h, "spp": spp, "recursions": 0,
                   "jitter": "fixed 0.5" if spp == 1 else "hashed seed 0",
                   "accel": "reference octree, triangles_per_leaf 70", "scene_loader": "tests/collada_ref.py (independent of the product)",
                   "step": ("one full frame" if parts == 1 else "%d rows (1/%d of the frame, rotating)" % (rows, parts)) + " + get_tonemapped_pixels"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": single,
                         "omp_num_threads_env_at_start": os.environ.get("OMP_NUM_THREADS"),
                         "sample": ("%d full frames" % args.steps if parts == 1 else
                                    "%d steps of %d rows each (1/%d of the frame, rotating over it)" % (args.steps, rows, parts))
                                   + " (%d rays) of the same workload, OpenMP over rows on %d threads" % (rays, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, rt, scene):
    """Oracle timed on the GPU box's host cores: all threads (value) and one thread (what the reference ships)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import JITTER_FIXED, Oracle

    _f, w, h, spp = WORKLOADS[workload]
    spp = min(spp, 1)  # bounded sample: one sample per pixel
    orc = Oracle(scene, w, h, rt.DEFAULT_TRIANGLES_PER_LEAF)
    orc.configure(recursions=0, jitter=JITTER_FIXED)
    threads = host_threads()
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


def oracle_band_check(rt, np, tracer_factory, fname, w, h, spp, band):
    """Parity of one configuration inside the benchmark run: a fresh handle renders the whole frame once, the oracle (scene from
    the independent loader) renders `band` = (first row, rows) with the same settings; returns primitive-id agreement and the
    largest 8-bit channel difference over the band."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import JITTER_FIXED, JITTER_HASHED, Oracle

    first, rows = band
    orc = Oracle(load_scene_for_oracle(fname), w, h, rt.DEFAULT_TRIANGLES_PER_LEAF)
    orc.configure(recursions=0, jitter=JITTER_FIXED if spp == 1 else JITTER_HASHED, seed=0)
    orc.trace_rows(first, rows, spp, threads=host_threads())
    t = tracer_factory()
    t.trace_rows(0, h, spp)
    sl = slice(first * w, (first + rows) * w)
    ids, ids_o = t.get_primary_ids()[sl], orc.get_primary_ids()[sl]
    ldr, ldr_o = t.get_tonemapped_pixels()[sl], orc.get_tonemapped_pixels()[sl]
    lsb = np.max([np.abs(((ldr >> k) & 255).astype(np.int32) - ((ldr_o >> k) & 255).astype(np.int32)) for k in (0, 8, 16, 24)], axis=0)
    t.close()
    # with several jittered samples per pixel the id buffer holds the LAST sample's hit; a sample on a grazing ray may pick the
    # neighbouring triangle in the BVH (true closest hit) vs the octree order: report the fraction of pixels within the bar
    return {"band_rows": [first, first + rows], "ids_agree": float((ids == ids_o).mean()), "max_lsb_diff": int(lsb.max()),
            "pixels_within_1_lsb": float((lsb <= 1).mean())}


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
    ap.add_argument("--no-extras", action="store_true", help="skip value_long, configs, recursions2, strong, first-frame-after-move blocks")
    ap.add_argument("--reference-budget", type=float, default=150.0,
                    help="--impl reference: seconds the whole run may take; longer runs trace a rotating part of the frame per step")
    ap.add_argument("--launch-timing", action="store_true",
                    help="developer: keep RT_TUNE_TIME_LAUNCHES = 1 (two CUDA events per trace call behind launch_stats().trace_kernel_ms) in the "
                         "device-timed and end-to-end legs; by default only the roofline leg, which needs that figure, records them")
    ap.add_argument("--e2e-gather", default="host", choices=["host", "nvlink"],
                    help="N>1, end-to-end leg: 'host' = every rank copies the rows it owns over its own PCIe link into one frame in shared, "
                         "page-locked host memory (multi_gpu.HostFrameGather); 'nvlink' = the frame is gathered on rank 0 over NVLink (--gather) "
                         "and copied to the host over rank 0's link alone")
    ap.add_argument("--fence", default="kernel", choices=["kernel", "memops"],
                    help="N>1, --gather peer: per-frame fence made of flag kernels (bounded waits, default) or of stream memory operations "
                         "(cuStreamWriteValue32 / cuStreamWaitValue32: no launch, no timeout; faster at 2 GPUs, slower at 8)")
    ap.add_argument("--fused-signal", action="store_true",
                    help="N>1, --gather peer: publish 'frame done' from the trace kernel's last warp out (rt_set_done_signal) instead of a signal launch "
                         "of its own; measured slower (every warp fences its peer stores at system scope), kept as an option")
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

    fname, W, H, base_spp = WORKLOADS[args.workload]
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    accel = {"bvh": rt.ACCEL_BVH, "octree": rt.ACCEL_OCTREE, "cwbvh": rt.ACCEL_CWBVH, "bvh4": rt.ACCEL_BVH4, "lbvh": rt.ACCEL_LBVH}[args.accel]

    grids_off = any(kv.strip() == "22=0" for kv in args.tune.split(","))  # RT_TUNE_CAMERA_GRID = 0: every ray walks the tree

    def jitter_for(spp):
        return rt.JITTER_FIXED_HALF if spp == 1 else rt.JITTER_HASHED

    def make_tracer(scene_, w, h, spp, sharded=True, recursions=0):
        t = rt.RayTracer.from_scene(scene_, rt.Config(w, h, recursions=recursions, jitter_mode=jitter_for(spp), seed=0, accel=accel, device=local_rank,
                                                      shard_index=rank if sharded else 0, shard_count=world if sharded else 1, band_rows=8))
        for kv in filter(None, args.tune.split(",")):
            k, v = kv.split("=")
            t.set_tuning(int(k), int(v))
        return t

    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    from raytracer_rs_b200 import multi_gpu

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        v = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)

    def run_legs(tracer, gather, w, h, spp, steps, warmup, e2e=True, kernel_timing=True):
        """device-timed leg (+ kernel duration for the roofline) and end-to-end leg of one configuration on the current handles"""
        host_frames = [torch.empty(w * h, dtype=torch.int32).pin_memory(), torch.empty(w * h, dtype=torch.int32).pin_memory()]
        pipelined = world == 1 and not args.sync_readback and not args.zero_copy
        tracer.set_tuning(10, 1 if args.launch_timing else 0)

        def device_step():
            if gather is not None:
                gather.begin_frame()
            tracer.trace_rows(0, h, spp, want_shadow=False)
            if gather is not None:
                gather.device_gather(release=True)  # device-timed leg: the frame stays on rank 0

        def global_ray_totals():
            t = tracer.ray_totals()  # exact device counters since the handle was created (synchronises this rank's stream)
            v = torch.tensor([t["primary"], t["shadow"]], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(v)
            return int(v[0]), int(v[1])

        def kernels_now():
            return tracer.kernels_launched() + (gather.kernels if gather else 0)

        def timed(n_steps):
            rays0, launches0 = global_ray_totals(), kernels_now()
            t_region0 = time.perf_counter()
            events = []
            with torch.cuda.stream(stream):
                for _ in range(n_steps):
                    flush.zero_()  # L2 flush between timed iterations (not inside the event pair)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    device_step()
                    e1.record(stream)
                    events.append((e0, e1))
            barrier()
            clocks = sampler.window(t_region0, time.perf_counter())
            launches = kernels_now() - launches0
            rays1 = global_ray_totals()
            # rays of the timed steps, counted on the device (with hashed jitter every step draws new samples, so the shadow-ray
            # count differs slightly from step to step)
            n_primary, n_shadow = (rays1[0] - rays0[0]) / n_steps, (rays1[1] - rays0[1]) / n_steps
            total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in events))
            return {"value": (n_primary + n_shadow) * n_steps / (total_ms * 1e-3) / 1e6, "ms_per_step": total_ms / n_steps, "clocks": clocks,
                    "launches": launches, "primary": n_primary, "shadow": n_shadow, "steps": n_steps}

        with torch.cuda.stream(stream):
            for _ in range(warmup):
                flush.zero_()
                device_step()
        barrier()
        sampler.ready.wait(timeout=10)
        res = {"dev": timed(steps)}
        rays_per_step = res["dev"]["primary"] + res["dev"]["shadow"]

        # ---- trace-kernel duration for the roofline (library's own CUDA events around the kernel, same stream) ----
        if kernel_timing:
            tracer.set_tuning(10, 1)
            kernel_ms = []
            with torch.cuda.stream(stream):
                for _ in range(20):
                    flush.zero_()
                    tracer.trace_rows(0, h, spp, want_shadow=False)
                    kernel_ms.append(tracer.launch_stats()["trace_kernel_ms"])
            tracer.set_tuning(10, 1 if args.launch_timing else 0)
            res["kernel_ms"] = max_over_ranks(float(np.mean(kernel_ms)))
        if not e2e:
            return res

        # ---- end-to-end leg: public API, host buffers, every step ---------------------------------------------------
        hgather = None
        if gather is not None and args.e2e_gather == "host":
            shm_seq[0] += 1
            hgather = multi_gpu.HostFrameGather(tracer, rank, world, dev, stream, "rtb200_%s_%d" % (os.environ.get("MASTER_PORT", "0"), shm_seq[0]))

        def e2e_step(i):
            # per-step input: camera state from the host (travels to the device as the kernel's launch parameters)
            tracer.camera.set_state(0.0, 0.0, (0.0, 0.0, 0.0))
            if hgather is not None:
                # every rank delivers the rows it owns over its own PCIe link (pipelined: the copy of frame i overlaps the trace of i+1)
                hgather.begin_frame()
                tracer.trace_rows(0, h, spp, want_shadow=False)
                hgather.publish()
                if rank == 0:
                    hgather.wait_frame(keep=1)
                return
            if gather is not None:
                gather.begin_frame()
            tracer.trace_rows(0, h, spp, want_shadow=False)
            if gather is not None:
                gather.device_gather()
            if pipelined:
                # hand frame i to the copy stream, THEN make sure frame i-1 (the other host buffer, copied while frame i traced) has
                # arrived: the copy engine always has the next frame queued
                tracer.get_tonemapped_pixels_async(host_frames[i & 1].data_ptr())
                tracer.wait_pixels(1)
            elif gather is not None and not args.sync_readback:
                if rank == 0:  # same pipelining on rank 0 of a multi-GPU run: hand frame i to the copy stream, then make sure
                    # frame i-1 (the other host buffer) has arrived
                    gather.read_frame_async(host_frames[i & 1])
                    gather.wait_frame(keep=1)
            elif rank == 0:
                if gather is not None:
                    gather.read_frame_into(host_frames[0])
                else:
                    tracer.get_tonemapped_pixels_into(host_frames[0].data_ptr())

        def e2e_drain():
            if hgather is not None:
                if rank == 0:
                    hgather.wait_frame()
            elif pipelined:
                tracer.wait_pixels()  # the last frame is delivered inside the timed region too
            elif gather is not None and rank == 0:
                gather.wait_frame()

        if gather is None and args.zero_copy:
            tracer.set_host_frame(host_frames[0].data_ptr())  # the kernel stores packed pixels straight into the pinned frame
        for i in range(warmup):
            e2e_step(i)
        e2e_drain()
        barrier()
        rays0 = global_ray_totals()
        t0 = time.perf_counter()
        for i in range(steps):
            e2e_step(i)
        e2e_drain()
        barrier()
        e2e_s = time.perf_counter() - t0
        rays1 = global_ray_totals()
        e2e_rays = (rays1[0] - rays0[0]) + (rays1[1] - rays0[1])
        e2e_s = max_over_ranks(e2e_s)
        res["e2e"] = {"value": e2e_rays / e2e_s / 1e6, "ms_per_step": e2e_s / steps * 1e3}

        # ---- the frame the end-to-end path delivers must be the frame the device holds ----
        frame_ok = None
        if gather is None:
            check = np.empty(w * h, np.uint32)
            tracer.set_host_frame(None)
            tracer.get_tonemapped_pixels(check)
            last = host_frames[(steps - 1) & 1] if pipelined else host_frames[0]
            frame_ok = bool(np.array_equal(check, last.numpy().view(np.uint32)))
        else:
            # N > 1: every rank clears its film, ONE more frame goes through the same gather + host copy, and rank 0 compares it bit
            # for bit with its own UNSHARDED render of that frame (same samples: a sample's number is the pixel's film count)
            tracer.film.clear()
            e2e_step(0)
            e2e_drain()
            barrier()
            if rank == 0:
                solo = make_tracer(scene_of[(w, h)], w, h, spp, sharded=False)
                solo.trace_rows(0, h, spp, want_shadow=False)
                got = np.array(hgather.frame(hgather.frame_no - 1)) if hgather is not None else host_frames[0].numpy().view(np.uint32)
                frame_ok = bool(np.array_equal(solo.get_tonemapped_pixels(), got))
                solo.close()
        if hgather is not None:
            res["e2e"]["kernels"] = hgather.kernels
            hgather.close()
            gather.rearm()
        res["e2e"]["host_gather"] = hgather is not None
        res["e2e"]["frame_matches_device"] = frame_ok
        res["e2e"]["pipelined"] = pipelined
        res["rays_per_step"] = rays_per_step
        return res

    scene_of = {(W, H): scene}
    shm_seq = [0]
    spp = base_spp * (world if (world > 1 and args.scaling == "weak") else 1)
    tracer = make_tracer(scene, W, H, spp)
    tracer.set_stream(stream.cuda_stream)
    gather = multi_gpu.FrameGather(tracer, rank, world, dev, stream, mode=args.gather, fused_signal=args.fused_signal, fence=args.fence) if world > 1 else None
    main_res = run_legs(tracer, gather, W, H, spp, args.steps, args.warmup)
    extras = not args.no_extras

    # ---- a long run of the same step (>= 0.5 s of GPU time): the driver-sized figure above lasts a few milliseconds ----
    long_res = None
    if extras:
        n_long = int(max(args.steps, min(20000, 0.5e3 / max(main_res["dev"]["ms_per_step"], 1e-3))))
        tracer.film.clear()
        long_res = run_legs(tracer, gather, W, H, spp, n_long, 3, e2e=False, kernel_timing=False)["dev"]

    # ---- N > 1, weak line: the fixed frame (1 sample per pixel) over the same N GPUs ----
    strong_res = None
    if extras and world > 1 and args.scaling == "weak":
        tracer.configure(recursions=0, jitter_mode=jitter_for(base_spp), seed=0, accel=accel)
        tracer.film.clear()
        strong_res = run_legs(tracer, gather, W, H, base_spp, args.steps, args.warmup, kernel_timing=True)
        tracer.configure(recursions=0, jitter_mode=jitter_for(spp), seed=0, accel=accel)

    timeouts = tracer.sync_timeouts() if world > 1 else 0
    timeouts = int(max_over_ranks(float(timeouts)))

    # ---- single-GPU extras on the headline handle ----
    band = first_move = None
    if world == 1 and spp == 1 and extras:
        # the reference's own call pattern: 50-row bands, frame readback after every band (main.rs:200-201)
        host_frame = torch.empty(W * H, dtype=torch.int32).pin_memory()
        clone = np.empty(W * H, np.uint32)
        tracer.set_rows_per_call(50)
        calls = (H + 49) // 50
        rays_frame = main_res["rays_per_step"] * (calls * 50 / H)

        def band_loop(readback, reps):
            for _ in range(reps * calls):
                tracer.trace_frame_additive()
                readback()

        def band_rate(readback):
            band_loop(readback, 3)  # every band's schedule is learned
            torch.cuda.synchronize(dev)
            reps = max(1, min(args.steps, 20))
            t0 = time.perf_counter()
            band_loop(readback, reps)
            return rays_frame * reps / (time.perf_counter() - t0) / 1e6

        def delta_and_clone():
            tracer.get_tonemapped_pixels_delta_into(host_frame.data_ptr())
            np.copyto(clone, host_frame.numpy().view(np.uint32))  # the fresh Vec<u32> the reference's signature returns (mod.rs:120)

        v_delta = band_rate(lambda: tracer.get_tonemapped_pixels_delta_into(host_frame.data_ptr()))
        check = np.empty(W * H, np.uint32)
        tracer.get_tonemapped_pixels(check)
        delta_ok = bool(np.array_equal(check, host_frame.numpy().view(np.uint32)))
        band = {"value": v_delta, "unit": UNIT, "calls_per_frame": calls, "d2h_bytes_per_call": 50 * W * 4,
                "readback": "rt_get_tonemapped_pixels_delta into one pinned frame the host keeps: only the 50 rows traced by the call are copied",
                "frame_matches_device": delta_ok,
                "value_with_host_vec_clone": band_rate(delta_and_clone),
                "value_full_frame_readback": band_rate(lambda: tracer.get_tonemapped_pixels_into(host_frame.data_ptr())),
                "d2h_bytes_per_call_full_frame": W * H * 4,
                "note": "trace_frame_additive (50 rows) + frame readback per call, synchronous, as main.rs:200-201 does"}

        # first frame after a camera key (main.rs:124-162 moves the camera and clears the film): the tile schedule of the old view
        # is kept for it; compared with the steady state of the same handle
        def frame_ms(prepare=None):
            ms = []
            for _ in range(8):
                if prepare:
                    prepare()
                    torch.cuda.synchronize(dev)
                with torch.cuda.stream(stream):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    tracer.trace_rows(0, H, 1, want_shadow=False)
                    e1.record(stream)
                torch.cuda.synchronize(dev)
                ms.append(e0.elapsed_time(e1))
            return float(np.median(ms))

        for _ in range(4):
            tracer.trace_rows(0, H, 1, want_shadow=False)
        steady = frame_ms()

        def key_press():
            for _ in range(3):  # the frames in between re-learn the schedule, as during interactive use
                tracer.trace_rows(0, H, 1, want_shadow=False)
            tracer.camera.move_rel(0.0, 0.0, 0.1)  # the 'w' key of the native binary (main.rs:125-127)
            tracer.film.clear()

        first_move = {"first_frame_after_move_ms": frame_ms(key_press), "steady_state_ms": steady,
                      "note": "median of 8: camera.move_rel(0, 0, 0.1) + film.clear(), then one frame, timed with CUDA events"}
        first_move["ratio"] = first_move["first_frame_after_move_ms"] / steady
        tracer.camera.set_state(0.0, 0.0, (0.0, 0.0, 0.0))
        tracer.film.clear()

    # ---- the other BASELINE configurations and the reference's default mode, in the same run ----
    configs, rec2, tree_walk = None, None, None
    if extras and args.workload == "thai2_1080p":
        configs = {}
        names = ["ico2_1024x768", "4boxes_1080p", "ico3_tex_1080p", "thai2_4k_16spp"] if world == 1 else ["thai2_4k_16spp"]
        if gather is not None:
            gather.close()
            gather = None
        tracer.close()
        tracer = None
        for name in names:
            f2, w2, h2, spp2 = WORKLOADS[name]
            sc2 = scene if f2 == fname else rt.load_scene(os.path.join(ROOT, "data", f2))
            scene_of[(w2, h2)] = sc2
            t2 = make_tracer(sc2, w2, h2, spp2)
            t2.set_stream(stream.cuda_stream)
            g2 = multi_gpu.FrameGather(t2, rank, world, dev, stream, mode=args.gather, fused_signal=args.fused_signal, fence=args.fence) if world > 1 else None
            r2 = run_legs(t2, g2, w2, h2, spp2, max(5, min(args.steps, 20)), 3, kernel_timing=False)
            entry = {"value": r2["dev"]["value"], "unit": UNIT, "ms_per_step": r2["dev"]["ms_per_step"], "e2e": r2["e2e"]["value"],
                     "e2e_frame_matches_device": r2["e2e"]["frame_matches_device"], "spp": spp2, "n_gpus": world,
                     "rays_per_step": r2["rays_per_step"]}
            if g2 is not None:
                g2.close()
            t2.close()
            if rank == 0:
                # parity inside the run: an oracle band through the middle of the geometry (Q1 squeezes it into the upper ~56 % of the rows)
                band_rows = (h2 // 4, 16 if spp2 > 1 else 48)
                entry["oracle_check"] = oracle_band_check(rt, np, lambda: make_tracer(sc2, w2, h2, spp2, sharded=False), f2, w2, h2, spp2, band_rows)
            barrier()
            configs[name] = entry
        if world == 1:
            # RECURSIONS = 2, SUB_SPREAD = 1, jittered: what the reference binary always runs (mod.rs:81-82); bounce wavefront
            t3 = rt.RayTracer.from_scene(scene, rt.Config(W, H, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, seed=0, accel=accel, device=local_rank))
            t3.set_stream(stream.cuda_stream)
            for _ in range(4):
                t3.trace_rows(0, H, 1, want_shadow=False)
            tot0 = t3.ray_totals()
            ev = []
            n3 = max(5, min(args.steps, 20))
            with torch.cuda.stream(stream):
                for _ in range(n3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    t3.trace_rows(0, H, 1, want_shadow=False)
                    e1.record(stream)
                    ev.append((e0, e1))
            torch.cuda.synchronize(dev)
            tot1 = t3.ray_totals()
            ms3 = sum(a.elapsed_time(b) for a, b in ev) / n3
            rays3 = {k: (tot1[k] - tot0[k]) / n3 for k in tot0}
            rec2 = {"frame_ms": ms3, "rays_per_frame": rays3, "grays_per_s_all_rays": sum(rays3.values()) / ms3 / 1e6,
                    "value_primary_shadow": (rays3["primary"] + rays3["shadow"]) / ms3 / 1e3, "unit": UNIT,
                    "note": "thai2 1080p, recursions 2, sub_spread 1, hashed jitter, 1 spp per step; primary + shadow + bounce rays"}
            t3.close()
            # the headline workload with every ray walking the tree (RT_TUNE_CAMERA_GRID = 0): what the perspective grids replace
            if accel in (rt.ACCEL_BVH, rt.ACCEL_LBVH) and not grids_off:
                t4 = rt.RayTracer.from_scene(scene, rt.Config(W, H, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, seed=0, accel=accel, device=local_rank))
                t4.set_stream(stream.cuda_stream)
                t4.set_tuning(22, 0)
                t4.set_tuning(10, 0)
                for _ in range(4):
                    t4.trace_rows(0, H, 1, want_shadow=False)
                tot0 = t4.ray_totals()
                ev = []
                with torch.cuda.stream(stream):
                    for _ in range(n3):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        t4.trace_rows(0, H, 1, want_shadow=False)
                        e1.record(stream)
                        ev.append((e0, e1))
                torch.cuda.synchronize(dev)
                tot1 = t4.ray_totals()
                ms4 = sum(a.elapsed_time(b) for a, b in ev) / n3
                # and the two paths render the same thing: one fresh sample each, frame and film compared bit for bit
                t4.film.clear()
                t4.trace_rows(0, H, 1)
                t5 = rt.RayTracer.from_scene(scene, rt.Config(W, H, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, seed=0, accel=accel, device=local_rank))
                t5.set_stream(stream.cuda_stream)
                for _ in range(3):  # (a view gets its grid on its second launch)
                    t5.trace_rows(0, H, 1)
                t5.film.clear()
                t5.trace_rows(0, H, 1)
                same = bool(np.array_equal(t4.get_tonemapped_pixels(), t5.get_tonemapped_pixels()) and
                            t4.film.pixel_datas().tobytes() == t5.film.pixel_datas().tobytes())
                tree_walk = {"value": (tot1["primary"] - tot0["primary"] + tot1["shadow"] - tot0["shadow"]) / n3 / ms4 / 1e3, "unit": UNIT, "ms_per_step": ms4,
                             "frame_and_film_equal_the_grid_path": same,
                             "note": "same workload and timing as `value`, camera and shadow rays through the BVH instead of the perspective grids"}
                t4.close()
                t5.close()

    sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if timeouts:
        raise SystemExit("bench.py: %d frame-fence waits timed out (rt_sync_timeouts): a frame may be torn, no number is reported" % timeouts)

    dev_res, e2e_res, kms = main_res["dev"], main_res["e2e"], main_res["kernel_ms"]
    clocks = dev_res["clocks"]
    rays_per_step = main_res["rays_per_step"]
    peak, peak_src = measured_hbm_peak()
    alg = algorithmic_bytes(args.workload)
    prof, traffic, warp_inst, stale = {}, None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            prof = json.load(f)
        traffic = prof.get("%s/%s" % (args.workload, args.accel))
        warp_inst = prof.get("warp_instructions", {}).get("%s/%s" % (args.workload, args.accel))
        # the ncu counts describe ONE build of the kernels: compare the digest recorded with them against the digest compiled into
        # the loaded library and against the sources in the tree
        lib_hash = rt.lib().rt_kernels_hash().decode()
        stale = not (prof.get("kernels_hash") == lib_hash == kernel_source_hash())
    except Exception:
        pass
    roofline = None
    if alg is not None:
        note = ("achieved/peak: traffic-equivalent of the reference algorithm's reads (the <= 2 MB scene is L1/L2 resident, so it exceeds the HBM "
                "peak); frac: the binding physical resource")
        if world == 1 and spp == 1:
            per_launch = float(alg)  # exact oracle counters of the pinned frame
        else:
            # jittered / sharded launches: the pinned frame's average bytes per ray x the rays one rank traces per launch
            p0, s0 = REFERENCE_WORK[args.workload][:2]
            per_launch = alg / (p0 + s0) * rays_per_step / world
            note += "; per-launch bytes estimated from the pinned frame's average bytes per ray"
        achieved = per_launch / (kms * 1e-3) / 1e9
        single = world == 1 and spp == 1
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "traffic_equivalent": achieved / peak,
                    "traffic": traffic if single else None, "peak_source": peak_src, "kernel": "trace_shade_persistent_kernel<%s>" % ("bvh + perspective grids" if args.accel in ("bvh", "lbvh") and not grids_off else args.accel),
                    "kernel_ms": kms, "algorithmic_bytes_per_launch": per_launch, "note": note, "stale": stale,
                    "counters_from": "profiles/traffic.json (ncu --set full capture of one launch; kernels_hash %s)" % prof.get("kernels_hash")}
        if roofline["traffic"]:
            # what actually crosses the HBM interface (ncu dram__bytes of one launch, profiles/traffic.json): the film
            roofline["dram_achieved"] = roofline["traffic"] / (kms * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_achieved"] / peak
        if warp_inst and single and clocks.get("sm_mhz"):
            # the resource that actually binds: warp instructions issued (ncu smsp__inst_executed.sum of one launch) / measured
            # kernel time, against 4 schedulers x SMs x the SM clock measured during the run
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
            roofline["issue_slots"] = {"warp_instructions_per_launch": warp_inst, "achieved_ginst_s": warp_inst / (kms * 1e-3) / 1e9,
                                       "peak_ginst_s": peak_issue / 1e9, "frac": warp_inst / (kms * 1e-3) / peak_issue,
                                       "lanes_active_per_instruction": prof.get("lanes_per_instruction", {}).get("%s/%s" % (args.workload, args.accel))}
            roofline["frac"] = roofline["issue_slots"]["frac"]
            roofline["frac_of"] = "SM issue slots over the whole launch (one warp instruction per scheduler and cycle)"
        else:
            roofline["frac"] = roofline["traffic_equivalent"]
            roofline["frac_of"] = "traffic equivalent (no instruction count recorded for this configuration)"
    gather_note = "" if world == 1 else (" (%d spp per GPU-count unit: %s scaling), rows sharded by interleaved 8-row bands, every rank traces its rows x all "
                                          "samples in one launch, packed frame stored into rank 0's buffer over NVLink (%s, frame-done signal %s, fence: %s)"
                                          % (base_spp, args.scaling, args.gather, "fused into the trace kernel" if args.fused_signal else "from a separate launch", args.fence))
    readback = (("every rank copies the rows it owns over its own PCIe link into one frame in shared page-locked host memory (copy stream, overlaps "
                 "the next frame; arrival flags in the same shared memory)") if e2e_res.get("host_gather") else
                ("gather to rank 0 (%s) + %s copy to pinned host memory" % (args.gather, "blocking" if args.sync_readback else "pipelined (copy stream, overlaps the next frame)"))) if world > 1 else (
        "zero-copy stores from the trace kernel into the pinned frame" if args.zero_copy else (
            "pipelined: device snapshot + copy stream, the copy of frame k overlaps the trace of frame k+1, every frame delivered inside the timed region"
            if e2e_res["pipelined"] else "blocking cudaMemcpyAsync after the kernel"))
    line = {
        "metric": METRIC,
        "value": dev_res["value"],
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dev_res["ms_per_step"],
        "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "reference scene fixture data/%s (the upstream repository's own scene), pinned camera" % fname,
        "config": {"workload": args.workload, "width": W, "height": H, "spp": spp, "recursions": 0,
                   "jitter": "fixed 0.5" if spp == 1 else "hashed seed 0", "accel": args.accel,
                   "ray_index": ("none: every ray walks the tree" if grids_off or args.accel not in ("bvh", "lbvh") else
                                 "camera rays through the perspective grid of the view, shadow rays through the cube of grids around the light "
                                 "(built on the device; same hits as the tree walk, bit for bit), bounce rays and everything else through the tree"),
                   "l2_flush_between_steps": True,
                   "step": "one full frame: trace %d rows x %d spp%s" % (H, spp, gather_note),
                   "primary_rays_per_step": dev_res["primary"], "shadow_rays_per_step": dev_res["shadow"]},
        "frames_per_s": 1e3 / dev_res["ms_per_step"],
        "clocks": clocks,
        "e2e": {"value": e2e_res["value"], "unit": UNIT, "h2d_bytes_per_step": int(rt.lib().rt_launch_param_bytes()), "d2h_bytes_per_step": W * H * 4 + 32,
                "ms_per_step": e2e_res["ms_per_step"], "readback": readback, "frame_matches_device": e2e_res["frame_matches_device"],
                "sync_timeouts": timeouts,
                "note": "camera state in (launch parameters), packed LDR frame out to pinned host memory, wall clock"},
        "gpu_launches": int(dev_res["launches"]),
        "roofline": roofline,
    }
    if long_res:
        line["value_long"] = {"value": long_res["value"], "unit": UNIT, "steps": long_res["steps"], "ms_per_step": long_res["ms_per_step"],
                              "gpu_seconds": long_res["ms_per_step"] * long_res["steps"] * 1e-3, "clocks": long_res["clocks"]}
    if strong_res:
        line["strong"] = {"value": strong_res["dev"]["value"], "unit": UNIT, "ms_per_step": strong_res["dev"]["ms_per_step"], "spp": base_spp,
                          "kernel_ms": strong_res["kernel_ms"], "e2e": strong_res["e2e"],
                          "note": "the fixed frame (%d sample per pixel) sharded over the %d GPUs: strong scaling of BASELINE configs[3]" % (base_spp, world)}
    if band:
        line["e2e_reference_call_pattern"] = band
    if first_move:
        line["first_frame_after_move"] = first_move
    if configs:
        line["configs"] = configs
    if rec2:
        line["recursions2"] = rec2
    if tree_walk is not None:
        line["value_tree_walk"] = tree_walk
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.workload, rt, scene)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
