"""Developer probe: the reference's band loop (trace_frame_additive 50 rows + readback, main.rs:200-201) on thai2 1080p under
different schedule settings, full vs incremental readback; kernel time per band from the library's CUDA events."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracer_rs_b200 as rt
w, h = 1920, 1080
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
host = torch.empty(w * h, dtype=torch.int32).pin_memory()
calls = (h + 49) // 50
rays_frame = 2599194 * (calls * 50 / h)
for label, tune in (("default: lap traced ahead", {}), ("band by band, static parts", {19: 0}), ("band by band, whole tiles", {19: 0, 1: 0}),
                    ("band by band, schedules", {19: 0, 11: 1024})):
    t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
    for k, v in tune.items():
        t.set_tuning(k, v)
    for _ in range(4 * calls):  # four frames: the schedules of all bands are learned
        t.trace_frame_additive()
    kms = []
    for _ in range(calls):
        t.trace_frame_additive()
        kms.append(t.launch_stats()["trace_kernel_ms"])
    res = {}
    for mode in ("full", "delta", "none"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5 * calls):
            t.trace_frame_additive()
            if mode == "full":
                t.get_tonemapped_pixels_into(host.data_ptr())
            elif mode == "delta":
                t.get_tonemapped_pixels_delta_into(host.data_ptr())
        torch.cuda.synchronize()
        res[mode] = rays_frame * 5 / (time.perf_counter() - t0) / 1e6
    print(f"{label:26s} band kernel ms: sum {sum(kms):.3f} max {max(kms):.4f} min {min(kms):.4f} | Mrays/s full readback {res['full']:.0f}, delta {res['delta']:.0f}, no readback {res['none']:.0f}")
    t.close()
