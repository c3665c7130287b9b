"""Profiling target: a handful of pinned-mode frames of one workload through the C ABI (no torch, no oracle).
Usage: python tools/profile_step.py [workload] [accel] [frames]      (run under ncu per B200_PROFILING.md)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt

workloads = {"thai2_1080p": ("thai2.dae", 1920, 1080), "ico2_1024x768": ("ico2.dae", 1024, 768),
             "4boxes_1080p": ("4boxes.dae", 1920, 1080), "ico3_tex_1080p": ("ico3_tex.dae", 1920, 1080)}
wl = sys.argv[1] if len(sys.argv) > 1 else "thai2_1080p"
accel = sys.argv[2] if len(sys.argv) > 2 else "bvh"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 6
f, w, h = workloads[wl]
scene = rt.load_scene(os.path.join(ROOT, "data", f))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF,
                                             accel=rt.ACCEL_BVH if accel == "bvh" else rt.ACCEL_OCTREE))
ms = []
for i in range(frames):
    n_primary, n_shadow = t.trace_rows(0, h, 1)
    ms.append(t.launch_stats()["trace_kernel_ms"])
t.get_tonemapped_pixels()
print(wl, accel, "rays/frame", n_primary + n_shadow, "kernel ms", [round(x, 4) for x in ms])
