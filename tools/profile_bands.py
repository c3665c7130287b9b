"""Profiling target: the reference's band loop (50-row calls) on thai2 1080p with per-band tile schedules.
Usage: python tools/profile_bands.py [frames] [tune key=value,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
w, h = 1920, 1080
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
    k, v = kv.split("=")
    t.set_tuning(int(k), int(v))
calls = (h + 49) // 50
for f in range(frames):
    ms = []
    for _ in range(calls):
        t.trace_frame_additive()
        st = t.launch_stats()
        ms.append((round(st["trace_kernel_ms"], 4), st["kernels_launched"]))
    print("frame", f, ms)
t.close()
