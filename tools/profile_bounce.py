"""Profiling target: bounce-wavefront frames (RECURSIONS = 2, mod.rs:81-82) of thai2 1080p through the C ABI.
Usage: python tools/profile_bounce.py [frames] [tune key=value,...]   (run under ncu -k regex:wf_ per B200_PROFILING.md)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
w, h = 1920, 1080
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH))
for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
    k, v = kv.split("=")
    t.set_tuning(int(k), int(v))
ms = []
for _ in range(frames):
    n = t.trace_rows(0, h, 1)
    ms.append(round(t.launch_stats()["trace_kernel_ms"], 4))
print("bounce frames", n, t.launch_stats(), "frame ms", ms)
t.close()
