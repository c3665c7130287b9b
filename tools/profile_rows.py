"""Profiling target: a few launches over a narrow row range (latency-critical case: heavy tiles running almost alone)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
first, n = int(sys.argv[1]), int(sys.argv[2])
w, h = 1920, 1080
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
ms = []
for i in range(6):
    t.trace_rows(first, n, 1)
    ms.append(t.launch_stats()["trace_kernel_ms"])
print("rows", first, n, "kernel ms", [round(x, 4) for x in ms])
