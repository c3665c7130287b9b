"""Profiling target: tools/profile_step.py on another build of the library. Usage: python tools/profile_lib.py librt_<name>.so [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200.api as api
if sys.argv[1] != "default":
    path = os.path.join(ROOT, "tools", sys.argv[1])
    api.lib_path = lambda: path
import raytracer_rs_b200 as rt
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(1920, 1080, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
ms = []
for i in range(frames):
    t.trace_rows(0, 1080, 1)
    ms.append(round(t.launch_stats()["trace_kernel_ms"], 4))
print(sys.argv[1], "kernel ms", ms)
