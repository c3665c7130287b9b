"""Profiling target for the launch list of the auxiliary kernels: GPU tree build, bounce wavefront, multi-sample pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
w, h = 1920, 1080
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_LBVH))
print("lbvh", t.lbvh_build())
t.trace_rows(0, h, 1)
t.close()
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH))
for _ in range(4):
    n = t.trace_rows(0, h, 1)
print("bounce frame", n, t.launch_stats())
t.close()
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH))
for _ in range(4):
    n = t.trace_rows(0, h, 4)
print("4 spp pass", n, t.launch_stats())
t.close()
