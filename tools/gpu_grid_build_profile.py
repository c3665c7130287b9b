"""ncu target: a few camera moves with the grid built at once (RT_TUNE_CAMERA_GRID_AFTER = 0), so that the launch list shows the build kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
s = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
r = rt.RayTracer.from_scene(s, rt.Config(1920, 1080, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
r.set_tuning(23, 0)
for i in range(6):
    r.camera.move_rel(0.0, 0.0, 0.002 if i % 2 else -0.002)
    r.trace_rows(0, 1080, 1)
    r.trace_rows(0, 1080, 1)
r.close()
