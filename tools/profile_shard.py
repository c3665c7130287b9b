"""Profiling target: what ONE rank of an 8-GPU weak-scaling step runs (1/8 of the rows x 8 samples in one launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
w, h, n = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH, shard_index=0, shard_count=n, band_rows=8))
ms = []
for i in range(12):
    t.trace_rows(0, h, n, want_shadow=False)
    ms.append(round(t.launch_stats()["trace_kernel_ms"], 4))
print("shard 0 of", n, "pass ms (sample lanes: one trace launch; odd sample counts: trace + accumulate)", ms)
