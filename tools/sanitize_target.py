"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
w, h = 200, 113  # not multiples of the tile size
for name in ("thai2", "ico3_tex"):
    s = rt.load_scene(os.path.join(ROOT, f"data/{name}.dae"))
    for accel in (rt.ACCEL_OCTREE, rt.ACCEL_BVH, rt.ACCEL_CWBVH, rt.ACCEL_BVH4, rt.ACCEL_LBVH):
        for variant in (0, 1, 2):
            if variant == 2 and accel != rt.ACCEL_BVH:
                continue
            t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_HASHED, seed=1, accel=accel))
            t.set_tuning(0, variant)
            t.trace_rows(0, h, 1)
            t.trace_rows(7, 50, 3)      # sample planes, wrapped range
            t.trace_frame_additive()
            t.get_tonemapped_pixels()
            t.film.clear()
            t.close()
        # bounce rays: wavefront and depth first
        for wf in (1, 0):
            t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, seed=1, accel=accel))
            t.set_tuning(6, wf)
            t.trace_rows(0, h, 2)
            t.get_primary_ids()
            t.close()
    # schedule with sort + split (needs >= 4096 tiles), sharded handle
    t = rt.RayTracer.from_scene(s, rt.Config(1024, 576, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH, shard_index=1, shard_count=2, band_rows=8))
    for _ in range(4):
        t.trace_rows(0, 576, 1)
    t.close()
    t = rt.RayTracer.from_scene(s, rt.Config(1024, 576, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
    for _ in range(4):
        t.trace_rows(0, 576, 1)
    t.lbvh_build()
    t.close()
print("sanitize target done")
