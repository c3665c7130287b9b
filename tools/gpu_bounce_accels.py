"""Developer probe: RECURSIONS = 2 frame on thai2 / ico3_tex with every acceleration structure (lockstep wavefront, occupancy-sized grids;
the binary BVH also as a ray stream). Frame time from the library's CUDA events, median of 8."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
for fname, w, h in (("thai2.dae", 1920, 1080), ("ico3_tex.dae", 1920, 1080)):
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    ref = None
    for label, accel, tune in (("bvh stream blocks 4", rt.ACCEL_BVH, {18: 0, 16: 4, 14: 16, 15: 8}), ("bvh lockstep", rt.ACCEL_BVH, {13: 0}),
                               ("lbvh lockstep", rt.ACCEL_LBVH, {13: 0}), ("bvh4 lockstep", rt.ACCEL_BVH4, {}), ("cwbvh lockstep", rt.ACCEL_CWBVH, {}),
                               ("octree lockstep", rt.ACCEL_OCTREE, {})):
        t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, accel=accel))
        for k, v in tune.items():
            t.set_tuning(k, v)
        ms = []
        for _ in range(11):
            t.trace_rows(0, h, 1)
            ms.append(t.launch_stats()["trace_kernel_ms"])
        st = t.launch_stats()
        film = t.film.pixel_datas().view(np.uint32)
        if ref is None:
            ref = film
        rays = st["n_primary"] + st["n_shadow"] + st["n_bounce"]
        m = float(np.median(ms[3:]))
        print(f"{fname:13s} {label:22s} frame {m:.4f} ms  {rays / m / 1e3:.0f} Mrays/s  film identical to the first row: {bool(np.array_equal(film, ref))}")
        t.close()
