"""Developer probe: cost of a camera move (perspective-grid rebuild + first frame) with the grid on / off, wall clock per frame over a run
of moves, everything synchronous (trace_rows reads the counters back)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
s = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
for g, after in ((0, 1), (3, 1), (3, 0), (2, 0)):
    r = rt.RayTracer.from_scene(s, rt.Config(1920, 1080, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH))
    r.set_tuning(22, g)
    r.set_tuning(23, after)
    for _ in range(5):
        r.trace_rows(0, 1080, 1)
    t0 = time.perf_counter()
    for _ in range(200):
        r.trace_rows(0, 1080, 1)
    still = (time.perf_counter() - t0) / 200
    ks = []
    t0 = time.perf_counter()
    for i in range(200):
        r.camera.move_rel(0.0, 0.0, 0.002 if i % 2 else -0.002)
        r.trace_rows(0, 1080, 1)
        ks.append(r.launch_stats()["trace_kernel_ms"])
    moving = (time.perf_counter() - t0) / 200
    print("grid %d, built after %d launches: static camera %.4f ms per synchronous frame, camera moved every frame %.4f ms (trace events %.4f ms)" % (g, after, still * 1e3, moving * 1e3, float(np.median(ks))))
    r.close()
