"""Developer probe: refill / inner-loop-yield thresholds of the ray-stream kernel (forced on) against the lockstep wavefront, RECURSIONS = 2 frame."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
for fname, w, h in (("thai2.dae", 1920, 1080), ("ico3_tex.dae", 1920, 1080), ("ico2.dae", 1024, 768)):
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    rows = []
    for label, tune in [("stream r%d m%d" % (r, m), {13: 1, 14: r, 15: m}) for r in (8, 12, 16, 20, 24) for m in (4, 8, 12)] + [("lockstep", {13: 0})]:
        t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH))
        for k, v in tune.items():
            t.set_tuning(k, v)
        ms = []
        for _ in range(11):
            t.trace_rows(0, h, 1, want_shadow=False)
            ms.append(t.launch_stats()["trace_kernel_ms"])
        rows.append("%s %.4f" % (label, float(np.median(ms[3:]))))
        t.close()
    print(fname, " | ".join(rows), flush=True)
