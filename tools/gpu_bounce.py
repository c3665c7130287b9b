"""Developer probe: RECURSIONS = 2 frame (the reference's default mode, mod.rs:81-82) on the BASELINE scenes: ray-stream levels with several
refill thresholds vs the lockstep wavefront vs depth first. Frame time from the library's CUDA events, median of 8."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
for fname, w, h in (("thai2.dae", 1920, 1080), ("ico3_tex.dae", 1920, 1080), ("ico2.dae", 1024, 768)):
    scene = rt.load_scene(os.path.join(ROOT, "data", fname))
    settings = [("stream chain %d blocks %d refill %d min_inner %d" % (c, b, r, m), {18: c, 16: b, 14: r, 15: m})
                for c in (1, 0) for b in (4, 5) for (r, m) in ((16, 8), (8, 4), (24, 8))]
    for label, tune in settings + [("lockstep, 3 blocks/SM", {13: 0, 17: 3}), ("lockstep, occupancy", {13: 0}), ("depth first", {6: 0})]:
        t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH))
        for k, v in tune.items():
            t.set_tuning(k, v)
        ms = []
        for _ in range(11):
            n = t.trace_rows(0, h, 1)
            ms.append(t.launch_stats()["trace_kernel_ms"])
        st = t.launch_stats()
        rays = st["n_primary"] + st["n_shadow"] + st["n_bounce"]
        m = float(np.median(ms[3:]))
        print(f"{fname:13s} {label:44s} frame {m:.4f} ms  {rays / m / 1e3:.0f} Mrays/s (primary+shadow+bounce {rays})  kernels {st['kernels_launched']}")
        t.close()
