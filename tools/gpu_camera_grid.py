"""Developer probe: camera rays through the perspective grid (RT_TUNE_CAMERA_GRID = 2..5) and shadow rays through the light grids (RT_TUNE_LIGHT_GRID = 6..9) against the BVH walk (0): warm trace-kernel time and
a digest of film + ids + frame, which must not depend on the setting."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
for name, w, h in (("thai2", 1920, 1080), ("ico2", 1024, 768), ("ico3_tex", 1920, 1080), ("4boxes", 1920, 1080)):
    s = rt.load_scene(os.path.join(ROOT, "data", name + ".dae"))
    for rec, jit in ((0, rt.JITTER_FIXED_HALF), (0, rt.JITTER_HASHED), (2, rt.JITTER_HASHED)):
        out = []
        for g, lg in ((0, 0), (3, 0), (3, 7), (3, 8), (3, 9), (2, 8)):
            r = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=rec, sub_spread=1, jitter_mode=jit, seed=7, accel=rt.ACCEL_BVH))
            r.set_tuning(22, g)
            r.set_tuning(24, lg)
            r.trace_rows(0, h, 1)
            ts = []
            for i in range(24):
                r.trace_rows(0, h, 1, want_shadow=False)
                ts.append(r.launch_stats()["trace_kernel_ms"])
            dig = hashlib.sha256(r.film.pixel_datas().tobytes() + r.get_primary_ids().tobytes() + r.get_tonemapped_pixels().tobytes()).hexdigest()[:8]
            out.append("grid %d/%d: %.4f ms %s" % (g, lg, float(np.median(ts[8:])), dig))
            r.close()
        print("%-9s rec %d jitter %d | %s" % (name, rec, jit, " | ".join(out)), flush=True)
