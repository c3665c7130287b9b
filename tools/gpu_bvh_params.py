"""Developer probe: SAH leaf size / triangle cost of the host BVH (env RT_BVH_MAX_LEAF, RT_BVH_TRI_COST, read by the library at build time)
against the frame time of the default frame (recursions 0) and the RECURSIONS = 2 frame. One process per setting (the env is read once).
Usage: python tools/gpu_bvh_params.py [leaf:cost ...]            (parent: runs the grid)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import raytracer_rs_b200 as rt
    import hashlib
    out = []
    for fname, w, h in (("thai2.dae", 1920, 1080), ("ico3_tex.dae", 1920, 1080)):
        scene = rt.load_scene(os.path.join(ROOT, "data", fname))
        for rec in (0, 2):
            t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=rec, sub_spread=1, jitter_mode=rt.JITTER_HASHED if rec else rt.JITTER_FIXED_HALF,
                                                         accel=rt.ACCEL_BVH))
            ms = []
            for _ in range(14):
                t.trace_rows(0, h, 1)
                ms.append(t.launch_stats()["trace_kernel_ms"])
            dig = hashlib.sha1(t.get_tonemapped_pixels().tobytes()).hexdigest()[:8]
            out.append("%s rec%d %.4f ms %s" % (fname[:-4], rec, float(np.median(ms[4:])), dig))
            t.close()
    print(" | ".join(out))
else:
    grid = [tuple(float(x) for x in a.split(':')) for a in sys.argv[1:]] or [(l, c) for l in (2, 3, 4, 6, 8) for c in (0.6, 1.2, 2.0)]
    for leaf, cost in grid:
        leaf = int(leaf)
        if True:
            env = dict(os.environ, RT_BVH_MAX_LEAF=str(leaf), RT_BVH_TRI_COST=str(cost))
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print("leaf<=%d tri_cost %.1f: %s" % (leaf, cost, r.stdout.strip() or r.stderr[-300:]), flush=True)
