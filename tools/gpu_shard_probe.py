"""Developer probe: what one rank of N runs in a weak-scaling step (its interleaved 8-row bands x N samples per pixel, hashed jitter, sample
lanes), timed alone on one GPU with the library's CUDA events; grids on / off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
s = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
for grid in (3, 0):
    for n in (1, 2, 4, 8):
        r = rt.RayTracer.from_scene(s, rt.Config(1920, 1080, recursions=0, jitter_mode=rt.JITTER_HASHED, seed=0, accel=rt.ACCEL_BVH,
                                                 shard_index=0, shard_count=n, band_rows=8))
        r.set_tuning(22, grid)
        ts = []
        for i in range(30):
            r.trace_rows(0, 1080, n, want_shadow=False)
            ts.append(r.launch_stats()["trace_kernel_ms"])
        st = r.launch_stats()
        print("grid %d  rank 0 of %d x %d spp: %.4f ms  (kernels per call %d)" % (grid, n, n, float(np.median(ts[10:])), st["kernels_launched"]), flush=True)
        r.close()
