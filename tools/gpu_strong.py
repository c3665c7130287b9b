"""Developer probe: what ONE rank of a strong-scaling step runs (its interleaved 8-row bands of the 1-spp thai2 1080p frame) on one GPU,
for different heavy-tile split limits. Kernel time from the library's CUDA events, median of 20 after the schedule is learned."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:  # another build of the library (tools/librt_<name>.so)
    import raytracer_rs_b200.api as api
    _path = os.path.join(ROOT, "tools", sys.argv[1])
    api.lib_path = lambda: _path
import raytracer_rs_b200 as rt
w, h = 1920, 1080
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
print("library:", rt.lib_path())
for n in (1, 2, 4, 8):
    out = []
    for label, tune in (("split<=4", {12: 1}), ("split<=8", {12: 2}), ("split<=16", {12: 3}), ("no split", {12: 0}), ("image order", {1: 0})):
        t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH, shard_index=0, shard_count=n, band_rows=8))
        for k, v in tune.items():
            t.set_tuning(k, v)
        ms = []
        for _ in range(30):
            t.trace_rows(0, h, 1, want_shadow=False)
            ms.append(t.launch_stats()["trace_kernel_ms"])
        out.append("%s %.4f" % (label, float(np.median(ms[10:]))))
        t.close()
    print("shard 0 of %d:" % n, " | ".join(out), flush=True)
# critical path of the shard one rank of 8 owns (split <= 16): what bounds the launch
t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH, shard_index=0, shard_count=8, band_rows=8))
ms = []
for _ in range(30):
    t.trace_rows(0, h, 1, want_shadow=False)
    ms.append(t.launch_stats()["trace_kernel_ms"])
cost, items = t.tile_costs()
mhz = 1965.0
warps = 148 * 3 * 8
print("shard 0 of 8, split <= 16: kernel %.4f ms | %d tiles -> %d queue items | sum of tile costs / %d resident warps = %.1f us | heaviest tile (parts x slowest part) %.1f us, 99th percentile %.1f us, median %.2f us (cycles at %d MHz)"
      % (float(np.median(ms[10:])), len(cost), items, warps, cost.sum() / warps / mhz, cost.max() / mhz, np.percentile(cost, 99) / mhz, np.median(cost) / mhz, mhz))
t.close()
