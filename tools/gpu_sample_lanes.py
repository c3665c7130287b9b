"""A/B of the multi-sample launch forms (RT_TUNE_MULTI_SAMPLE_LAUNCH) on ONE GPU, in the shape a rank of a weak-scaling
run sees: shard 0 of N (interleaved 8-row bands) of thai2 1920x1080 with N hashed samples per pixel per pass, and the
4K x 16 spp frame. Prints the median pass time (library CUDA events around the whole rt_trace_rows call, which includes
the accumulation pass of the plane form) and checks that the film is bit-identical between the forms.
Usage: python tools/gpu_sample_lanes.py [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
FORMS = {1: "sample lanes", 2: "planes + accumulate", 0: "launch per sample"}


def run(w, h, shards, spp, reps):
    films = {}
    for form, name in FORMS.items():
        t = rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=rt.JITTER_HASHED, accel=rt.ACCEL_BVH,
                                                     shard_index=0, shard_count=shards, band_rows=8))
        t.set_tuning(5, form)
        ms = []
        for i in range(reps):
            n_primary, n_shadow = t.trace_rows(0, h, spp)
            ms.append(t.launch_stats()["trace_kernel_ms"])
        t.film.clear()
        t.trace_rows(0, h, spp)
        films[form] = t.film.pixel_datas().view(np.uint32).copy()
        ms = sorted(ms[3:])
        print("%dx%d shard 1/%d spp %2d  %-20s median %.4f ms  min %.4f ms  (%d primary + ~%d shadow rays, %d kernels per pass)"
              % (w, h, shards, spp, name, ms[len(ms) // 2], ms[0], n_primary, n_shadow, t.launch_stats()["kernels_launched"]), flush=True)
        t.close()
    same = all(np.array_equal(films[1], films[f]) for f in (2, 0))
    print("    films bit-identical across the three forms:", same, flush=True)
    return same


ok = True
for n in (2, 4, 8):
    ok &= run(1920, 1080, n, n, reps)
ok &= run(1920, 1080, 1, 1, reps)
ok &= run(3840, 2160, 1, 16, max(6, reps // 5))
sys.exit(0 if ok else 1)
