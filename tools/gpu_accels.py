"""Developer probe: trace time of every acceleration structure / kernel variant on the four BASELINE scenes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
scenes = [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]
cases = [('bvh', rt.ACCEL_BVH, 1), ('bvh/pool', rt.ACCEL_BVH, 2), ('bvh/1thr', rt.ACCEL_BVH, 0), ('octree', rt.ACCEL_OCTREE, 1),
         ('bvh4', rt.ACCEL_BVH4, 1), ('cwbvh', rt.ACCEL_CWBVH, 1), ('lbvh', rt.ACCEL_LBVH, 1)]
for name,w,h in scenes:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    line=f'{name:9s}'
    for cn, accel, variant in cases:
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
        r.set_tuning(0, variant)
        npri, nsh = r.trace_rows(0,h,1)
        ts=[]
        for i in range(40):
            r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        ms=float(np.median(ts[12:]))
        line += f' | {cn}: {ms:.4f} ms {(npri+nsh)/ms/1e3:6.0f}'
        r.close()
    print(line, flush=True)
