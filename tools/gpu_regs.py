import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200.api as api
lib = sys.argv[1]
api.lib_path = lambda: os.path.join(ROOT, lib)
import raytracer_rs_b200 as rt
for name,w,h in [('ico2',1024,768),('thai2',1920,1080)]:
    s = rt.load_scene(os.path.join(ROOT,f'data/{name}.dae'))
    for accel,an in [(rt.ACCEL_BVH,'bvh'),(rt.ACCEL_OCTREE,'octree')]:
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
        ts=[]
        for i in range(30):
            r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        print(lib, name, an, 'ms %.4f'%float(np.median(ts[8:])), flush=True)
