"""Developer probe: throughput of the secondary modes (bounce rays, 4K x 16 spp, other scenes) with the default kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
def run(name, w, h, spp, rec, accel=rt.ACCEL_BVH, reps=5, wavefront=1):
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=rec,sub_spread=1,jitter_mode=rt.JITTER_HASHED,seed=1,accel=accel))
    r.set_tuning(6, wavefront)
    for _ in range(3): r.trace_rows(0,h,spp)
    ms=[]; rays=None
    for _ in range(reps):
        npri, nsh = r.trace_rows(0,h,spp)
        st = r.launch_stats(); ms.append(st['trace_kernel_ms']); rays = npri + nsh + st['n_bounce']
    m = float(np.median(ms))
    print(f'{name:9s} {w}x{h} spp {spp} recursions {rec}: {m:.3f} ms, {rays} rays (bounce {st["n_bounce"]}), {rays/m/1e3:.0f} Mrays/s', flush=True)
    r.close()
for wf in (0, 1):
    print('bounce wavefront', wf)
    run('thai2',1920,1080,1,2, wavefront=wf)
    run('ico2',1024,768,1,2, wavefront=wf)
    run('ico3_tex',1920,1080,1,2, wavefront=wf)
    run('thai2',1920,1080,1,2, accel=rt.ACCEL_OCTREE, wavefront=wf)
