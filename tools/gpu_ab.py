"""Developer A/B: kernel variants on every workload; checks frames are identical between variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
scenes = [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]
# (variant, lpt schedule, pool_refill, pool_min_inner, accel)
variants = [(1,1,8,16,rt.ACCEL_BVH),(2,1,16,8,rt.ACCEL_BVH),(2,1,8,8,rt.ACCEL_BVH),(2,1,24,8,rt.ACCEL_BVH)]
if len(sys.argv) > 1:
    scenes = [s for s in scenes if s[0] in sys.argv[1].split(',')]
for name,w,h in scenes:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    ref=None; line=f'{name:9s}'
    for v in variants:
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=v[4]))
        r.set_tuning(0, v[0]); r.set_tuning(1, v[1]); r.set_tuning(2, v[2]); r.set_tuning(4, v[3])
        npri, nsh = r.trace_rows(0,h,1)
        frame = (r.get_primary_ids(), r.get_tonemapped_pixels(), r.film.pixel_datas())
        if ref is None: ref = (frame, nsh)
        same = all(np.array_equal(a.view(np.uint32),b.view(np.uint32)) for a,b in zip(ref[0],frame)) and nsh == ref[1]
        ts=[]
        for i in range(40):
            r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        ms=float(np.median(ts[12:]))
        line += f' | {v}: {ms:.4f} {(npri+nsh)/ms/1e3:6.0f} {"ok" if same else "DIFF"}'
        r.close()
    print(line, flush=True)
