"""Developer A/B: kernel variants on every workload; checks frames are identical between variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
scenes = [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]
variants = [(1,1)]
for name,w,h in scenes:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    for accel, an in [(rt.ACCEL_BVH,'bvh'),(rt.ACCEL_CWBVH,'cwbvh')]:
        ref=None; line=f'{name:9s} {an:7s}'
        for v in variants:
            r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
            r.set_tuning(0, v[0]); r.set_tuning(1, v[1])
            npri, nsh = r.trace_rows(0,h,1)
            frame = (r.get_primary_ids(), r.get_tonemapped_pixels(), r.film.pixel_datas())
            if ref is None: ref = frame
            same = all(np.array_equal(a.view(np.uint32),b.view(np.uint32)) for a,b in zip(ref,frame))
            ts=[]
            for i in range(30):
                r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
            ms=float(np.median(ts[5:]))
            line += f' | v{v[0]}{v[1]}: {ms:.4f} ms {(npri+nsh)/ms/1e3:8.0f} Mrays/s same={same}'
            r.close()
        print(line, flush=True)
