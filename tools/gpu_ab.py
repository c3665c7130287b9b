"""Developer A/B: tuning values on every workload; checks frames are identical to the first configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
scenes = [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]
# list of {tuning key: value}
variants = [{}, {7: 3}, {7: 5}, {7: 6}, {7: 8}, {7: 12}]
if len(sys.argv) > 1:
    scenes = [s for s in scenes if s[0] in sys.argv[1].split(',')]
for name,w,h in scenes:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    ref=None; line=f'{name:9s}'
    for v in variants:
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
        for k, val in v.items(): r.set_tuning(k, val)
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
