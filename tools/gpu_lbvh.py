"""Developer probe: GPU tree build time (CUDA events) and trace time with the GPU-built tree vs the host SAH tree."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
for name, w, h in [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    line = f'{name:9s}'
    for accel, an in [(rt.ACCEL_BVH,'sah(host)'),(rt.ACCEL_LBVH,'lbvh(gpu)')]:
        t0 = time.perf_counter()
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
        create_ms = (time.perf_counter()-t0)*1e3
        extra = ''
        if accel == rt.ACCEL_LBVH:
            b = [r.lbvh_build() for _ in range(5)]
            extra = ' build_ms %s depth %d' % ([round(x['build_ms'],3) for x in b], b[-1]['depth'])
        npri, nsh = r.trace_rows(0,h,1)
        ts=[]
        for i in range(30):
            r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        ms=float(np.median(ts[12:]))
        line += f' | {an}: create {create_ms:.1f} ms trace {ms:.4f} ms {(npri+nsh)/ms/1e3:6.0f} Mr/s{extra}'
        r.close()
    print(line, flush=True)
