"""Developer probe: collapse parameters of the 4-wide BVH vs trace time (env overrides RT_BVH4_*)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import raytracer_rs_b200 as rt
out=[]
for name,w,h in [('thai2',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080)]:
    s = rt.load_scene(os.path.join(%r, 'data/%%s.dae' %% name))
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH4))
    npri,nsh = r.trace_rows(0,h,1)
    ts=[]
    for i in range(40):
        r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
    out.append('%%s %%.4f ms nodes %%d' %% (name, float(np.median(ts[12:])), r.bvh4_stats()['nodes']))
    r.close()
print(' | '.join(out))
''' % (ROOT, ROOT)
for base in (2, 4):
    for leaf in (2, 4, 6):
        for cost in (0.3, 0.6, 1.0, 1.5):
            env = dict(os.environ, RT_BVH4_MAX_LEAF=str(leaf), RT_BVH4_PRIM_COST=str(cost), RT_BVH4_BASE_LEAF=str(base), RT_BVH_TRI_COST='0.6' if base == 2 else '1.2')
            r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
            print('base', base, 'leaf', leaf, 'prim_cost', cost, '->', r.stdout.strip() or r.stderr[-300:], flush=True)
