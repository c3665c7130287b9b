"""Developer probe: the same frame with differently built libraries (tools/librt_*.so)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import raytracer_rs_b200.api as api
lib = os.environ.get("RT_LIB")
if lib: api.lib_path = lambda: lib
import raytracer_rs_b200 as rt
out=[]
for name,w,h in [('thai2',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080)]:
    s = rt.load_scene(os.path.join(%r, 'data/%%s.dae' %% name))
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
    npri,nsh = r.trace_rows(0,h,1)
    ts=[]
    for i in range(40):
        r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
    out.append('%%s %%.4f ms' %% (name, float(np.median(ts[12:]))))
    r.close()
print(' | '.join(out))
''' % (ROOT, ROOT)
for lib in [None] + sys.argv[1:]:
    env = dict(os.environ)
    if lib: env['RT_LIB'] = os.path.join(ROOT, 'tools', lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(lib or 'default', '->', r.stdout.strip() or r.stderr[-400:], flush=True)
