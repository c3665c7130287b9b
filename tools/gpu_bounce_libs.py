"""Developer probe: the RECURSIONS = 2 frame of thai2 / ico3_tex with differently built libraries (tools/librt_<name>.so), ray-stream kernel
forced on, 3 / 4 / 5 resident blocks per SM; a digest of the film must not depend on the build."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, hashlib
sys.path.insert(0, %r)
import numpy as np
import raytracer_rs_b200.api as api
lib = os.environ.get("RT_LIB")
if lib: api.lib_path = lambda: lib
import raytracer_rs_b200 as rt
out=[]
for name,w,h in [('thai2',1920,1080),('ico3_tex',1920,1080)]:
    s = rt.load_scene(os.path.join(%r, 'data/%%s.dae' %% name))
    for blocks in (3, 4, 5):
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=2,sub_spread=1,jitter_mode=rt.JITTER_HASHED,accel=rt.ACCEL_BVH))
        r.set_tuning(13, 1); r.set_tuning(16, blocks)
        ts=[]
        for i in range(11):
            r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        dig = hashlib.sha256(r.film.pixel_datas().tobytes()).hexdigest()[:6]
        out.append('%%s b%%d %%.4f %%s' %% (name, blocks, float(np.median(ts[3:])), dig))
        r.close()
print(' | '.join(out))
''' % (ROOT, ROOT)
for lib in [None] + sys.argv[1:]:
    env = dict(os.environ)
    if lib: env['RT_LIB'] = os.path.join(ROOT, 'tools', lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print('%-22s' % (lib or 'default'), '->', r.stdout.strip() or r.stderr[-400:], flush=True)
