"""Developer probe: the same frames with differently built libraries (tools/librt_<name>.so from tools/build_variants.sh): median trace
kernel time (the library's CUDA events, warm L2) and a digest of film + frame, which must not depend on the build."""
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
for name,w,h in [('thai2',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080)]:
    s = rt.load_scene(os.path.join(%r, 'data/%%s.dae' %% name))
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
    npri,nsh = r.trace_rows(0,h,1)
    ts=[]
    for i in range(40):
        r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
    dig = hashlib.sha256(r.film.pixel_datas().tobytes() + r.get_tonemapped_pixels().tobytes()).hexdigest()[:8]
    out.append('%%s %%.4f ms %%s' %% (name, float(np.median(ts[12:])), dig))
    r.close()
print(' | '.join(out))
''' % (ROOT, ROOT)
for lib in [None] + sys.argv[1:]:
    env = dict(os.environ)
    if lib: env['RT_LIB'] = os.path.join(ROOT, 'tools', lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print('%-24s' % (lib or 'default'), '->', r.stdout.strip() or r.stderr[-400:], flush=True)
