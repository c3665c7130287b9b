"""Developer probe (debug build with -DRT_DEBUG_WARP_EXIT): when does every warp of the ray-pool kernel run out of work?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200.api as api
api.lib_path = lambda: os.path.join(ROOT, 'tools', 'librt_b200_dbg.so')
import raytracer_rs_b200 as rt
w,h=1920,1080
s = rt.load_scene(os.path.join(ROOT,'data/thai2.dae'))
for lpt in (1,0):
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
    r.set_tuning(1, lpt)
    for i in range(12):
        r.trace_rows(0,h,1)
    ms = r.launch_stats()['trace_kernel_ms']
    raw = r.get_primary_ids()
    nw = 444*8
    ex = raw[:nw].astype(np.int64); st = raw[nw:2*nw].astype(np.int64)
    t0 = st.min()
    ex = (ex - t0) & 0xffffffff; st = (st - t0) & 0xffffffff
    print('lpt',lpt,'kernel ms',ms,'start spread us', st.max()/1e3, 'exit us: min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f' % tuple(np.percentile(ex,[0,10,50,90,99,100])/1e3))
    per_sm = ex.reshape(444,8).max(1)
    print('   block exit us p10 %.1f p50 %.1f p90 %.1f max %.1f; mean warp busy fraction %.3f' % (*(np.percentile(per_sm,[10,50,90,100])/1e3), ex.mean()/ex.max()))
    r.close()
