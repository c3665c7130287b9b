import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200 as rt
w,h=1920,1080
s = rt.load_scene(os.path.join(ROOT,'data/thai2.dae'))
r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
def t(first, n, reps=12):
    ts=[]
    for i in range(reps):
        r.trace_rows(first,n,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
    return float(np.median(ts[4:]))
print('full frame', t(0,h))
for first,n in [(128,4),(128,12),(132,4),(136,4),(596,4),(596,20),(12,4),(0,64),(0,128),(0,256),(0,540),(700,380),(540,540)]:
    print(f'rows {first}..{first+n}: {t(first,n)*1000:.1f} us')
r.set_tuning(1,0)
print('image-order schedule: full', t(0,h), 'rows 128..132', t(128,4)*1000, 'us')
