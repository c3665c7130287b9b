"""Developer probe (not part of the product): parity + timing of both traversal kernels on every scene."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import raytracer_rs_b200 as rt
from oracle_lib import Oracle, JITTER_FIXED

out = {}
scenes = [('4boxes',1920,1080),('ico2',1024,768),('ico3_tex',1920,1080),('thai2',1920,1080)]
if len(sys.argv) > 1: scenes = [s for s in scenes if s[0] in sys.argv[1:]]
for name,w,h in scenes:
    s = rt.load_scene(os.path.join(ROOT, f'data/{name}.dae'))
    o = Oracle(s, w, h, 70); o.configure(recursions=0, jitter=JITTER_FIXED)
    t0=time.time(); o.trace_rows(0,h,1,threads=0); cpu_s=time.time()-t0
    ids_o = o.get_primary_ids(); ldr_o = o.get_tonemapped_pixels(); film_o = o.get_film()
    c = o.counters(); n_shadow_o = c['rays']['shadow']
    for accel, an in [(rt.ACCEL_OCTREE,'octree'),(rt.ACCEL_BVH,'bvh')]:
        r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
        npri, nsh = r.trace_rows(0,h,1)
        ids = r.get_primary_ids(); ldr = r.get_tonemapped_pixels(); film = r.film.pixel_datas()
        idmis = int((ids!=ids_o).sum())
        ch = lambda a,k: ((a>>k)&255).astype(np.int32)
        maxlsb = int(max(np.abs(ch(ldr,k)-ch(ldr_o,k)).max() for k in (0,8,16,24)))
        ldrmis = int((ldr!=ldr_o).sum())
        same_id = ids==ids_o
        maxlsb_sameid = int(max(np.abs(ch(ldr,k)-ch(ldr_o,k))[same_id].max() for k in (0,8,16)))
        film_bits = int((film.view(np.uint32)!=film_o.view(np.uint32)).any(axis=1).sum())
        # timing: 20 frames
        ts=[]
        for i in range(20):
            r.film.clear(); r.trace_rows(0,h,1,want_shadow=False); ts.append(r.launch_stats()['trace_kernel_ms'])
        ms = float(np.median(ts[3:]))
        rays = npri+nsh
        res = dict(id_mismatch=idmis, id_agree=1-idmis/ids.size, ldr_mismatch=ldrmis, max_lsb=maxlsb, max_lsb_same_id=maxlsb_sameid,
                   film_pixels_not_bit_equal=film_bits, n_primary=npri, n_shadow=nsh, n_shadow_oracle=n_shadow_o,
                   kernel_ms=ms, mrays_s=rays/ms/1e3, cpu_s_allcores=cpu_s)
        out[f'{name}/{an}'] = res
        print(name, an, json.dumps(res), flush=True)
        r.close()
os.makedirs(os.path.join(ROOT,'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT,'gpurun_out/probe.json'),'w'), indent=1)
