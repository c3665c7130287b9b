import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200.api as api
api.lib_path = lambda: os.path.join(ROOT, 'tools', 'librt_b200_dbg.so')
import raytracer_rs_b200 as rt
w,h=1920,1080
s = rt.load_scene(os.path.join(ROOT,'data/thai2.dae'))
r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
r.trace_rows(0,h,1)
raw = r.get_primary_ids().reshape(h,w)
nodes = (raw & 0xffff).astype(np.float64); tris = (raw>>16).astype(np.float64)
print('nodes/pixel: mean %.1f p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f' % (nodes.mean(), *np.percentile(nodes,[50,90,99,99.9]), nodes.max()))
print('tris/pixel:  mean %.1f p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f' % (tris.mean(), *np.percentile(tris,[50,90,99,99.9]), tris.max()))
tn = nodes.reshape(h//4,4,w//8,8)
tile_max = tn.max(axis=(1,3)); tile_sum = tn.sum(axis=(1,3))
print('per tile: sum of nodes mean %.0f; max-lane nodes mean %.1f p99 %.0f max %.0f; SIMD bound = sum/(32*max) mean over busy tiles %.3f' % (tile_sum.mean(), tile_max.mean(), np.percentile(tile_max,99), tile_max.max(), (tile_sum[tile_max>0]/(32*tile_max[tile_max>0])).mean()))
print('total nodes %.3g total tris %.3g ; total if every tile ran at its max lane: nodes %.3g' % (nodes.sum(), tris.sum(), 32*tile_max.sum()))
for (ty,tx) in [(34,140),(33,139),(149,76),(3,136)]:
    print('tile',ty,tx,'nodes', nodes[ty*4:ty*4+4, tx*8:tx*8+8].astype(int).tolist(), 'tris', tris[ty*4:ty*4+4, tx*8:tx*8+8].astype(int).tolist())
np.save(os.path.join(ROOT,'gpurun_out','steps_nodes.npy'), nodes.astype(np.uint16)); np.save(os.path.join(ROOT,'gpurun_out','steps_tris.npy'), tris.astype(np.uint16))
