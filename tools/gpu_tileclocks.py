import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_rs_b200.api as api
api.lib_path = lambda: os.path.join(ROOT, 'tools', 'librt_b200_dbg.so')
import raytracer_rs_b200 as rt
w,h=1920,1080
s = rt.load_scene(os.path.join(ROOT,'data/thai2.dae'))
for accel,an in [(rt.ACCEL_BVH,'bvh'),(rt.ACCEL_OCTREE,'octree')]:
    r = rt.RayTracer.from_scene(s, rt.Config(w,h,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=accel))
    r.trace_rows(0,h,1); r.trace_rows(0,h,1)
    clk = r.get_primary_ids().reshape(h,w)
    tiles = clk.reshape(h//4,4,w//8,8).max(axis=(1,3)).astype(np.float64)
    print(an, 'kernel ms', r.launch_stats()['trace_kernel_ms'], 'tiles', tiles.size, 'sum Mcycles', tiles.sum()/1e6, 'mean', tiles.mean(), 'p50', np.percentile(tiles,50), 'p90', np.percentile(tiles,90), 'p99', np.percentile(tiles,99), 'p99.9', np.percentile(tiles,99.9), 'max', tiles.max())
    idx = np.argsort(tiles.reshape(-1))[::-1][:12]
    print('  heaviest tiles (tile_y, tile_x, kcycles):', [(int(i//(w//8)), int(i%(w//8)), int(tiles.reshape(-1)[i]/1000)) for i in idx])
    # rows profile
    rowsum = tiles.sum(axis=1)
    print('  per tile-row Mcycles (every 10th):', [round(x/1e6,2) for x in rowsum[::10]])
    np.save(os.path.join(ROOT,'gpurun_out',f'tileclk_{an}.npy'), tiles.astype(np.float32))
