import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import raytracer_rs_b200 as rt
dev = torch.device('cuda',0)
scene = rt.load_scene(os.path.join(ROOT,'data/thai2.dae'))
W,H=1920,1080
t = rt.RayTracer.from_scene(scene, rt.Config(W,H,recursions=0,jitter_mode=rt.JITTER_FIXED_HALF,accel=rt.ACCEL_BVH))
stream = torch.cuda.Stream(device=dev); t.set_stream(stream.cuda_stream)
flush = torch.empty(256<<20, dtype=torch.uint8, device=dev)
small = torch.empty(1<<20, dtype=torch.uint8, device=dev)
def run(label, do_flush, n=50, sync_each=False, flush_t=flush):
    evs=[]
    with torch.cuda.stream(stream):
        for i in range(n):
            if do_flush: flush_t.zero_()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record(stream); t.trace_rows(0,H,1,want_shadow=False); e1.record(stream); evs.append((e0,e1))
            if sync_each: stream.synchronize()
    torch.cuda.synchronize()
    ms=[a.elapsed_time(b) for a,b in evs][5:]
    kms = t.launch_stats()['trace_kernel_ms']
    print(f"{label:40s} step mean {np.mean(ms):.4f} min {np.min(ms):.4f} max {np.max(ms):.4f} | last kernel_ms {kms:.4f}", flush=True)
run('no flush, async', False)
run('no flush, sync each', False, sync_each=True)
run('flush 256MB, async', True)
run('flush 256MB, sync each', True, sync_each=True)
run('flush 1MB (small), async', True, flush_t=small)
# time pieces: a trivial torch op between events
with torch.cuda.stream(stream):
    evs=[]
    for i in range(30):
        flush.zero_()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(stream); small.zero_(); e1.record(stream); evs.append((e0,e1))
torch.cuda.synchronize()
print('tiny kernel after flush:', np.mean([a.elapsed_time(b) for a,b in evs][5:]))
