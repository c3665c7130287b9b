"""Writes profiles/traffic.json from an `ncu --set full` report of the default trace kernel: the two per-launch counts bench.py quotes
(DRAM bytes, warp instructions) plus lanes per instruction, stamped with the digest of the kernel sources of the CURRENT build (the
report must have been captured with this build: run right after the capture, before touching csrc/kernels.cu).
Usage: python tools/update_traffic.py gpurun_out/<report>.ncu-rep [workload/accel]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_rs_b200 as rt
rep = sys.argv[1]
key = sys.argv[2] if len(sys.argv) > 2 else "thai2_1080p/bvh"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, v = rows[0], rows[-1]
def metric(name, unit_scale=None):
    x = float(v[hdr.index(name)].replace(",", ""))
    unit = rows[1][hdr.index(name)]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    return x * scale
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    d = json.load(open(path))
except Exception:
    d = {}
d["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum and smsp__thread_inst_executed_per_inst_executed.ratio of ONE launch "
                 "of the trace kernel (ncu --set full, full 1080p frame, learned tile schedule), written by tools/update_traffic.py; kernels_hash = digest of "
                 "csrc/kernels.cu + csrc/device_types.h of the build the counts were measured on (rt_kernels_hash of that library)")
d["kernels_hash"] = rt.kernels_hash()
d["report"] = os.path.basename(rep)
d["kernel"] = v[hdr.index("Kernel Name")]
d[key] = int(metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum"))
d.setdefault("warp_instructions", {})[key] = int(metric("smsp__inst_executed.sum"))
d.setdefault("lanes_per_instruction", {})[key] = round(metric("smsp__thread_inst_executed_per_inst_executed.ratio"), 2)
d["gpu_time_us_under_ncu"] = metric("gpu__time_duration.sum")
for k in ("_comment_warp_instructions",):
    d.pop(k, None)
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
