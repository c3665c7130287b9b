"""Generates tests/golden/*.npz + manifest.json from the CPU oracle (oracle/rt_oracle.cpp).

These are ORACLE outputs, not reference outputs: the Rust reference cannot be built or run in this environment and
holds no golden images. They pin the oracle against drift and give the GPU suite fixed
vectors that do not need the oracle at run time.   Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import raytracer_rs_b200 as rt  # noqa: E402  (product loader flattens the scene files)
from oracle_lib import JITTER_FIXED, JITTER_HASHED, Oracle  # noqa: E402

W, H, SEED = 160, 90, 12345
manifest = {}
for name in ["4boxes", "ico2", "ico3_tex", "thai2"]:
    scene = rt.load_scene(os.path.join(ROOT, "data", name + ".dae"))
    o = Oracle(scene, W, H)
    o.configure(recursions=0, jitter=JITTER_FIXED)
    o.trace_rows(0, H, 1)
    ids, ldr, c = o.get_primary_ids(), o.get_tonemapped_pixels(), o.counters()
    o.film_clear()
    o.configure(recursions=0, jitter=JITTER_HASHED, seed=SEED)
    o.trace_rows(0, H, 2)
    np.savez_compressed(os.path.join(HERE, f"{name}_{W}x{H}.npz"), ids=ids, ldr=ldr, ldr_jitter2=o.get_tonemapped_pixels())
    manifest[name] = dict(width=W, height=H, seed=SEED, shadow_rays=c["rays"]["shadow"], primary_hits=c["primary_hits"],
                          triangles=int(scene.vertices.shape[0]))
    print(name, manifest[name])
json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
