"""raytracer_rs_b200 — B200-native (sm_100a) replacement for raytracer_lib's per-pixel render loop.

Python mirror of the reference crate's public surface (raytracer_lib/src/lib.rs:5-27):

    create_raytracer(collada_doc, triangles_per_leaf, width, height) -> RayTracer
    create_raytracer_from_file(collada_filename, triangles_per_leaf, width, height) -> RayTracer
    RayTracer.trace_frame_additive() -> int          (raytracer/mod.rs:80)
    RayTracer.get_tonemapped_pixels() -> np.ndarray  (raytracer/mod.rs:120)
    RayTracer.camera.{move_rel, add_x_angle, add_y_angle}, RayTracer.film.clear()
    stats.Stats, DEFAULT_TRIANGLES_PER_LEAF, timing.BenchMark

Everything goes through the C ABI of include/rt_b200.h (raytracer_rs_b200/librt_b200.so, built by
`__graft_entry__.build()` or `make -C raytracer_rs_b200/csrc`). There is no CPU fallback: importing works
without a GPU (so the ABI can be inspected), rendering raises RtError.
"""
from .api import (  # noqa: F401
    DEFAULT_TRIANGLES_PER_LEAF,
    ACCEL_BVH,
    ACCEL_BVH4,
    ACCEL_CWBVH,
    ACCEL_LBVH,
    ACCEL_OCTREE,
    JITTER_FIXED_HALF,
    JITTER_HASHED,
    BenchMark,
    Config,
    RayTracer,
    RtError,
    Scene,
    Stats,
    create_raytracer,
    create_raytracer_from_file,
    lib,
    lib_path,
    load_scene,
    version,
    kernels_hash,
)

__all__ = [
    "DEFAULT_TRIANGLES_PER_LEAF",
    "ACCEL_BVH",
    "ACCEL_BVH4",
    "ACCEL_CWBVH",
    "ACCEL_LBVH",
    "ACCEL_OCTREE",
    "JITTER_FIXED_HALF",
    "JITTER_HASHED",
    "BenchMark",
    "Config",
    "RayTracer",
    "RtError",
    "Scene",
    "Stats",
    "create_raytracer",
    "create_raytracer_from_file",
    "lib",
    "lib_path",
    "load_scene",
    "version",
    "kernels_hash",
]
