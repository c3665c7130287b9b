"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (the CPU checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


class OrcMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rgb", C.c_float * 3), ("texture_id", C.c_uint32)]


class OrcLight(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("color", C.c_float * 3)]


class OrcTexture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb", C.POINTER(C.c_float))]


class OrcSceneDesc(C.Structure):
    _fields_ = [
        ("num_triangles", C.c_uint32),
        ("vertices", C.POINTER(C.c_float)),
        ("tri_geom", C.POINTER(C.c_uint32)),
        ("num_geometries", C.c_uint32),
        ("materials", C.POINTER(OrcMaterial)),
        ("num_lights", C.c_uint32),
        ("lights", C.POINTER(OrcLight)),
        ("num_textures", C.c_uint32),
        ("textures", C.POINTER(OrcTexture)),
        ("camera_orientation", C.c_float * 16),
        ("camera_fov_deg", C.c_float),
    ]


_lib = None


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "rt_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcSceneDesc), C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_configure.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_int]
        L.orc_camera_move_rel.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.orc_camera_add_x_angle.argtypes = [C.c_void_p, C.c_float]
        L.orc_camera_add_y_angle.argtypes = [C.c_void_p, C.c_float]
        L.orc_camera_get.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_camera_get_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.orc_film_clear.argtypes = [C.c_void_p]
        L.orc_trace_frame_additive.restype = C.c_uint32
        L.orc_trace_frame_additive.argtypes = [C.c_void_p, C.c_int]
        L.orc_trace_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_get_tonemapped_pixels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_primary_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_film.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_estimated_variances.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_film.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_octree_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_octree_export.restype = C.c_uint64
        L.orc_octree_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_intersect_cube_inverse_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_moller_trumbore.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_triangle_cube_intersection.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_collada_matrix_to_vecmath.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_matrix_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_matrix_mul_vec4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tonemap_pack.restype = C.c_uint32
        L.orc_tonemap_pack.argtypes = [C.c_void_p]
        L.orc_hash4.restype = C.c_uint32
        L.orc_hash4.argtypes = [C.c_uint32] * 4
        L.orc_sample_table.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


COUNTER_FIELDS = ["rays", "cube_tests", "tri_tests", "inner_nodes", "leaves", "leaf_rejects"]
RAY_KINDS = ["primary", "shadow", "bounce"]

JITTER_FIXED = 0
JITTER_HASHED = 1
ISECT_OCTREE = 0
ISECT_BRUTE = 1


class Oracle:
    """RayTracer of the CPU oracle (restated reference)."""

    def __init__(self, scene, width: int, height: int, triangles_per_leaf: int = 70):
        """`scene` has: vertices [T,9] f32, tri_geom [T] u32, materials [(kind,(r,g,b),tex)], lights [(pos,color)],
        textures [(w,h,rgb)], camera_orientation [16], camera_fov_deg."""
        L = lib()
        self._keep = []
        d = OrcSceneDesc()
        verts = np.ascontiguousarray(scene.vertices, dtype=np.float32)
        geom = np.ascontiguousarray(scene.tri_geom, dtype=np.uint32)
        self._keep += [verts, geom]
        d.num_triangles = verts.shape[0]
        d.vertices = verts.ctypes.data_as(C.POINTER(C.c_float))
        d.tri_geom = geom.ctypes.data_as(C.POINTER(C.c_uint32))
        mats = (OrcMaterial * max(1, len(scene.materials)))()
        for i, (kind, rgb, tex) in enumerate(scene.materials):
            mats[i].kind = int(kind)
            mats[i].rgb[:] = [float(x) for x in rgb]
            mats[i].texture_id = int(tex)
        d.num_geometries = len(scene.materials)
        d.materials = mats
        lights = (OrcLight * max(1, len(scene.lights)))()
        for i, (pos, col) in enumerate(scene.lights):
            lights[i].pos[:] = [float(x) for x in pos]
            lights[i].color[:] = [float(x) for x in col]
        d.num_lights = len(scene.lights)
        d.lights = lights
        texs = (OrcTexture * max(1, len(scene.textures)))()
        for i, (w, h, rgb) in enumerate(scene.textures):
            arr = np.ascontiguousarray(rgb, dtype=np.float32)
            self._keep.append(arr)
            texs[i].width, texs[i].height = int(w), int(h)
            texs[i].rgb = arr.ctypes.data_as(C.POINTER(C.c_float))
        d.num_textures = len(scene.textures)
        d.textures = texs
        d.camera_orientation[:] = [float(x) for x in scene.camera_orientation]
        d.camera_fov_deg = float(scene.camera_fov_deg)
        self._keep += [mats, lights, texs]
        self.width, self.height = width, height
        self.h = L.orc_create(C.byref(d), width, height, triangles_per_leaf)
        self.L = L

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, recursions=0, sub_spread=1, jitter=JITTER_FIXED, seed=0, intersector=ISECT_OCTREE):
        self.L.orc_configure(self.h, recursions, sub_spread, jitter, seed, intersector)

    def move_rel(self, x, y, z):
        self.L.orc_camera_move_rel(self.h, x, y, z)

    def add_x_angle(self, r):
        self.L.orc_camera_add_x_angle(self.h, r)

    def add_y_angle(self, r):
        self.L.orc_camera_add_y_angle(self.h, r)

    def camera(self):
        out = np.zeros(34, np.float32)
        self.L.orc_camera_get(self.h, _p(out))
        return out

    def get_ray(self, u, v, xi1=0.5, xi2=0.5):
        out = np.zeros(6, np.float32)
        self.L.orc_camera_get_ray(self.h, u, v, xi1, xi2, _p(out))
        return out

    def film_clear(self):
        self.L.orc_film_clear(self.h)

    def trace_frame_additive(self, threads=1):
        return self.L.orc_trace_frame_additive(self.h, threads)

    def trace_rows(self, first_row, n_rows, spp=1, threads=0):
        self.L.orc_trace_rows(self.h, first_row, n_rows, spp, threads)

    def get_tonemapped_pixels(self):
        out = np.zeros(self.width * self.height, np.uint32)
        self.L.orc_get_tonemapped_pixels(self.h, _p(out))
        return out

    def get_primary_ids(self):
        out = np.zeros(self.width * self.height, np.uint32)
        self.L.orc_get_primary_ids(self.h, _p(out))
        return out

    def get_film(self):
        out = np.zeros((self.width * self.height, 7), np.float32)
        self.L.orc_get_film(self.h, _p(out))
        return out

    def get_estimated_variances(self):
        """Film::get_estimated_variances (film.rs:50-67): [W*H, 3]"""
        out = np.zeros((self.width * self.height, 3), np.float32)
        self.L.orc_get_estimated_variances(self.h, _p(out))
        return out

    def set_film(self, film7, n=None):
        """overwrite the film: film7 [W*H, 7] as get_film returns it; n (uint32 per pixel) overrides column 6 when given"""
        a = np.ascontiguousarray(film7, np.float32)
        assert a.shape == (self.width * self.height, 7)
        nn = None if n is None else np.ascontiguousarray(n, np.uint32)
        self.L.orc_set_film(self.h, _p(a), _p(nn) if nn is not None else None)

    def counters(self, reset=False):
        raw = np.zeros(20, np.uint64)
        self.L.orc_get_counters(self.h, _p(raw), 1 if reset else 0)
        out = {}
        for fi, f in enumerate(COUNTER_FIELDS):
            out[f] = {k: int(raw[3 * fi + ki]) for ki, k in enumerate(RAY_KINDS)}
        out["primary_hits"] = int(raw[18])
        out["shadow_blocked"] = int(raw[19])
        return out

    def octree_stats(self):
        raw = np.zeros(6, np.uint64)
        self.L.orc_octree_stats(self.h, _p(raw))
        return dict(zip(["nodes", "inner", "leaves", "empty_leaves", "tri_refs", "depth"], (int(x) for x in raw)))

    def octree_export(self):
        n = self.octree_stats()["nodes"]
        refs = self.octree_stats()["tri_refs"]
        cubes = np.zeros((n, 6), np.float32)
        first_child = np.zeros(n, np.int32)
        leaf_offset = np.zeros(n + 1, np.uint32)
        leaf_tris = np.zeros(max(refs, 1), np.uint32)
        self.L.orc_octree_export(self.h, _p(cubes), _p(first_child), _p(leaf_offset), _p(leaf_tris))
        return cubes, first_child, leaf_offset, leaf_tris[:refs]

    def sample_table(self):
        out = np.zeros((65536, 3), np.float32)
        self.L.orc_sample_table(self.h, _p(out))
        return out
