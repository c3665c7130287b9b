"""ctypes binding of include/rt_b200.h plus a thin object layer that mirrors raytracer_lib's API.

No rendering logic lives here: every method is one call into librt_b200.so. If the shared library is missing
the import fails loudly (no eager / CPU fallback exists anywhere in this package).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

DEFAULT_TRIANGLES_PER_LEAF = 70  # oct_tree_intersector.rs:12, re-exported lib.rs:7
DEVICE_NONE = -2  # RT_DEVICE_NONE: host-side handle for CPU-only tests of the host logic
ACCEL_OCTREE, ACCEL_BVH, ACCEL_CWBVH, ACCEL_BVH4, ACCEL_LBVH = 0, 1, 2, 3, 4
JITTER_FIXED_HALF, JITTER_HASHED = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))


class RtError(RuntimeError):
    """Error string returned by the library (the reference returns Result<_, String>, lib.rs:15-27)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"rt_b200 error {code}: {message}")
        self.code = code
        self.message = message


# ---- C structs -------------------------------------------------------------------------------------------


class CMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rgb", C.c_float * 3), ("texture_id", C.c_uint32)]


class CLight(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("color", C.c_float * 3)]


class CTexture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb", C.POINTER(C.c_float))]


class CSceneDesc(C.Structure):
    _fields_ = [
        ("num_triangles", C.c_uint32),
        ("vertices", C.POINTER(C.c_float)),
        ("tri_geom", C.POINTER(C.c_uint32)),
        ("num_geometries", C.c_uint32),
        ("materials", C.POINTER(CMaterial)),
        ("num_lights", C.c_uint32),
        ("lights", C.POINTER(CLight)),
        ("num_textures", C.c_uint32),
        ("textures", C.POINTER(CTexture)),
        ("camera_orientation", C.c_float * 16),
        ("camera_fov_deg", C.c_float),
        ("normals", C.POINTER(C.c_float)),  # reserved, may be NULL: ignored as the reference ignores them (colladaloader.rs:587-593)
        ("uvs", C.POINTER(C.c_float)),
    ]


class CConfig(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("triangles_per_leaf", C.c_uint32),
        ("rows_per_call", C.c_uint32),
        ("recursions", C.c_int32),
        ("sub_spread", C.c_uint32),
        ("jitter_mode", C.c_int32),
        ("seed", C.c_uint32),
        ("accel", C.c_int32),
        ("device", C.c_int32),
        ("shard_index", C.c_uint32),
        ("shard_count", C.c_uint32),
        ("band_rows", C.c_uint32),
    ]


class CLaunchStats(C.Structure):
    _fields_ = [
        ("kernels_launched", C.c_uint32),
        ("trace_kernel_ms", C.c_float),
        ("n_primary", C.c_uint64),
        ("n_shadow", C.c_uint64),
        ("n_bounce", C.c_uint64),
    ]


# every symbol include/rt_b200.h declares (tests/test_abi.py checks the header against this list)
ABI_SYMBOLS = [
    "rt_config_default", "rt_scene_load_file", "rt_scene_load_str", "rt_scene_get_desc", "rt_scene_free",
    "rt_create_raytracer", "rt_create_raytracer_from_file", "rt_create", "rt_destroy", "rt_last_error",
    "rt_configure", "rt_set_rows_per_call", "rt_trace_frame_additive", "rt_trace_rows",
    "rt_get_tonemapped_pixels", "rt_get_tonemapped_pixels_delta", "rt_get_tonemapped_pixels_async", "rt_wait_pixels", "rt_wait_pixels_keep", "rt_film_clear", "rt_get_film",
    "rt_set_film", "rt_get_estimated_variances", "rt_get_primary_ids", "rt_camera_move_rel",
    "rt_camera_add_x_angle", "rt_camera_add_y_angle", "rt_camera_get", "rt_camera_set_state", "rt_get_camera_plane_matrix", "rt_set_stream",
    "rt_get_ldr_device_ptr", "rt_set_ldr_target", "rt_get_owned_ldr_rows_device", "rt_get_launch_stats",
    "rt_device_alloc", "rt_device_free", "rt_ipc_export", "rt_ipc_open", "rt_ipc_close", "rt_get_counters_device_ptr",
    "rt_stream_signal_flag", "rt_stream_signal_then_wait", "rt_stream_wait_flags", "rt_sync_timeouts", "rt_set_done_signal", "rt_stream_write_value", "rt_stream_wait_value",
    "rt_launch_param_bytes", "rt_set_tuning", "rt_set_host_frame", "rt_host_register", "rt_host_unregister", "rt_copy_owned_rows",
    "rt_signal_flag_on_stream",
    "rt_kernels_launched", "rt_get_ray_totals", "rt_get_tile_costs", "rt_octree_stats", "rt_octree_export", "rt_bvh_stats", "rt_bvh_export", "rt_bvh4_stats", "rt_bvh4_export", "rt_lbvh_build", "rt_lbvh_export", "rt_cwbvh_stats",
    "rt_cwbvh_export", "rt_stats_new",
    "rt_stats_free", "rt_stats_stats", "rt_stats_mean_stats", "rt_benchmark_new", "rt_benchmark_free",
    "rt_benchmark_start", "rt_benchmark_stop", "rt_benchmark_report", "rt_version", "rt_kernels_hash",
]  # fmt: skip

_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, "librt_b200.so")


def lib() -> C.CDLL:
    """Loads librt_b200.so. Raises ImportError when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C raytracer_rs_b200/csrc`. raytracer_rs_b200 has no CPU or eager fallback."
        )
    L = C.CDLL(path)
    vp, u32, i32, f32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_float, C.c_uint64
    sz, cp = C.c_size_t, C.c_char_p
    P = C.POINTER
    sig = {
        "rt_config_default": (None, [P(CConfig), u32, u32]),
        "rt_scene_load_file": (C.c_int, [cp, P(vp), cp, sz]),
        "rt_scene_load_str": (C.c_int, [cp, cp, P(vp), cp, sz]),
        "rt_scene_get_desc": (C.c_int, [vp, P(CSceneDesc)]),
        "rt_scene_free": (None, [vp]),
        "rt_create_raytracer": (C.c_int, [cp, sz, sz, sz, P(vp), cp, sz]),
        "rt_create_raytracer_from_file": (C.c_int, [cp, sz, sz, sz, P(vp), cp, sz]),
        "rt_create": (C.c_int, [P(CSceneDesc), P(CConfig), P(vp), cp, sz]),
        "rt_destroy": (None, [vp]),
        "rt_last_error": (cp, [vp]),
        "rt_configure": (C.c_int, [vp, i32, u32, i32, u32, i32]),
        "rt_set_rows_per_call": (C.c_int, [vp, u32]),
        "rt_trace_frame_additive": (C.c_int, [vp, P(u32)]),
        "rt_trace_rows": (C.c_int, [vp, u32, u32, u32, P(u64), P(u64)]),
        "rt_get_tonemapped_pixels": (C.c_int, [vp, vp]),
        "rt_get_tonemapped_pixels_delta": (C.c_int, [vp, vp]),
        "rt_get_tonemapped_pixels_async": (C.c_int, [vp, vp]),
        "rt_wait_pixels": (C.c_int, [vp]),
        "rt_wait_pixels_keep": (C.c_int, [vp, u32]),
        "rt_film_clear": (C.c_int, [vp]),
        "rt_get_film": (C.c_int, [vp, vp]),
        "rt_set_film": (C.c_int, [vp, vp]),
        "rt_get_estimated_variances": (C.c_int, [vp, vp]),
        "rt_get_primary_ids": (C.c_int, [vp, vp]),
        "rt_camera_move_rel": (C.c_int, [vp, f32, f32, f32]),
        "rt_camera_add_x_angle": (C.c_int, [vp, f32]),
        "rt_camera_add_y_angle": (C.c_int, [vp, f32]),
        "rt_camera_get": (C.c_int, [vp, vp]),
        "rt_get_camera_plane_matrix": (C.c_int, [vp, vp]),
        "rt_camera_set_state": (C.c_int, [vp, f32, f32, P(f32)]),
        "rt_set_stream": (C.c_int, [vp, vp]),
        "rt_get_ldr_device_ptr": (C.c_int, [vp, P(vp)]),
        "rt_set_ldr_target": (C.c_int, [vp, vp]),
        "rt_get_owned_ldr_rows_device": (C.c_int, [vp, vp, P(u32)]),
        "rt_get_launch_stats": (C.c_int, [vp, P(CLaunchStats)]),
        "rt_device_alloc": (C.c_int, [vp, sz, P(vp)]),
        "rt_device_free": (C.c_int, [vp, vp]),
        "rt_ipc_export": (C.c_int, [vp, vp, vp]),
        "rt_ipc_open": (C.c_int, [vp, vp, P(vp)]),
        "rt_ipc_close": (C.c_int, [vp, vp]),
        "rt_get_counters_device_ptr": (C.c_int, [vp, P(vp)]),
        "rt_stream_signal_flag": (C.c_int, [vp, vp, u32]),
        "rt_stream_signal_then_wait": (C.c_int, [vp, vp, u32, vp, u32]),
        "rt_stream_wait_flags": (C.c_int, [vp, vp, u32, u32, i32, i32]),
        "rt_sync_timeouts": (C.c_int, [vp, P(u32)]),
        "rt_set_done_signal": (C.c_int, [vp, vp, u32]),
        "rt_stream_write_value": (C.c_int, [vp, vp, u32]),
        "rt_stream_wait_value": (C.c_int, [vp, vp, u32]),
        "rt_launch_param_bytes": (u32, []),
        "rt_set_tuning": (C.c_int, [vp, i32, i32]),
        "rt_set_host_frame": (C.c_int, [vp, vp]),
        "rt_host_register": (C.c_int, [vp, vp, sz, P(vp)]),
        "rt_host_unregister": (C.c_int, [vp, vp]),
        "rt_copy_owned_rows": (C.c_int, [vp, vp, vp, vp]),
        "rt_signal_flag_on_stream": (C.c_int, [vp, vp, u32, vp]),
        "rt_kernels_launched": (u64, [vp]),
        "rt_octree_stats": (C.c_int, [vp, vp]),
        "rt_octree_export": (C.c_int, [vp, vp, vp, vp, vp, P(u64)]),
        "rt_bvh_stats": (C.c_int, [vp, vp]),
        "rt_bvh_export": (C.c_int, [vp, vp, vp, vp, vp]),
        "rt_get_ray_totals": (C.c_int, [vp, vp]),
        "rt_get_tile_costs": (C.c_int, [vp, vp, u32, P(u32), P(u32)]),
        "rt_lbvh_build": (C.c_int, [vp, vp, P(f32)]),
        "rt_lbvh_export": (C.c_int, [vp, vp, vp, vp, vp]),
        "rt_bvh4_stats": (C.c_int, [vp, vp]),
        "rt_bvh4_export": (C.c_int, [vp, vp, vp, vp, vp]),
        "rt_cwbvh_stats": (C.c_int, [vp, vp]),
        "rt_cwbvh_export": (C.c_int, [vp, vp, vp]),
        "rt_stats_new": (vp, []),
        "rt_stats_free": (None, [vp]),
        "rt_stats_stats": (C.c_int, [vp, u32, cp, sz]),
        "rt_stats_mean_stats": (C.c_int, [vp, cp, sz]),
        "rt_benchmark_new": (vp, []),
        "rt_benchmark_free": (None, [vp]),
        "rt_benchmark_start": (C.c_int, [vp, cp]),
        "rt_benchmark_stop": (C.c_int, [vp, cp]),
        "rt_benchmark_report": (C.c_int, [vp, cp, sz]),
        "rt_version": (cp, []),
        "rt_kernels_hash": (cp, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def version() -> str:
    return lib().rt_version().decode()


def kernels_hash() -> str:
    """digest of the kernel sources the loaded library was built from (rt_kernels_hash)"""
    return lib().rt_kernels_hash().decode()


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---- flattened scene ---------------------------------------------------------------------------------------


class Scene:
    """Flattened `Scene` (scene/mod.rs:24-29) as produced by the product's Collada loader."""

    def __init__(self, handle):
        self._h = handle
        d = CSceneDesc()
        lib().rt_scene_get_desc(self._h, C.byref(d))
        self.desc = d
        n = d.num_triangles
        self.vertices = np.ctypeslib.as_array(d.vertices, shape=(n, 9)).copy() if n else np.zeros((0, 9), np.float32)
        self.tri_geom = np.ctypeslib.as_array(d.tri_geom, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        self.materials = [
            (d.materials[g].kind, tuple(np.float32(x) for x in d.materials[g].rgb), d.materials[g].texture_id)
            for g in range(d.num_geometries)
        ]
        self.lights = [
            (np.array(d.lights[i].pos[:], np.float32), np.array(d.lights[i].color[:], np.float32)) for i in range(d.num_lights)
        ]
        self.textures = []
        for k in range(d.num_textures):
            t = d.textures[k]
            self.textures.append((t.width, t.height, np.ctypeslib.as_array(t.rgb, shape=(t.width * t.height * 3,)).copy()))
        self.camera_orientation = np.array(d.camera_orientation[:], np.float32)
        self.camera_fov_deg = np.float32(d.camera_fov_deg)

    def close(self):
        if self._h:
            lib().rt_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_scene(collada_filename: str) -> Scene:
    """ColladaLoader::from_file (scene/loaders/colladaloader.rs:34-45) through rt_scene_load_file."""
    h = C.c_void_p()
    err = C.create_string_buffer(1024)
    rc = lib().rt_scene_load_file(collada_filename.encode(), C.byref(h), err, len(err))
    if rc != 0:
        raise RtError(rc, err.value.decode(errors="replace"))
    return Scene(h)


@dataclass
class Config:
    """rt_config with the reference's defaults (RECURSIONS = 2, SUB_SPREAD = 1, 50 rows per call)."""

    width: int
    height: int
    triangles_per_leaf: int = DEFAULT_TRIANGLES_PER_LEAF
    rows_per_call: int = 50
    recursions: int = 2
    sub_spread: int = 1
    jitter_mode: int = JITTER_HASHED
    seed: int = 0
    accel: int = ACCEL_BVH
    device: int = -1
    shard_index: int = 0
    shard_count: int = 1
    band_rows: int = 8

    def to_c(self) -> CConfig:
        c = CConfig()
        for f, _ in CConfig._fields_:
            setattr(c, f, getattr(self, f))
        return c


# ---- RayTracer ----------------------------------------------------------------------------------------------


class _Camera:
    """`pub camera: Camera` (raytracer/mod.rs:38; scene/camera.rs:63-78)."""

    def __init__(self, rt: "RayTracer"):
        self._rt = rt

    def move_rel(self, x: float, y: float, z: float) -> None:
        self._rt._check(lib().rt_camera_move_rel(self._rt._h, x, y, z))

    def add_x_angle(self, radians: float) -> None:
        self._rt._check(lib().rt_camera_add_x_angle(self._rt._h, radians))

    def add_y_angle(self, radians: float) -> None:
        self._rt._check(lib().rt_camera_add_y_angle(self._rt._h, radians))

    def set_state(self, x_angle: float, y_angle: float, pos) -> None:
        p = (C.c_float * 3)(*[float(v) for v in pos])
        self._rt._check(lib().rt_camera_set_state(self._rt._h, x_angle, y_angle, p))

    def plane_matrix(self):
        """(A, origin) of rt_get_camera_plane_matrix: (X, Y, Z) = A (p - origin) puts a point of the camera ray of pixel (u, v) at
        X / Z = u + xi1, Y / Z = v + xi2, Z = t"""
        out = np.zeros(12, np.float64)
        self._rt._check(lib().rt_get_camera_plane_matrix(self._rt._h, _ptr(out)))
        return out[:9].reshape(3, 3), out[9:]

    def matrices(self) -> np.ndarray:
        """rotation_matrix[16], orientation_matrix[16], max_x, max_y"""
        out = np.zeros(34, np.float32)
        self._rt._check(lib().rt_camera_get(self._rt._h, _ptr(out)))
        return out


class _Film:
    """`pub film: Film` (raytracer/mod.rs:41; film.rs:31-48)."""

    def __init__(self, rt: "RayTracer"):
        self._rt = rt

    def clear(self) -> None:
        self._rt._check(lib().rt_film_clear(self._rt._h))

    def pixel_datas(self) -> np.ndarray:
        """[W*H, 7]: pixel_sum rgb, pixel_sum_squared rgb, num_samples"""
        out = np.zeros((self._rt.width * self._rt.height, 7), np.float32)
        self._rt._check(lib().rt_get_film(self._rt._h, _ptr(out)))
        return out

    def set_pixel_datas(self, film7: np.ndarray) -> None:
        """overwrite the film (`pub pixel_datas`, film.rs:27-29) from the layout pixel_datas() returns"""
        a = np.ascontiguousarray(film7, np.float32)
        assert a.shape == (self._rt.width * self._rt.height, 7)
        self._rt._check(lib().rt_set_film(self._rt._h, _ptr(a)))

    def get_estimated_variances(self) -> np.ndarray:
        """Film::get_estimated_variances (film.rs:50-67): [W*H, 3]"""
        out = np.zeros((self._rt.width * self._rt.height, 3), np.float32)
        self._rt._check(lib().rt_get_estimated_variances(self._rt._h, _ptr(out)))
        return out


class RayTracer:
    """Device-resident `RayTracer` (raytracer/mod.rs:32-128)."""

    def __init__(self, handle, width: int, height: int, num_triangles: int = 0):
        self._h = handle
        self.width, self.height = width, height
        self.num_triangles = num_triangles
        self.camera = _Camera(self)
        self.film = _Film(self)

    # -- construction helpers --
    @staticmethod
    def from_scene(scene, config: Config) -> "RayTracer":
        """build_raytracer (lib.rs:29-44) on a flattened scene: `scene` is a Scene or any object with the same arrays."""
        keep = []
        d = CSceneDesc()
        verts = np.ascontiguousarray(scene.vertices, dtype=np.float32)
        geom = np.ascontiguousarray(scene.tri_geom, dtype=np.uint32)
        keep += [verts, geom]
        d.num_triangles = verts.shape[0]
        d.vertices = verts.ctypes.data_as(C.POINTER(C.c_float))
        d.tri_geom = geom.ctypes.data_as(C.POINTER(C.c_uint32))
        mats = (CMaterial * max(1, len(scene.materials)))()
        for i, (kind, rgb, tex) in enumerate(scene.materials):
            mats[i].kind = int(kind)
            mats[i].rgb[:] = [float(x) for x in rgb]
            mats[i].texture_id = int(tex)
        d.num_geometries = len(scene.materials)
        d.materials = mats
        lights = (CLight * max(1, len(scene.lights)))()
        for i, (pos, col) in enumerate(scene.lights):
            lights[i].pos[:] = [float(x) for x in pos]
            lights[i].color[:] = [float(x) for x in col]
        d.num_lights = len(scene.lights)
        d.lights = lights
        texs = (CTexture * max(1, len(scene.textures)))()
        for i, (w, h, rgb) in enumerate(scene.textures):
            arr = np.ascontiguousarray(rgb, dtype=np.float32)
            keep.append(arr)
            texs[i].width, texs[i].height = int(w), int(h)
            texs[i].rgb = arr.ctypes.data_as(C.POINTER(C.c_float))
        d.num_textures = len(scene.textures)
        d.textures = texs
        d.camera_orientation[:] = [float(x) for x in scene.camera_orientation]
        d.camera_fov_deg = float(scene.camera_fov_deg)
        for name in ("normals", "uvs"):  # reserved attributes: passed through when the scene object has them, ignored by the path
            arr = getattr(scene, name, None)
            if arr is not None:
                arr = np.ascontiguousarray(arr, dtype=np.float32)
                keep.append(arr)
                setattr(d, name, arr.ctypes.data_as(C.POINTER(C.c_float)))
        cfg = config.to_c()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().rt_create(C.byref(d), C.byref(cfg), C.byref(h), err, len(err))
        if rc != 0:
            raise RtError(rc, err.value.decode(errors="replace"))
        return RayTracer(h, config.width, config.height, int(verts.shape[0]))

    def close(self) -> None:
        if self._h:
            lib().rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RtError(rc, lib().rt_last_error(self._h).decode(errors="replace"))

    # -- the reference's two render calls --
    def trace_frame_additive(self) -> int:
        n = C.c_uint32()
        self._check(lib().rt_trace_frame_additive(self._h, C.byref(n)))
        return n.value

    def get_tonemapped_pixels(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.width * self.height, np.uint32)
        self._check(lib().rt_get_tonemapped_pixels(self._h, _ptr(out)))
        return out

    def get_tonemapped_pixels_into(self, host_ptr: int) -> None:
        """Same, into a caller-owned (ideally pinned) host buffer given by address."""
        self._check(lib().rt_get_tonemapped_pixels(self._h, C.c_void_p(host_ptr)))

    def get_tonemapped_pixels_delta_into(self, host_ptr: int) -> None:
        """Incremental readback into a buffer the caller keeps across calls: only the rows that changed since the previous
        call with the same buffer are copied (rt_get_tonemapped_pixels_delta)."""
        self._check(lib().rt_get_tonemapped_pixels_delta(self._h, C.c_void_p(host_ptr)))

    def get_tonemapped_pixels_async(self, pinned_host_ptr: int) -> None:
        """Pipelined readback: snapshot now, device -> host copy on a second stream while the next trace call runs.
        The pixels are valid after wait_pixels()."""
        self._check(lib().rt_get_tonemapped_pixels_async(self._h, C.c_void_p(pinned_host_ptr)))

    def wait_pixels(self, keep: int = 0) -> None:
        """blocks until at most `keep` pipelined readbacks are still in flight"""
        self._check(lib().rt_wait_pixels_keep(self._h, keep))

    # -- extensions used by tests / bench --
    def configure(self, recursions=0, sub_spread=1, jitter_mode=JITTER_FIXED_HALF, seed=0, accel=ACCEL_BVH) -> None:
        self._check(lib().rt_configure(self._h, recursions, sub_spread, jitter_mode, seed, accel))

    def set_rows_per_call(self, rows: int) -> None:
        self._check(lib().rt_set_rows_per_call(self._h, rows))

    def trace_rows(self, first_row: int, n_rows: int, spp: int = 1, want_shadow: bool = True):
        """Returns (n_primary, n_shadow); n_shadow forces a stream synchronisation, pass want_shadow=False to stay async."""
        npri, nsh = C.c_uint64(), C.c_uint64()
        self._check(lib().rt_trace_rows(self._h, first_row, n_rows, spp, C.byref(npri), C.byref(nsh) if want_shadow else None))
        return npri.value, (nsh.value if want_shadow else None)

    def get_primary_ids(self) -> np.ndarray:
        out = np.empty(self.width * self.height, np.uint32)
        self._check(lib().rt_get_primary_ids(self._h, _ptr(out)))
        return out

    def launch_stats(self) -> dict:
        s = CLaunchStats()
        self._check(lib().rt_get_launch_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in CLaunchStats._fields_}

    def kernels_launched(self) -> int:
        return int(lib().rt_kernels_launched(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        self._check(lib().rt_set_stream(self._h, C.c_void_p(cuda_stream)))

    def ldr_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(lib().rt_get_ldr_device_ptr(self._h, C.byref(p)))
        return p.value

    def set_ldr_target(self, dev_ptr: int | None) -> None:
        self._check(lib().rt_set_ldr_target(self._h, C.c_void_p(dev_ptr or 0)))

    def owned_ldr_rows_to(self, dev_ptr: int | None) -> int:
        n = C.c_uint32()
        self._check(lib().rt_get_owned_ldr_rows_device(self._h, C.c_void_p(dev_ptr or 0), C.byref(n)))
        return n.value

    def set_host_frame(self, pinned_host_ptr: int | None) -> None:
        """Zero-copy readback into a page-locked host buffer (rt_set_host_frame)."""
        self._check(lib().rt_set_host_frame(self._h, C.c_void_p(pinned_host_ptr or 0)))

    def host_register(self, host_ptr: int, nbytes: int) -> int:
        """page-locks a host range (e.g. shared memory) for this process and returns its device address"""
        p = C.c_void_p()
        self._check(lib().rt_host_register(self._h, C.c_void_p(host_ptr), nbytes, C.byref(p)))
        return p.value

    def host_unregister(self, host_ptr: int) -> None:
        self._check(lib().rt_host_unregister(self._h, C.c_void_p(host_ptr)))

    def copy_owned_rows(self, src_frame: int, dst_frame: int, cuda_stream: int = 0) -> None:
        """copies the rows this shard owns from one frame to the same rows of another (device or page-locked host)"""
        self._check(lib().rt_copy_owned_rows(self._h, C.c_void_p(src_frame), C.c_void_p(dst_frame), C.c_void_p(cuda_stream or 0)))

    def signal_flag_on_stream(self, dev_flag: int, value: int, cuda_stream: int = 0) -> None:
        self._check(lib().rt_signal_flag_on_stream(self._h, C.c_void_p(dev_flag), value, C.c_void_p(cuda_stream or 0)))

    def set_tuning(self, key: int, value: int) -> None:
        self._check(lib().rt_set_tuning(self._h, key, value))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(lib().rt_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, dev_ptr: int) -> None:
        self._check(lib().rt_device_free(self._h, C.c_void_p(dev_ptr)))

    def ipc_export(self, dev_ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(lib().rt_ipc_export(self._h, C.c_void_p(dev_ptr), buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64)(*handle)
        p = C.c_void_p()
        self._check(lib().rt_ipc_open(self._h, buf, C.byref(p)))
        return p.value

    def ipc_close(self, dev_ptr: int) -> None:
        self._check(lib().rt_ipc_close(self._h, C.c_void_p(dev_ptr)))

    def signal_flag(self, dev_flag: int, value: int) -> None:
        self._check(lib().rt_stream_signal_flag(self._h, dev_flag, value))

    def signal_then_wait(self, dev_signal_flag: int, value: int, dev_wait_flag: int, target: int) -> None:
        self._check(lib().rt_stream_signal_then_wait(self._h, dev_signal_flag, value, dev_wait_flag, target))

    def wait_flags(self, dev_flags: int, n_flags: int, target: int, signal_slot: int = -1, release_slot: int = -1) -> None:
        self._check(lib().rt_stream_wait_flags(self._h, dev_flags, n_flags, target, signal_slot, release_slot))

    def set_done_signal(self, dev_flag: int | None, value: int) -> None:
        """the next trace call publishes `value` at *dev_flag when its last pixel is stored (rt_set_done_signal)"""
        self._check(lib().rt_set_done_signal(self._h, C.c_void_p(dev_flag or 0), value))

    def stream_write_value(self, dev_flag: int, value: int) -> None:
        """stream memory operation: *dev_flag = value after everything enqueued before (no kernel launch)"""
        self._check(lib().rt_stream_write_value(self._h, C.c_void_p(dev_flag), value))

    def stream_wait_value(self, dev_flag: int, value: int) -> None:
        """stream memory operation: hold the stream until *dev_flag >= value (cyclic); no timeout"""
        self._check(lib().rt_stream_wait_value(self._h, C.c_void_p(dev_flag), value))

    def sync_timeouts(self) -> int:
        n = C.c_uint32()
        self._check(lib().rt_sync_timeouts(self._h, C.byref(n)))
        return int(n.value)

    def counters_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(lib().rt_get_counters_device_ptr(self._h, C.byref(p)))
        return p.value

    def octree_stats(self) -> dict:
        raw = np.zeros(6, np.uint64)
        self._check(lib().rt_octree_stats(self._h, _ptr(raw)))
        return dict(zip(["nodes", "inner", "leaves", "empty_leaves", "tri_refs", "depth"], (int(x) for x in raw)))

    def octree_export(self):
        st = self.octree_stats()
        n, refs = st["nodes"], st["tri_refs"]
        cubes = np.zeros((n, 6), np.float32)
        first_child = np.zeros(n, np.int32)
        leaf_offset = np.zeros(n + 1, np.uint32)
        leaf_tris = np.zeros(max(refs, 1), np.uint32)
        nref = C.c_uint64()
        self._check(lib().rt_octree_export(self._h, _ptr(cubes), _ptr(first_child), _ptr(leaf_offset), _ptr(leaf_tris), C.byref(nref)))
        return cubes, first_child, leaf_offset, leaf_tris[:refs]

    def bvh_stats(self) -> dict:
        raw = np.zeros(4, np.uint64)
        self._check(lib().rt_bvh_stats(self._h, _ptr(raw)))
        return dict(zip(["nodes", "leaves", "max_leaf", "depth"], (int(x) for x in raw)))

    def bvh_export(self):
        """(boxes [N,2,2,3] lo/hi per child, children [N,2], counts [N,2], tri_order [T])"""
        n = self.bvh_stats()["nodes"]
        boxes = np.zeros((n, 2, 2, 3), np.float32)
        children = np.zeros((n, 2), np.int32)
        counts = np.zeros((n, 2), np.int32)
        order = np.zeros(max(1, self.num_triangles), np.uint32)
        self._check(lib().rt_bvh_export(self._h, _ptr(boxes), _ptr(children), _ptr(counts), _ptr(order)))
        return boxes, children, counts, order[: self.num_triangles]

    def ray_totals(self) -> dict:
        """rays issued since creation (exact, over any number of asynchronous trace calls); synchronises the stream"""
        raw = np.zeros(3, np.uint64)
        self._check(lib().rt_get_ray_totals(self._h, _ptr(raw)))
        return dict(zip(["primary", "shadow", "bounce"], (int(x) for x in raw)))

    def tile_costs(self):
        """(cycles per 32-lane tile of the last scheduled launch geometry, items in its sorted queue)"""
        n, items = C.c_uint32(), C.c_uint32()
        self._check(lib().rt_get_tile_costs(self._h, None, 0, C.byref(n), C.byref(items)))
        out = np.zeros(max(1, n.value), np.uint32)
        self._check(lib().rt_get_tile_costs(self._h, _ptr(out), n.value, C.byref(n), C.byref(items)))
        return out[: n.value], int(items.value)

    def lbvh_build(self) -> dict:
        """(re)builds the BVH on the GPU; returns nodes, depth, triangles and the device time of the build (ms)"""
        raw = np.zeros(3, np.uint64)
        ms = C.c_float()
        self._check(lib().rt_lbvh_build(self._h, _ptr(raw), C.byref(ms)))
        return {"nodes": int(raw[0]), "depth": int(raw[1]), "triangles": int(raw[2]), "build_ms": float(ms.value)}

    def lbvh_export(self):
        """GPU-built tree in the layout of bvh_export"""
        n = max(self.num_triangles - 1, 1)
        boxes = np.zeros((n, 2, 2, 3), np.float32)
        children = np.zeros((n, 2), np.int32)
        counts = np.zeros((n, 2), np.int32)
        order = np.zeros(max(1, self.num_triangles), np.uint32)
        self._check(lib().rt_lbvh_export(self._h, _ptr(boxes), _ptr(children), _ptr(counts), _ptr(order)))
        return boxes, children, counts, order[: self.num_triangles]

    def bvh4_stats(self) -> dict:
        raw = np.zeros(4, np.uint64)
        self._check(lib().rt_bvh4_stats(self._h, _ptr(raw)))
        return dict(zip(["nodes", "leaves", "max_leaf", "depth"], (int(x) for x in raw)))

    def bvh4_export(self):
        """(boxes [N,4,2,3] lo/hi per child, children [N,4], counts [N,4], tri_order [T])"""
        n = self.bvh4_stats()["nodes"]
        boxes = np.zeros((n, 4, 2, 3), np.float32)
        children = np.zeros((n, 4), np.int32)
        counts = np.zeros((n, 4), np.int32)
        order = np.zeros(max(1, self.num_triangles), np.uint32)
        self._check(lib().rt_bvh4_export(self._h, _ptr(boxes), _ptr(children), _ptr(counts), _ptr(order)))
        return boxes, children, counts, order[: self.num_triangles]

    def cwbvh_stats(self) -> dict:
        raw = np.zeros(4, np.uint64)
        self._check(lib().rt_cwbvh_stats(self._h, _ptr(raw)))
        return dict(zip(["nodes", "leaf_children", "tri_slots", "depth"], (int(x) for x in raw)))

    def cwbvh_export(self):
        """(node words [N,5,4] u32 — layout in csrc/cwbvh_build.cpp —, tri_order [T])"""
        st = self.cwbvh_stats()
        words = np.zeros((st["nodes"], 5, 4), np.uint32)
        order = np.zeros(max(1, st["tri_slots"]), np.uint32)
        self._check(lib().rt_cwbvh_export(self._h, _ptr(words), _ptr(order)))
        return words, order[: st["tri_slots"]]


def _create(fn_name: str, arg: str, triangles_per_leaf: int, width: int, height: int) -> RayTracer:
    h = C.c_void_p()
    err = C.create_string_buffer(1024)
    rc = getattr(lib(), fn_name)(arg.encode(), triangles_per_leaf, width, height, C.byref(h), err, len(err))
    if rc != 0:
        raise RtError(rc, err.value.decode(errors="replace"))
    return RayTracer(h, width, height)


def create_raytracer(collada_doc: str, triangles_per_leaf: int, width: int, height: int) -> RayTracer:
    """lib.rs:15-20"""
    return _create("rt_create_raytracer", collada_doc, triangles_per_leaf, width, height)


def create_raytracer_from_file(collada_filename: str, triangles_per_leaf: int, width: int, height: int) -> RayTracer:
    """lib.rs:22-27"""
    return _create("rt_create_raytracer_from_file", collada_filename, triangles_per_leaf, width, height)


# ---- stats.rs / timing crate ---------------------------------------------------------------------------------


class Stats:
    """raytracer_lib/src/stats.rs"""

    def __init__(self):
        self._h = lib().rt_stats_new()

    def stats(self, num_primary_rays: int) -> str:
        buf = C.create_string_buffer(256)
        lib().rt_stats_stats(self._h, num_primary_rays, buf, len(buf))
        return buf.value.decode()

    def mean_stats(self) -> str:
        buf = C.create_string_buffer(256)
        lib().rt_stats_mean_stats(self._h, buf, len(buf))
        return buf.value.decode()

    def __del__(self):
        try:
            lib().rt_stats_free(self._h)
        except Exception:
            pass


class BenchMark:
    """timing/src/lib.rs `BenchMark` (start/stop/time_scope + Display)."""

    def __init__(self):
        self._h = lib().rt_benchmark_new()

    def start(self, name: str) -> None:
        lib().rt_benchmark_start(self._h, name.encode())

    def stop(self, name: str) -> None:
        if lib().rt_benchmark_stop(self._h, name.encode()) != 0:
            raise KeyError("unexpected name in stop()")

    def time_scope(self, name: str):
        bm = self

        class _Scope:
            def __enter__(self_inner):
                bm.start(name)
                return self_inner

            def __exit__(self_inner, *exc):
                bm.stop(name)
                return False

        return _Scope()

    def __str__(self) -> str:
        buf = C.create_string_buffer(16384)
        lib().rt_benchmark_report(self._h, buf, len(buf))
        return buf.value.decode()

    def __del__(self):
        try:
            lib().rt_benchmark_free(self._h)
        except Exception:
            pass
