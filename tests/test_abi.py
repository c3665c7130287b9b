"""The C-ABI library loads and exports every symbol include/rt_b200.h declares (no compute calls: CPU only)."""
import os
import re
import subprocess

import raytracer_rs_b200 as rt
from raytracer_rs_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_binding_table():
    assert header_functions() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", rt.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rt_[a-z0-9_]+)", out))
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing


def test_library_loads_and_identifies_itself():
    L = rt.lib()  # binds argtypes for every symbol; AttributeError if one is missing
    assert all(hasattr(L, s) for s in api.ABI_SYMBOLS)
    assert rt.version().startswith("rt_b200") and "sm_100a" in rt.version()
    assert L.rt_launch_param_bytes() > 100


def test_only_sm100a_code_is_shipped():
    out = subprocess.run(["cuobjdump", "-lelf", rt.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_config_default_mirrors_reference_constants():
    c = api.CConfig()
    rt.lib().rt_config_default(c, 1024, 768)
    assert (c.width, c.height) == (1024, 768)
    assert c.triangles_per_leaf == rt.DEFAULT_TRIANGLES_PER_LEAF == 70  # oct_tree_intersector.rs:12
    assert c.rows_per_call == 50  # mod.rs:87
    assert c.recursions == 2 and c.sub_spread == 1  # mod.rs:81-82
    assert c.shard_count == 1 and c.device == -1


def test_product_does_not_link_or_import_the_oracle():
    """The oracle is test infrastructure: nothing under raytracer_rs_b200/ may reference it."""
    out = subprocess.run(["ldd", rt.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "raytracer_rs_b200")):
        if "build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "orc_" not in text, (dirpath, f)


def build_c_client(tmp_path):
    """Compiles tests/c_client/host_client.c as strict C99 against include/rt_b200.h and links librt_b200.so."""
    exe = str(tmp_path / "host_client")
    libdir = os.path.dirname(rt.lib_path())
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_client", "host_client.c"), "-o", exe, "-L", libdir, "-l:" + os.path.basename(rt.lib_path()),
                    "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    return exe


def run_c_client(exe, *args):
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return dict(line.split(" ", 1) for line in out.stdout.splitlines() if " " in line)


def test_plain_c_client_drives_the_host_side_of_the_abi(tmp_path):
    """The boundary is a C ABI for a non-C++ host (the reference is Rust): the header compiles as strict C99 and a C
    program gets the same scene, tree and camera as the Python binding; a host-only handle refuses to render."""
    import numpy as np

    got = run_c_client(build_c_client(tmp_path), os.path.join(ROOT, "data", "thai2.dae"))
    assert got["version"] == rt.version()
    scene = rt.load_scene(os.path.join(ROOT, "data", "thai2.dae"))
    assert int(got["triangles"]) == scene.vertices.shape[0] == 20049 and int(got["lights"]) == len(scene.lights)
    assert np.float32(got["fov"]) == scene.camera_fov_deg
    assert got["reserved_arrays"] == "null"  # rt_scene_desc.normals / uvs: a slot the loader leaves empty, as the reference does
    t = rt.RayTracer.from_scene(scene, rt.Config(96, 54, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_OCTREE,
                                                 device=api.DEVICE_NONE))
    st = t.octree_stats()
    assert (int(got["octree_nodes"]), int(got["octree_refs"]), int(got["octree_depth"])) == (st["nodes"], st["tri_refs"], st["depth"])
    t.camera.move_rel(0.25, 0.0, -0.5)
    t.camera.add_x_angle(0.125)
    t.camera.add_y_angle(-0.0625)
    cam = t.camera.matrices()
    assert np.float32(got["camera_rot0"]) == cam[0] and np.float32(got["camera_max_x"]) == cam[32]
    assert [np.float32(x) for x in got["camera_pos"].split()] == [cam[28], cam[29], cam[30]]
    assert int(got["trace_rc"]) == -3 and "RT_DEVICE_NONE" in got["trace_error"]  # no CPU fallback
    assert got["stats_prefix"].startswith("fps:") and got["mean_stats_prefix"].startswith("mean fps")
    assert int(got["benchmark_unknown_rc"]) == -1 and got["benchmark_report"].startswith("frame total:")
    t.close()
