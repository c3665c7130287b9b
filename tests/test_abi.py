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
