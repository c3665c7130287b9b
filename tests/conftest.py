"""pytest configuration: `gpu` marker, shared fixtures.

CPU suite  : python -m pytest tests -x -q -m "not gpu"   (oracle vs golden vectors, host logic, ABI, gloo sharding)
GPU suite  : python -m pytest tests -x -q -m gpu         (CUDA path vs oracle through the C ABI, on a B200)
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = os.path.join(ROOT, "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# the BASELINE.json configurations: (scene, width, height)
CONFIGS = {
    "ico2": ("ico2.dae", 1024, 768),
    "4boxes": ("4boxes.dae", 1920, 1080),
    "ico3_tex": ("ico3_tex.dae", 1920, 1080),
    "thai2": ("thai2.dae", 1920, 1080),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The native libraries must exist; build them if the checkout is fresh (nvcc cross-compiles without a GPU)."""
    import raytracer_rs_b200 as rt

    if not os.path.exists(rt.lib_path()):
        import __graft_entry__ as g

        g.build()
    from oracle_lib import build_oracle

    build_oracle()
    return True


@pytest.fixture(scope="session")
def scenes():
    """Flattened scenes, loaded once through the product loader (rt_scene_load_file)."""
    import raytracer_rs_b200 as rt

    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = rt.load_scene(os.path.join(DATA, CONFIGS[name][0]))
        return cache[name]

    return get


@pytest.fixture(scope="session")
def ref_scenes():
    """The same scenes through the INDEPENDENT numpy Collada reader (tests/collada_ref.py, written from the Rust loader's sources): what
    the oracle is fed in the full-resolution parity tests, so that those tests do not share a loader with the product they check."""
    from collada_ref import load_collada

    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_collada(os.path.join(DATA, CONFIGS[name][0]))
        return cache[name]

    return get


def channel_diff(a, b):
    """max per-channel |difference| between two packed 0xAARRGGBB frames"""
    import numpy as np

    return max(int(np.abs(((a >> k) & 255).astype(np.int32) - ((b >> k) & 255).astype(np.int32)).max()) for k in (0, 8, 16, 24))
