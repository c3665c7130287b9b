"""The plain-C client (tests/c_client/host_client.c) rendering on the GPU: what a non-Python host gets through the C ABI
equals what the Python binding gets. (Named to run after the other GPU tests.)"""
import os

import numpy as np
import pytest

import raytracer_rs_b200 as rt
from test_abi import build_c_client, run_c_client

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_renders_the_same_frame_as_the_python_binding(tmp_path):
    dae = os.path.join(ROOT, "data", "ico3_tex.dae")
    got = run_c_client(build_c_client(tmp_path), dae, "0")
    t = rt.RayTracer.from_scene(rt.load_scene(dae), rt.Config(96, 54, recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_OCTREE, device=0))
    # the client applies these camera moves before it renders
    t.camera.move_rel(0.25, 0.0, -0.5)
    t.camera.add_x_angle(0.125)
    t.camera.add_y_angle(-0.0625)
    n = t.trace_frame_additive()  # 50 rows of 96 pixels
    frame = t.get_tonemapped_pixels()
    assert int(got["trace_rc"]) == 0 and int(got["primary_rays"]) == n == 50 * 96
    assert int(got["frame_checksum"]) == int(frame.astype(np.uint64).sum())
    assert (frame[: 50 * 96] != 0xFFFFFFFF).all() and (frame[50 * 96:] == 0xFFFFFFFF).all()  # rows 50..53 were never sampled
    # three more 50-row calls read back incrementally into one buffer, then Film::get_estimated_variances
    assert int(got["delta_matches_full"]) == 1
    for _ in range(3):
        t.trace_frame_additive()
    var = t.film.get_estimated_variances()
    assert int(got["variance_finite_values"]) == int(np.isfinite(var).sum()) > 0
    t.close()
