"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on identical flattened inputs.

Bars (BASELINE.json north star): primary-hit primitive ids equal on >= 99.99 % of pixels, 8-bit colour within 1 LSB
per channel. The exact-semantics octree kernel is additionally required to match ids on EVERY pixel. HDR film sums
are compared bit for bit except where the specular term differs by the documented <= 1 ulp of powf.
"""
import json
import os

import numpy as np
import pytest

import raytracer_rs_b200 as rt
from conftest import CONFIGS, GOLDEN, channel_diff
from oracle_lib import ISECT_BRUTE, JITTER_FIXED, JITTER_HASHED, Oracle

pytestmark = pytest.mark.gpu

ACCELS = [(rt.ACCEL_OCTREE, "octree"), (rt.ACCEL_BVH, "bvh"), (rt.ACCEL_CWBVH, "cwbvh"), (rt.ACCEL_BVH4, "bvh4"), (rt.ACCEL_LBVH, "lbvh")]
ID_BAR = 0.9999  # north star
LSB_BAR = 1


def gpu_tracer(scene, w, h, accel, jitter=rt.JITTER_FIXED_HALF, seed=0, **kw):
    return rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=jitter, seed=seed, accel=accel, **kw))


def check_frame(tracer, oracle, exact_ids):
    ids, ldr = tracer.get_primary_ids(), tracer.get_tonemapped_pixels()
    ids_o, ldr_o = oracle.get_primary_ids(), oracle.get_tonemapped_pixels()
    agree = float((ids == ids_o).mean())
    if exact_ids:
        assert agree == 1.0, f"ids differ on {(ids != ids_o).sum()} pixels"
    assert agree >= ID_BAR, agree
    same = ids == ids_o
    lsb = channel_diff(ldr[same], ldr_o[same])
    assert lsb <= LSB_BAR, lsb
    assert float((ldr == ldr_o).mean()) >= ID_BAR
    return agree, lsb


@pytest.fixture(scope="module")
def oracle_frames(ref_scenes):
    """Full-resolution pinned-mode oracle frames of the four BASELINE configurations (computed once, all host threads). The oracle's
    scene comes from the independent loader (tests/collada_ref.py), the GPU's from the product's: no common-mode loader."""
    cache = {}

    def get(name):
        if name not in cache:
            _, w, h = CONFIGS[name]
            o = Oracle(ref_scenes(name), w, h)
            o.configure(recursions=0, jitter=JITTER_FIXED)
            o.trace_rows(0, h, 1, threads=0)
            cache[name] = o
        return cache[name]

    return get


@pytest.mark.parametrize("accel,aname", ACCELS)
@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_resolution_parity(scenes, oracle_frames, name, accel, aname):
    """BASELINE.json configs[0..3] at their full sizes, pinned mode (xi = 0.5, recursions = 0, 1 spp)."""
    _, w, h = CONFIGS[name]
    o = oracle_frames(name)
    t = gpu_tracer(scenes(name), w, h, accel)
    n_primary, n_shadow = t.trace_rows(0, h, 1)
    c = o.counters()
    assert n_primary == w * h == c["rays"]["primary"]
    # a shadow ray is issued per (hit, light) with n.l >= 0 (mod.rs:218-226): identical hits -> identical count
    assert n_shadow == c["rays"]["shadow"]
    check_frame(t, o, exact_ids=True)  # on these four scenes the BVH path also agrees on every pixel
    film, film_o = t.film.pixel_datas(), o.get_film()
    assert np.array_equal(film[:, 6], film_o[:, 6])
    bits_differ = (film.view(np.uint32) != film_o.view(np.uint32)).any(axis=1)
    assert bits_differ.mean() < 1e-3  # only the <= 1 ulp powf cases
    # <= 1 ulp in the specular term -> a few ulps in sum / sum of squares; atol absorbs the subnormal range, where
    # one ulp is a large relative step (x^32 underflows for small |V.R|)
    assert np.allclose(film[bits_differ], film_o[bits_differ], rtol=1e-6, atol=1e-30)
    t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
@pytest.mark.parametrize("name", list(CONFIGS))
def test_against_committed_golden(scenes, name, accel, aname):
    """Committed oracle vectors (tests/golden): no oracle needed at run time."""
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))[name]
    g = np.load(os.path.join(GOLDEN, f"{name}_{man['width']}x{man['height']}.npz"))
    w, h = man["width"], man["height"]
    t = gpu_tracer(scenes(name), w, h, accel)
    _, n_shadow = t.trace_rows(0, h, 1)
    assert n_shadow == man["shadow_rays"]
    assert np.array_equal(t.get_primary_ids(), g["ids"])
    assert channel_diff(t.get_tonemapped_pixels(), g["ldr"]) <= LSB_BAR
    t.film.clear()
    t.configure(recursions=0, jitter_mode=rt.JITTER_HASHED, seed=man["seed"], accel=accel)
    t.trace_rows(0, h, 2)
    assert channel_diff(t.get_tonemapped_pixels(), g["ldr_jitter2"]) <= LSB_BAR
    t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_hashed_jitter_multi_sample(scenes, accel, aname):
    """configs[4] in miniature: thai2, hashed jitter, several samples per pixel accumulated in the film."""
    w, h, spp, seed = 480, 270, 4, 7
    s = scenes("thai2")
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_HASHED, seed=seed)
    o.trace_rows(0, h, spp, threads=0)
    t = gpu_tracer(s, w, h, accel, jitter=rt.JITTER_HASHED, seed=seed)
    n_primary, n_shadow = t.trace_rows(0, h, spp)
    assert n_primary == w * h * spp and n_shadow == o.counters()["rays"]["shadow"]
    check_frame(t, o, exact_ids=(accel == rt.ACCEL_OCTREE))
    assert np.array_equal(t.film.pixel_datas()[:, 6], np.full(w * h, spp, np.float32))
    t.close()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_gpu_built_bvh_invariants(scenes, name):
    """f-3: the tree built on the device (lbvh_build.cu: Morton codes, radix sort, Karras hierarchy, refit). Every
    triangle sits in exactly one reachable leaf of at most 4 triangles, every stored child box strictly contains its
    subtree, and the build is deterministic."""
    s = scenes(name)
    n_tri = s.vertices.shape[0]
    t = rt.RayTracer.from_scene(s, rt.Config(64, 64, recursions=0, accel=rt.ACCEL_LBVH))
    st = t.lbvh_build()
    assert st["triangles"] == n_tri and st["nodes"] == n_tri - 1 and 0 < st["depth"] <= 46 and st["build_ms"] > 0
    boxes, children, counts, order = t.lbvh_export()
    st2 = t.lbvh_build()
    again = t.lbvh_export()
    assert st2["depth"] == st["depth"]
    for a, b in zip((boxes, children, counts, order), again):
        assert np.array_equal(a, b)
    assert sorted(order.tolist()) == list(range(n_tri))
    verts = s.vertices.reshape(n_tri, 3, 3)
    seen = np.zeros(n_tri, bool)
    depth_seen = [0]

    def subtree_box(node, level):
        depth_seen[0] = max(depth_seen[0], level)
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        for k in range(2):
            c, cnt = int(children[node, k]), int(counts[node, k])
            if c >= 0:
                lo, hi = subtree_box(c, level + 1)
            else:
                assert 1 <= cnt <= 4
                tri = order[~c:~c + cnt]
                assert not seen[tri].any()
                seen[tri] = True
                lo, hi = verts[tri].reshape(-1, 3).min(0), verts[tri].reshape(-1, 3).max(0)
            assert (boxes[node, k, 0] < lo).all() and (boxes[node, k, 1] > hi).all()
            lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
        return lo_all, hi_all

    import sys
    sys.setrecursionlimit(10000)
    subtree_box(0, 1)
    assert seen.all() and depth_seen[0] == st["depth"]
    t.close()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_kernel_variants_are_bit_identical(scenes, name):
    """The three schedules of the trace kernel (one thread per pixel / persistent tile warps / ray pool with
    shared-memory ray rings) evaluate the same f32 expressions per ray: ids, packed frame, HDR film and ray counts
    must agree bit for bit, with hashed jitter and two samples per pixel."""
    _, w, h = CONFIGS[name]
    ref = None
    for variant in (1, 0, 2):
        t = gpu_tracer(scenes(name), w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=5)
        t.set_tuning(0, variant)
        n_primary, n_shadow = t.trace_rows(0, h, 2)
        got = (t.get_primary_ids(), t.get_tonemapped_pixels(), t.film.pixel_datas().view(np.uint32), n_shadow)
        if ref is None:
            ref = got
        else:
            assert got[3] == ref[3]
            for a, b in zip(got[:3], ref[:3]):
                assert np.array_equal(a, b), variant
        t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_multi_sample_launch_equals_one_launch_per_sample(scenes, accel, aname):
    """rt_trace_rows with spp > 1 has three forms (RT_TUNE_MULTI_SAMPLE_LAUNCH): sample lanes (even spp: the lanes of a warp
    item hold 8, 4 or 2 samples of a pixel and add them to the film in order), sample planes + ordered accumulation pass
    (one launch), one launch per sample. Film, frame, ids and ray counts must be bit-identical in all three — also for a
    shard of interleaved bands, a wrapped row range, and when the film already holds samples."""
    s = scenes("thai2")
    w, h = 640, 360
    cases = [({}, 0, h, 1, (5, 8, 6)), ({}, 300, 101, 1, (4, 16)), ({"shard_index": 1, "shard_count": 3, "band_rows": 8}, 0, h, 1, (2, 8)),
             ({"shard_index": 0, "shard_count": 7, "band_rows": 3}, 5, 333, 1, (5, 12)), ({}, 0, h, 0, (5, 4))]
    for kw, first, rows, variant, spps in cases:
        tracers = []
        for multi in (1, 2, 0):
            t = gpu_tracer(s, w, h, accel, jitter=rt.JITTER_HASHED, seed=9, **kw)
            t.set_tuning(0, variant)
            t.set_tuning(5, multi)
            tracers.append(t)
        for spp in spps:
            got = []
            for t in tracers:
                t.film.clear()
                t.trace_rows(first, rows, 1)  # the film already holds one sample of these rows
                n_primary, n_shadow = t.trace_rows(first, rows, spp)
                got.append((n_primary, n_shadow, t.get_primary_ids(), t.get_tonemapped_pixels(), t.film.pixel_datas().view(np.uint32),
                            t.launch_stats()["kernels_launched"]))
            lanes, planes, single = got
            for other in (planes, single):
                assert lanes[0] == other[0] and lanes[1] == other[1], (kw, spp)
                for x, y in zip(lanes[2:5], other[2:5]):
                    assert np.array_equal(x, y), (kw, spp)
            if spp > 2:
                assert planes[5] < single[5]  # one trace launch (+ accumulation) instead of spp
            if variant == 1 and spp % 2 == 0 and spp >= 4:  # sample lanes: spp / 8, spp / 4 or spp / 2 launches
                assert lanes[5] < single[5], (kw, spp)
        for t in tracers:
            t.close()


def test_4k_16spp_properties(scenes):
    """configs[4] at full size (3840x2160, 16 jittered spp = 132.7 M primary rays): size-independent properties
    instead of a full oracle frame — sample counts, ray-count bounds, determinism (two runs give identical frames),
    octree kernel == BVH kernel, and an exact oracle comparison on a band of rows."""
    w, h, spp, seed = 3840, 2160, 16, 0
    s = scenes("thai2")
    frames = []
    for accel, _ in ACCELS:
        t = gpu_tracer(s, w, h, accel, jitter=rt.JITTER_HASHED, seed=seed)
        n_primary, n_shadow = t.trace_rows(0, h, spp)
        assert n_primary == w * h * spp
        assert 0.2 * n_primary < n_shadow < 0.3 * n_primary  # ~0.25 shadow rays per primary ray on thai2
        frames.append((t.get_tonemapped_pixels(), t.get_primary_ids(), n_shadow))
        if accel == rt.ACCEL_BVH:
            t.film.clear()
            t.trace_rows(0, h, spp, want_shadow=False)
            assert np.array_equal(t.get_tonemapped_pixels(), frames[-1][0])  # deterministic
        t.close()
    for other in frames[1:]:
        assert frames[0][2] == other[2]
        assert float((frames[0][1] == other[1]).mean()) >= ID_BAR
        assert float((frames[0][0] == other[0]).mean()) >= ID_BAR
    # exact check of 24 rows around the statue's centre against the oracle
    r0, nr = 600, 24
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_HASHED, seed=seed)
    o.trace_rows(r0, nr, spp, threads=0)
    band = slice(r0 * w, (r0 + nr) * w)
    assert channel_diff(frames[1][0][band], o.get_tonemapped_pixels()[band]) <= LSB_BAR


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_trace_frame_additive_call_pattern(scenes, accel, aname):
    """The reference's loop (raytracer/src/main.rs:200-201): 50-row bands wrapping modulo height, full-frame
    get_tonemapped_pixels after every band, never-sampled pixels white (Q15), progressive accumulation."""
    w, h = 320, 120
    s = scenes("ico2")
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_HASHED, seed=3)
    t = gpu_tracer(s, w, h, accel, jitter=rt.JITTER_HASHED, seed=3)
    assert (t.get_tonemapped_pixels() == 0xFFFFFFFF).all()
    for call in range(6):  # 300 rows = 2.5 laps
        assert t.trace_frame_additive() == 50 * w == o.trace_frame_additive(threads=0)
        ldr, ldr_o = t.get_tonemapped_pixels(), o.get_tonemapped_pixels()
        assert np.array_equal(ldr == 0xFFFFFFFF, ldr_o == 0xFFFFFFFF), call
        assert channel_diff(ldr, ldr_o) <= LSB_BAR, call
    assert np.array_equal(t.film.pixel_datas()[:, 6], o.get_film()[:, 6])
    t.close()


def test_pipelined_readback_delivers_every_frame(scenes):
    """rt_get_tonemapped_pixels_async / rt_wait_pixels: the copy of frame k overlaps the trace of frame k+1, and every
    frame arrives intact (compared with the blocking readback of the same film state)."""
    import torch

    s = scenes("ico2")
    w, h = 512, 384
    t = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=2)
    ref = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=2)
    bufs = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(2)]
    expect = []
    for k in range(6):
        t.trace_rows(0, h, 1, want_shadow=False)
        t.wait_pixels()  # frame k-1 is complete now
        if k > 0:
            assert np.array_equal(bufs[(k - 1) & 1].numpy().view(np.uint32), expect[k - 1]), k
        t.get_tonemapped_pixels_async(bufs[k & 1].data_ptr())
        ref.trace_rows(0, h, 1, want_shadow=False)
        expect.append(ref.get_tonemapped_pixels())
    t.wait_pixels()
    assert np.array_equal(bufs[5 & 1].numpy().view(np.uint32), expect[5])
    assert not np.array_equal(expect[0], expect[5])  # progressive accumulation changed the frame in between
    t.close()
    ref.close()


def test_pipelined_readback_two_frames_in_flight(scenes):
    """rt_wait_pixels_keep(1): frame k is handed to the copy stream BEFORE the host waits for frame k-1 (two snapshots, two host
    buffers), and every frame still arrives intact; a third call in a row (no wait in between) must not overwrite a snapshot that
    is still being copied."""
    import torch

    s = scenes("ico3_tex")
    w, h = 640, 360
    t = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=5)
    ref = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=5)
    bufs = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(3)]
    expect = []
    for k in range(7):
        t.trace_rows(0, h, 1, want_shadow=False)
        t.get_tonemapped_pixels_async(bufs[k % 3].data_ptr())
        if k != 3:  # frame 3: three copies queued back to back before the host looks at any of them
            t.wait_pixels(1)
            if k > 0:
                assert np.array_equal(bufs[(k - 1) % 3].numpy().view(np.uint32), expect[k - 1]), k
        ref.trace_rows(0, h, 1, want_shadow=False)
        expect.append(ref.get_tonemapped_pixels())
    t.wait_pixels()
    for k in (4, 5, 6):
        assert np.array_equal(bufs[k % 3].numpy().view(np.uint32), expect[k]), k
    t.close()
    ref.close()


def _subscene(scene, tri_index, lights=None):
    """A scene made of some triangles of `scene` (same camera, materials and textures; optionally other lights)."""
    from types import SimpleNamespace

    idx = np.asarray(tri_index, dtype=np.int64)
    return SimpleNamespace(vertices=scene.vertices[idx].copy(), tri_geom=scene.tri_geom[idx].copy(), materials=scene.materials,
                           lights=scene.lights if lights is None else lights, textures=scene.textures,
                           camera_orientation=scene.camera_orientation, camera_fov_deg=scene.camera_fov_deg)


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_degenerate_scenes_and_ragged_image_sizes(scenes, accel, aname):
    """Edge cases the tree builders and the tile schedule must survive: no triangle at all, one, two, a handful; image
    sizes that are not multiples of the 8x4 tile; every structure against the oracle (ids exact except where a BVH may
    resolve an exact tie differently, colour within 1 LSB)."""
    full = scenes("ico2")
    n = full.vertices.shape[0]
    picks = [[], [n - 1], [n - 1, n - 2], list(range(n - 12, n)), list(range(0, n, 7))]
    for tri in picks:
        sub = _subscene(full, tri)
        for w, h in [(101, 67), (64, 3)]:
            o = Oracle(sub, w, h)
            o.configure(recursions=0, jitter=JITTER_FIXED)
            o.trace_rows(0, h, 1, threads=0)
            t = gpu_tracer(sub, w, h, accel)
            n_primary, n_shadow = t.trace_rows(0, h, 1)
            assert n_primary == w * h and n_shadow == o.counters()["rays"]["shadow"], (len(tri), w, h)
            check_frame(t, o, exact_ids=True)
            t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_several_lights_and_no_light(scenes, accel, aname):
    """shade (mod.rs:214) loops over every light: two and three lights accumulate in light order; without a light every
    hit is black. (The ray-pool kernel only takes single-light scenes; these run the tile kernel.)"""
    full = scenes("ico2")
    l0 = full.lights[0]
    extra = [(np.array([-6.0, 9.0, 4.0], np.float32), np.array([3.0, 2.0, 1.0], np.float32)),
             (np.array([2.0, 12.0, -7.0], np.float32), np.array([0.5, 4.0, 2.5], np.float32))]
    w, h = 320, 240
    for lights in ([l0, extra[0]], [l0] + extra, []):
        sub = _subscene(full, range(full.vertices.shape[0]), lights=lights)
        o = Oracle(sub, w, h)
        o.configure(recursions=0, jitter=JITTER_FIXED)
        o.trace_rows(0, h, 1, threads=0)
        for variant in (1, 2):
            t = gpu_tracer(sub, w, h, accel)
            t.set_tuning(0, variant)
            _, n_shadow = t.trace_rows(0, h, 1)
            assert n_shadow == o.counters()["rays"]["shadow"]
            check_frame(t, o, exact_ids=True)
            t.close()


def test_height_smaller_than_band(scenes):
    """height < 50: one call visits rows more than once, sequentially (mod.rs:87-114)."""
    w, h = 64, 20
    s = scenes("4boxes")
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_HASHED, seed=1)
    t = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=1)
    t.trace_frame_additive(), o.trace_frame_additive()
    film, film_o = t.film.pixel_datas(), o.get_film()
    assert np.array_equal(film[:, 6], film_o[:, 6]) and set(film[:, 6]) == {2.0, 3.0}
    assert channel_diff(t.get_tonemapped_pixels(), o.get_tonemapped_pixels()) <= LSB_BAR
    t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
def test_camera_moves_and_film_clear(scenes, accel, aname):
    """The key handlers of raytracer/src/main.rs:124-162: camera.move_rel / add_*_angle then film.clear()."""
    w, h = 384, 216
    s = scenes("thai2")
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_FIXED)
    t = gpu_tracer(s, w, h, accel)
    for fn, args in [("move_rel", (0.1, 0.0, 0.0)), ("add_y_angle", (0.05,)), ("add_x_angle", (-0.03,)), ("move_rel", (0.0, 0.3, -0.5))]:
        getattr(t.camera, fn)(*args)
        getattr(o, fn)(*args)
        t.film.clear()
        o.film_clear()
        t.trace_rows(0, h, 1)
        o.trace_rows(0, h, 1, threads=0)
        check_frame(t, o, exact_ids=(accel == rt.ACCEL_OCTREE))
    t.close()


@pytest.mark.parametrize("tpl", [5, 20, 100])
def test_triangles_per_leaf_parameter(scenes, tpl):
    """create_raytracer's triangles_per_leaf (lib.rs:15, collect.ps1 sweeps 5..100): the exact octree kernel follows
    whatever tree the parameter produces."""
    w, h = 320, 180
    s = scenes("ico2")
    o = Oracle(s, w, h, tpl)
    o.configure(recursions=0, jitter=JITTER_FIXED)
    o.trace_rows(0, h, 1, threads=0)
    t = gpu_tracer(s, w, h, rt.ACCEL_OCTREE, triangles_per_leaf=tpl)
    t.trace_rows(0, h, 1)
    assert t.octree_stats() == o.octree_stats()
    check_frame(t, o, exact_ids=True)
    t.close()


def test_bvh_equals_brute_force_closest_hit_inside_root_cube(scenes):
    """The BVH only changes which triangles are tested: with the root-cube acceptance rule it must reproduce the
    brute-force intersector (no_acceleration_intersector.rs) wherever the hit point lies inside the scene AABB."""
    w, h = 480, 270
    for name in ("thai2", "ico2"):
        s = scenes(name)
        o = Oracle(s, w, h)
        o.configure(recursions=0, jitter=JITTER_FIXED, intersector=ISECT_BRUTE)
        o.trace_rows(0, h, 1, threads=0)
        t = gpu_tracer(s, w, h, rt.ACCEL_BVH)
        t.trace_rows(0, h, 1)
        assert np.array_equal(t.get_primary_ids(), o.get_primary_ids())
        t.close()


def test_create_raytracer_from_file_defaults(scenes):
    """lib.rs:22-27 with the reference's compile-time defaults; a full default-mode frame must at least agree with the
    pinned oracle on which pixels are hit (jitter moves edges by < 1 pixel) and accumulate one sample everywhere."""
    path = os.path.join(os.path.dirname(GOLDEN), "..", "data", "ico2.dae")
    t = rt.create_raytracer_from_file(os.path.abspath(path), rt.DEFAULT_TRIANGLES_PER_LEAF, 256, 150)
    t.configure(recursions=0, jitter_mode=rt.JITTER_HASHED, seed=0, accel=rt.ACCEL_BVH)
    for _ in range(3):
        assert t.trace_frame_additive() == 50 * 256
    assert (t.film.pixel_datas()[:, 6] == 1).all()
    doc = open(os.path.abspath(path)).read()
    t2 = rt.create_raytracer(doc, 70, 256, 150)
    t2.configure(recursions=0, jitter_mode=rt.JITTER_HASHED, seed=0, accel=rt.ACCEL_BVH)
    for _ in range(3):
        t2.trace_frame_additive()
    assert np.array_equal(t.get_tonemapped_pixels(), t2.get_tonemapped_pixels())
    t.close(), t2.close()


def test_sharded_handles_reassemble_the_frame(scenes):
    """Two shards on one GPU (sequential, independent handles): owned rows compacted on the device, reassembled on
    the host, equal to the unsharded frame — the single-GPU stand-in for the N-GPU gather."""
    import torch

    from raytracer_rs_b200.multi_gpu import _DevPtr, assemble_frame, band_partition

    w, h, world = 320, 200, 3
    s = scenes("ico3_tex")
    full = gpu_tracer(s, w, h, rt.ACCEL_BVH)
    full.trace_rows(0, h, 1)
    ref = full.get_tonemapped_pixels()
    parts = band_partition(h, world, 8)
    compact = []
    total_primary = 0
    for r in range(world):
        t = gpu_tracer(s, w, h, rt.ACCEL_BVH, shard_index=r, shard_count=world, band_rows=8)
        n_primary, _ = t.trace_rows(0, h, 1)
        total_primary += n_primary
        nbytes = len(parts[r]) * w * 4
        buf = t.device_alloc(nbytes)
        assert t.owned_ldr_rows_to(buf) == len(parts[r])  # gather_rows_kernel: owned rows, compacted, on the device
        ldr = t.get_tonemapped_pixels().reshape(h, w)  # also synchronises the handle's stream
        assert (ldr[np.setdiff1d(np.arange(h), parts[r])] == 0xFFFFFFFF).all()  # rows of other shards stay unsampled
        host = torch.as_tensor(_DevPtr(buf, nbytes), device="cuda").cpu().numpy().view(np.uint32).copy()
        assert np.array_equal(host, ldr[parts[r]].reshape(-1))
        compact.append(host)
        t.device_free(buf)
        t.close()
    assert total_primary == w * h
    assert np.array_equal(assemble_frame(compact, parts, w, h), ref)
    full.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
@pytest.mark.parametrize("name,w,h", [("ico2", 256, 192), ("thai2", 320, 180), ("ico3_tex", 320, 180)])
def test_bounce_rays_reference_default_mode(scenes, name, w, h, accel, aname):
    """a6: RECURSIONS = 2, SUB_SPREAD = 1 (mod.rs:81-82, 132-196): up to 4 bounce + 5 shadow rays per primary hit.
    The reference's bounce directions are OS-random (Q5); oracle and GPU draw them from the shared counter hash and the
    shared 65 536-entry unit-vector table, so the comparison is exact: same rays, same counts, colour within 1 LSB."""
    s = scenes(name)
    o = Oracle(s, w, h)
    o.configure(recursions=2, sub_spread=1, jitter=JITTER_HASHED, seed=11)
    o.trace_rows(0, h, 2, threads=0)
    t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=2, sub_spread=1, jitter_mode=rt.JITTER_HASHED, seed=11, accel=accel))
    n_primary, n_shadow = t.trace_rows(0, h, 2)
    c = o.counters()
    st = t.launch_stats()
    assert n_primary == c["rays"]["primary"]
    ids_equal = float((t.get_primary_ids() == o.get_primary_ids()).mean())
    ldr, ldr_o = t.get_tonemapped_pixels(), o.get_tonemapped_pixels()
    if accel == rt.ACCEL_OCTREE:
        assert ids_equal == 1.0
        assert st["n_bounce"] == c["rays"]["bounce"] and n_shadow == c["rays"]["shadow"]
        assert channel_diff(ldr, ldr_o) <= LSB_BAR
    else:
        # BVH: a bounce ray may resolve an exact-t tie or a hit outside the root cube differently
        assert ids_equal >= ID_BAR
        assert abs(st["n_bounce"] - c["rays"]["bounce"]) <= 1e-4 * c["rays"]["bounce"]
        within = np.ones(ldr.shape, bool)
        for k in (0, 8, 16):
            within &= np.abs(((ldr >> k) & 255).astype(np.int32) - ((ldr_o >> k) & 255).astype(np.int32)) <= LSB_BAR
        assert within.mean() >= ID_BAR
    t.close()


@pytest.mark.parametrize("accel,aname", ACCELS)
@pytest.mark.parametrize("rec,spread", [(2, 1), (1, 2), (3, 1)])
def test_bounce_wavefront_equals_depth_first(scenes, accel, aname, rec, spread):
    """f-2: the bounce wavefront (hits of every level compacted into a dense list, next level's rays one per thread,
    bottom-up combine) traces the same rays and leaves the same film as the depth-first walk inside the trace kernel."""
    if accel == rt.ACCEL_LBVH and rec == 3:
        pytest.skip("same traversal kernel as bvh")
    s = scenes("ico3_tex")
    w, h = 384, 216
    got = []
    for wavefront in (1, 0):
        t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=rec, sub_spread=spread, jitter_mode=rt.JITTER_HASHED, seed=4, accel=accel))
        t.set_tuning(6, wavefront)
        n_primary, n_shadow = t.trace_rows(0, h, 2)
        st = t.launch_stats()
        got.append((n_shadow, st["n_bounce"], t.get_primary_ids(), t.get_tonemapped_pixels(), t.film.pixel_datas().view(np.uint32)))
        t.close()
    a, b = got
    assert a[0] == b[0] and a[1] == b[1] and a[1] > 0
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(x, y)


def test_bounce_sample_table_matches_oracle(scenes):
    """sample_generator.rs:9-53 stand-in: host-generated table == oracle table, unit length, and recursion depth 1 and 3
    also agree with the oracle."""
    s = scenes("4boxes")
    w, h = 128, 72
    for rec in (1, 3):
        o = Oracle(s, w, h)
        o.configure(recursions=rec, sub_spread=2, jitter=JITTER_FIXED, seed=5)
        o.trace_rows(0, h, 1, threads=0)
        t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=rec, sub_spread=2, jitter_mode=rt.JITTER_FIXED_HALF, seed=5, accel=rt.ACCEL_OCTREE))
        t.trace_rows(0, h, 1)
        assert t.launch_stats()["n_bounce"] == o.counters()["rays"]["bounce"] > 0
        assert channel_diff(t.get_tonemapped_pixels(), o.get_tonemapped_pixels()) <= LSB_BAR
        t.close()
    tab = o.sample_table()
    assert np.allclose(np.linalg.norm(tab, axis=1), 1.0, atol=1e-6)


def test_sharded_handle_with_a_row_range_longer_than_the_image(scenes):
    """rt_trace_rows(first, n_rows > height) passes over some rows twice. A launch must never hold a pixel twice (two
    warps would update its film record concurrently), so the call is split into laps — also on a sharded handle, whose
    laps are the rows it owns of each lap. The shards' films must equal the unsharded film on the rows they own.
    (Last test of the GPU suite on purpose: it exercises a host path no other test or benchmark uses.)"""
    w, h, world = 320, 90, 2
    first, n_rows = 70, 2 * h + 37  # wraps, 2.4 laps
    s = scenes("ico2")
    full = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=3)
    n_full, sh_full = full.trace_rows(first, n_rows, 1)
    assert n_full == n_rows * w
    ref = full.film.pixel_datas().view(np.uint32).reshape(h, w, 7)
    ref_ldr = full.get_tonemapped_pixels().reshape(h, w)
    total, total_shadow = 0, 0
    for r in range(world):
        t = gpu_tracer(s, w, h, rt.ACCEL_BVH, jitter=rt.JITTER_HASHED, seed=3, shard_index=r, shard_count=world, band_rows=8)
        n, sh = t.trace_rows(first, n_rows, 1)
        total, total_shadow = total + n, total_shadow + sh
        own = ((np.arange(h) // 8) % world) == r
        film = t.film.pixel_datas().view(np.uint32).reshape(h, w, 7)
        assert np.array_equal(film[own], ref[own])
        assert np.array_equal(t.get_tonemapped_pixels().reshape(h, w)[own], ref_ldr[own])
        assert (film[~own][:, :, 6] == 0).all()  # rows of the other shard stay unsampled
        t.close()
    assert (total, total_shadow) == (n_full, sh_full)
    full.close()
