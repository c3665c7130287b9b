"""GPU tests added in round 2: Film::get_estimated_variances (film.rs:50-67), the incremental readback of the reference's
band loop (main.rs:200-201), per-call ray counters without a memset, the per-geometry tile schedules (bands, camera moves,
finer heavy-tile splits) and the kernel-variant switch on a live handle. Everything goes through the C ABI."""
import numpy as np
import pytest

import raytracer_rs_b200 as rt
from conftest import CONFIGS
from oracle_lib import JITTER_FIXED, JITTER_HASHED, Oracle

pytestmark = pytest.mark.gpu


def tracer_for(scene, w, h, accel=rt.ACCEL_BVH, jitter=rt.JITTER_FIXED_HALF, seed=0, **kw):
    return rt.RayTracer.from_scene(scene, rt.Config(w, h, recursions=0, jitter_mode=jitter, seed=seed, accel=accel, **kw))


def same_bits(a, b):
    """bit-for-bit equality except that any NaN equals any NaN (0/0 is -qNaN on x86 and +qNaN on the GPU)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(nan | (a.view(np.uint32) == b.view(np.uint32))))


def test_estimated_variances_match_the_oracle(scenes):
    """f-4: rows with 3 samples, rows with 1 sample (x/0 - y/0 = NaN) and never-sampled rows (n = 0: the wrapping u32 product
    n * (n - 1) is 0, NaN again), hashed jitter so that the samples of a pixel differ. The oracle evaluates film.rs:50-67 on the
    GPU's own film (its sums differ from the oracle's render by the documented <= 1 ulp of powf on a few pixels)."""
    w, h = 160, 90
    s = scenes("ico3_tex")
    t = tracer_for(s, w, h, jitter=rt.JITTER_HASHED, seed=3)
    t.trace_rows(0, 60, 1)
    t.trace_rows(0, 30, 2)
    film = t.film.pixel_datas()
    assert set(np.unique(film[:, 6])) == {0.0, 1.0, 3.0}
    var = t.film.get_estimated_variances()
    o = Oracle(s, w, h)
    o.set_film(film)
    var_o = o.get_estimated_variances()
    assert same_bits(var, var_o)
    n = film[:, 6]
    assert np.isnan(var[n <= 1]).all() and np.isfinite(var[n == 3]).all()
    assert (var[n == 3] >= -1e-3).all() and var[n == 3].max() > 0  # an unbiased variance estimate, times 50
    # the whole render agrees with the oracle's as well wherever the two films are bit-identical
    o2 = Oracle(s, w, h)
    o2.configure(recursions=0, jitter=JITTER_HASHED, seed=3)
    o2.trace_rows(0, 60, 1, threads=0)
    o2.trace_rows(0, 30, 2, threads=0)
    film_o = o2.get_film()
    eq = (film.view(np.uint32) == film_o.view(np.uint32)).all(axis=1)
    assert eq.mean() > 0.99
    assert same_bits(var[eq], o2.get_estimated_variances()[eq])
    t.close()


def test_estimated_variances_wrapping_sample_count(scenes):
    """n * (n - 1) wraps in u32 before the conversion to f32 (the reference's release build): n = 65 537 gives 65 536, n = 2^31
    gives 2^31 (wrapped) — injected through rt_set_film, compared with the oracle on the same film."""
    w, h = 8, 4
    s = scenes("4boxes")
    t = tracer_for(s, w, h)
    rng = np.random.default_rng(5)
    film = np.zeros((w * h, 7), np.float32)
    film[:, 0:3] = rng.uniform(0, 50, (w * h, 3)).astype(np.float32)
    film[:, 3:6] = rng.uniform(0, 500, (w * h, 3)).astype(np.float32)
    film[:, 6] = np.array([0, 1, 2, 3, 65536, 65537, 92682, 2 ** 31] * 4, np.float32)
    t.film.set_pixel_datas(film)
    assert np.array_equal(t.film.pixel_datas(), film)
    o = Oracle(s, w, h)
    o.set_film(film)
    assert same_bits(t.film.get_estimated_variances(), o.get_estimated_variances())
    assert np.array_equal(t.get_tonemapped_pixels(), o.get_tonemapped_pixels())  # the packed frame follows the film
    t.close()


def test_delta_readback_follows_the_band_loop(scenes):
    """The reference's loop — trace_frame_additive (50 rows), get_tonemapped_pixels (whole frame), main.rs:200-201 — with the
    incremental readback into one buffer the host keeps: after every call the buffer equals the full readback, across the wrap
    at the bottom of the image, a camera move with film.clear() (main.rs:124-162) and a change of buffer."""
    w, h = 320, 180
    t = tracer_for(scenes("thai2"), w, h)
    keep = np.zeros(w * h, np.uint32)
    for call in range(9):  # 9 x 50 rows: two and a half laps
        t.trace_frame_additive()
        t.get_tonemapped_pixels_delta_into(keep.ctypes.data)
        assert np.array_equal(keep, t.get_tonemapped_pixels()), call
        if call == 4:
            t.camera.move_rel(0.1, 0.0, 0.2)
            t.film.clear()
    other = np.full(w * h, 0x12345678, np.uint32)  # a buffer the library has never seen: whole frame
    t.get_tonemapped_pixels_delta_into(other.ctypes.data)
    assert np.array_equal(other, keep)
    t.trace_rows(10, 3, 1)
    t.get_tonemapped_pixels_delta_into(other.ctypes.data)
    assert np.array_equal(other, t.get_tonemapped_pixels())
    t.close()


def test_per_call_ray_counters_without_host_resets(scenes):
    """Consecutive trace calls with nothing in between (no stream synchronisation, no counter read): every call's shadow-ray
    count is exact — the two per-call counter sets alternate and the last warp out of a launch zeroes the other one."""
    w, h = 256, 144
    s = scenes("ico2")
    o = Oracle(s, w, h)
    o.configure(recursions=0, jitter=JITTER_FIXED)
    expect = []
    for first, rows in ((0, 144), (0, 50), (50, 50), (100, 44), (0, 144)):
        o.counters(reset=True)
        o.trace_rows(first, rows, 1, threads=0)
        expect.append(o.counters()["rays"]["shadow"])
    for variant in (1, 2, 0):
        t = tracer_for(s, w, h)
        t.set_tuning(0, variant)
        t.trace_rows(0, 144, 1, want_shadow=False)  # asynchronous calls first: their counters are never read
        t.trace_rows(0, 50, 1, want_shadow=False)
        got = [t.trace_rows(first, rows, 1)[1] for first, rows in ((50, 50), (100, 44), (0, 144))]
        assert got == expect[2:], (variant, got, expect)
        tot = t.ray_totals()
        assert tot["shadow"] == sum(expect) and tot["primary"] == w * (144 + 50 + 50 + 44 + 144)
        t.close()


@pytest.mark.parametrize("name", ["thai2", "ico3_tex"])
def test_band_schedules_and_fine_splits_leave_the_film_alone(scenes, name):
    """Schedules only reorder and regroup work. A handle that walks over a 1080p frame in 50-row bands with per-band cost feedback,
    a split threshold low enough that heavy tiles become 8 and 16 items, and a camera move in between leaves the same film, ids
    and frame, bit for bit, as a handle that runs everything in image order."""
    _, w, h = CONFIGS[name]
    s = scenes(name)
    a, b = tracer_for(s, w, h), tracer_for(s, w, h)
    a.set_tuning(1, 0)  # image order, no cost feedback
    b.set_tuning(11, 256)  # schedule even small launches
    b.set_tuning(7, 1)  # split from a quarter of the balanced launch time on
    calls = 3 * ((h + 49) // 50)
    for c in range(calls):
        na, nb = a.trace_frame_additive(), b.trace_frame_additive()
        assert na == nb == 50 * w
        if c == calls // 2:
            for t in (a, b):
                t.camera.add_y_angle(0.05)  # the film is deliberately not cleared: old and new view accumulate
    assert np.array_equal(a.get_primary_ids(), b.get_primary_ids())
    assert np.array_equal(a.get_tonemapped_pixels(), b.get_tonemapped_pixels())
    assert np.array_equal(a.film.pixel_datas().view(np.uint32), b.film.pixel_datas().view(np.uint32))
    # whole frames on the same handles (another launch geometry, with its own schedule); 4 frames so that the sorted, split
    # queue is in use
    for _ in range(4):
        a.trace_rows(0, h, 1, want_shadow=False)
        b.trace_rows(0, h, 1, want_shadow=False)
    assert np.array_equal(a.film.pixel_datas().view(np.uint32), b.film.pixel_datas().view(np.uint32))
    assert np.array_equal(a.get_tonemapped_pixels(), b.get_tonemapped_pixels())
    a.close()
    b.close()


def test_sharded_sample_lanes_with_fine_splits(scenes):
    """The shard one rank of 8 owns, 8 samples per launch in sample lanes, low split threshold: parts must hold whole pixels
    (a part of an S = 8 item is 8 lanes), so the film equals one launch per sample."""
    w, h = 1920, 1080
    s = scenes("thai2")
    films = []
    for multi, split in ((0, 4), (1, 1)):
        t = tracer_for(s, w, h, jitter=rt.JITTER_HASHED, seed=11, shard_index=3, shard_count=8, band_rows=8)
        t.set_tuning(5, multi)
        t.set_tuning(7, split)
        for _ in range(4):
            t.trace_rows(0, h, 8, want_shadow=False)
        films.append(t.film.pixel_datas())
        t.close()
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))
    assert films[0][:, 6].max() == 32


def test_switching_the_kernel_variant_on_a_live_handle(scenes):
    """rt_set_tuning never changes results: after several 1080p frames of the persistent kernel (sorted queue with split
    tiles) the handle switches to the ray-pool kernel, which takes whole tiles — an order written for the other kernel must
    not reach it (it would trace split tiles once per part). Film and sample counts equal a fresh handle's."""
    _, w, h = CONFIGS["thai2"]
    s = scenes("thai2")
    t = tracer_for(s, w, h)
    for _ in range(4):
        t.trace_rows(0, h, 1, want_shadow=False)
    t.set_tuning(0, 2)
    for _ in range(4):
        t.trace_rows(0, h, 1, want_shadow=False)
    t.set_tuning(0, 1)
    t.trace_rows(0, h, 1, want_shadow=False)
    ref = tracer_for(s, w, h)
    for _ in range(9):
        ref.trace_rows(0, h, 1, want_shadow=False)
    f, fr = t.film.pixel_datas(), ref.film.pixel_datas()
    assert np.array_equal(f[:, 6], np.full(w * h, 9, np.float32))
    assert np.array_equal(f.view(np.uint32), fr.view(np.uint32))
    assert np.array_equal(t.get_tonemapped_pixels(), ref.get_tonemapped_pixels())
    # the same with the structure: costs recorded for one tree must not schedule another
    t.configure(recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH4)
    ref.configure(recursions=0, jitter_mode=rt.JITTER_FIXED_HALF, accel=rt.ACCEL_BVH4)
    for _ in range(3):
        t.trace_rows(0, h, 1, want_shadow=False)
        ref.trace_rows(0, h, 1, want_shadow=False)
    assert np.array_equal(t.film.pixel_datas().view(np.uint32), ref.film.pixel_datas().view(np.uint32))
    t.close()
    ref.close()


def test_reserved_normals_and_uvs_are_ignored(scenes):
    """rt_scene_desc.normals / uvs (north star: "triangle, normal, UV and texture buffers") are accepted and, as in the reference
    (colladaloader.rs:587-593 reads past them), have no influence on the image."""
    import types

    w, h = 160, 90
    s = scenes("ico3_tex")
    rng = np.random.default_rng(1)
    n = s.vertices.shape[0]
    with_attrs = types.SimpleNamespace(**{k: getattr(s, k) for k in ("vertices", "tri_geom", "materials", "lights", "textures",
                                                                     "camera_orientation", "camera_fov_deg")},
                                       normals=rng.normal(size=(n, 9)).astype(np.float32), uvs=rng.uniform(size=(n, 6)).astype(np.float32))
    a, b = tracer_for(s, w, h), tracer_for(with_attrs, w, h)
    a.trace_rows(0, h, 1)
    b.trace_rows(0, h, 1)
    assert np.array_equal(a.get_tonemapped_pixels(), b.get_tonemapped_pixels())
    assert np.array_equal(a.film.pixel_datas().view(np.uint32), b.film.pixel_datas().view(np.uint32))
    a.close()
    b.close()


@pytest.mark.parametrize("accel", [rt.ACCEL_BVH, rt.ACCEL_LBVH])
@pytest.mark.parametrize("name,rec,spread,lights", [("thai2", 2, 1, 1), ("ico3_tex", 2, 1, 1), ("ico2", 3, 1, 1), ("4boxes", 1, 3, 3)])
def test_bounce_ray_stream_equals_lockstep_wavefront_and_depth_first(scenes, accel, name, rec, spread, lights):
    """f-2, RECURSIONS > 0 (mod.rs:132-196): the ray-stream form of a wavefront level (bounce rays and the shadow rays of their hits
    share the lanes of a warp, finished lanes are refilled) traces the same rays and leaves the same film, bit for bit, as the
    lockstep wavefront and as the depth-first walk — for every refill / yield threshold, with the bounce levels chained in one launch (a lane continues with the bounce ray of the hit it
    just completed) or launched level by level, at every occupancy the kernel is compiled for, with textures, with several lights (shade()'s
    loop continues in place from one shadow ray to the next) and with deeper recursion."""
    import types

    s = scenes(name)
    if lights > 1:  # two more lights, one of them behind most surfaces
        s = types.SimpleNamespace(**{k: getattr(s, k) for k in ("vertices", "tri_geom", "materials", "textures", "camera_orientation", "camera_fov_deg")},
                                  lights=list(s.lights) + [(np.float32([-6, 8, -3]), np.float32([3, 2, 1])), (np.float32([0, -9, 0]), np.float32([1, 1, 4]))])
    w, h = 480, 270
    runs = []
    # (wavefront, ray stream, refill threshold, inner-loop yield threshold, resident blocks per SM, bounce levels chained in one launch)
    for wavefront, stream, refill, min_inner, blocks, chain in ((1, 1, 16, 8, 4, 1), (1, 1, 1, 0, 3, 1), (1, 1, 32, 32, 5, 0), (1, 1, 20, 4, 3, 0),
                                                                (1, 0, 8, 8, 3, 0), (0, 0, 8, 8, 3, 0)):
        t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=rec, sub_spread=spread, jitter_mode=rt.JITTER_HASHED, seed=6, accel=accel))
        t.set_tuning(6, wavefront)
        t.set_tuning(13, stream)
        t.set_tuning(14, refill)
        t.set_tuning(15, min_inner)
        t.set_tuning(16, blocks)
        t.set_tuning(18, chain)
        n_shadow = [t.trace_rows(0, h, 1)[1], t.trace_rows(40, 100, 1)[1]]
        st = t.launch_stats()
        runs.append((n_shadow, st["n_bounce"], t.get_primary_ids(), t.get_tonemapped_pixels(), t.film.pixel_datas().view(np.uint32)))
        t.close()
    ref = runs[-1]
    assert ref[1] > 0
    for r in runs[:-1]:
        assert r[0] == ref[0] and r[1] == ref[1]
        for x, y in zip(r[2:], ref[2:]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("world,height", [(2, 364), (3, 90), (8, 1080), (5, 37)])
def test_copy_owned_rows_moves_exactly_the_owned_bands(scenes, world, height):
    """rt_copy_owned_rows (one strided 2-D copy for the full bands a shard owns, one more for a partial band at the bottom): every shard
    copies its rows of a device frame into one host frame; together they reproduce the frame, and no shard touches a row it does not own."""
    import torch

    w = 96
    s = scenes("ico2")
    src = torch.arange(w * height, dtype=torch.int32, device="cuda") * 7 + 1
    assembled = torch.zeros(w * height, dtype=torch.int32).pin_memory()
    for r in range(world):
        t = tracer_for(s, w, height, shard_index=r, shard_count=world, band_rows=8)
        mine = torch.full((w * height,), -1, dtype=torch.int32).pin_memory()
        t.copy_owned_rows(src.data_ptr(), mine.data_ptr())
        t.copy_owned_rows(src.data_ptr(), assembled.data_ptr())
        torch.cuda.synchronize()
        rows = np.arange(height)
        owned = (rows // 8) % world == r
        got = mine.numpy().reshape(height, w)
        assert (got[owned] == src.cpu().numpy().reshape(height, w)[owned]).all() and (got[~owned] == -1).all()
        t.close()
    assert torch.equal(assembled, src.cpu())


@pytest.mark.parametrize("name,w,h,rec", [("thai2", 640, 360, 0), ("ico3_tex", 320, 170, 0), ("ico2", 256, 190, 2)])
def test_band_lookahead_equals_band_by_band_tracing(scenes, name, w, h, rec):
    """rt_trace_frame_additive with the lap traced ahead (one launch traces the next sample of every remaining row, band calls commit
    their rows) against the same handle settings tracing every band on its own: after EVERY call the frame, and at the end film, ids and
    per-call ray counts, are bit-identical — with hashed jitter (a sample's number is its pixel's film count), across the wrap at the
    bottom of the image (heights that are no multiple of 50), a camera move + film.clear() in the middle of a lap, an explicit
    rt_trace_rows in the middle of a lap, a change of rows_per_call, and with bounce rays (RECURSIONS = 2)."""
    s = scenes(name)
    cfg = dict(recursions=rec, sub_spread=1, jitter_mode=rt.JITTER_HASHED, seed=8, accel=rt.ACCEL_BVH)
    a = rt.RayTracer.from_scene(s, rt.Config(w, h, **cfg))
    b = rt.RayTracer.from_scene(s, rt.Config(w, h, **cfg))
    a.set_tuning(19, 0)  # every call traces its own rows
    keep_a, keep_b = np.zeros(w * h, np.uint32), np.zeros(w * h, np.uint32)
    calls = 4 * ((h + 49) // 50) + 3
    for c in range(calls):
        na, nb = a.trace_frame_additive(), b.trace_frame_additive()
        sa, sb = a.launch_stats(), b.launch_stats()
        assert na == nb and (sa["n_primary"], sa["n_shadow"], sa["n_bounce"]) == (sb["n_primary"], sb["n_shadow"], sb["n_bounce"]), c
        a.get_tonemapped_pixels_delta_into(keep_a.ctypes.data)
        b.get_tonemapped_pixels_delta_into(keep_b.ctypes.data)
        assert np.array_equal(keep_a, keep_b), c
        if c == 5:  # a key press in the middle of a lap (main.rs:124-162)
            for t in (a, b):
                t.camera.add_y_angle(0.07)
                t.film.clear()
        if c == 9:  # an explicit row range in the middle of a lap: its film counts move
            for t in (a, b):
                t.trace_rows(3, 21, 2)
        if c == 12:
            for t in (a, b):
                t.set_rows_per_call(70)
    assert np.array_equal(a.get_primary_ids(), b.get_primary_ids())
    assert np.array_equal(a.film.pixel_datas().view(np.uint32), b.film.pixel_datas().view(np.uint32))
    assert np.array_equal(a.get_tonemapped_pixels(), b.get_tonemapped_pixels())
    ta, tb = a.ray_totals(), b.ray_totals()
    assert ta["primary"] == tb["primary"] and ta["shadow"] == tb["shadow"] and ta["bounce"] == tb["bounce"]
    a.close()
    b.close()


@pytest.mark.parametrize("accel", [rt.ACCEL_BVH, rt.ACCEL_LBVH, rt.ACCEL_OCTREE])
def test_l2_prefetch_levels_leave_the_film_alone(scenes, accel):
    """RT_TUNE_FILM_PREFETCH 0..3 (film record, sums of squares of a hit, the whole tree + tile queue at the start of a launch) are
    hints to the cache: film, ids, frame and ray counts are the same for every level, on a full frame (cost-sorted queue from the
    second launch on) and on a wrapping band."""
    w, h = 320, 184
    s = scenes("thai2")
    want = None
    for level, rows_mb in ((0, 48), (1, 48), (2, 48), (3, 48), (3, 0)):
        t = tracer_for(s, w, h, accel=accel, jitter=rt.JITTER_HASHED, seed=5)
        t.set_tuning(20, level)
        t.set_tuning(21, rows_mb)  # hashed jitter: the launch also requests the film records of its rows up front
        shadow = [t.trace_rows(0, h, 1)[1] for _ in range(3)]
        shadow.append(t.trace_rows(h - 20, 50, 1)[1])
        got = (t.film.pixel_datas().tobytes(), t.get_primary_ids().tobytes(), t.get_tonemapped_pixels().tobytes(), tuple(shadow))
        t.close()
        if want is None:
            want = got
        assert got == want, (level, rows_mb)


def _render_digest(t, h, calls):
    """film + ids + frame + per-call shadow-ray counts of a sequence of trace calls"""
    shadow = [t.trace_rows(first, rows, spp)[1] for first, rows, spp in calls]
    return (t.film.pixel_datas().tobytes(), t.get_primary_ids().tobytes(), t.get_tonemapped_pixels().tobytes(), tuple(shadow))


@pytest.mark.parametrize("name,w,h", [("thai2", 480, 272), ("ico3_tex", 320, 180), ("4boxes", 256, 144)])
@pytest.mark.parametrize("accel", [rt.ACCEL_BVH, rt.ACCEL_LBVH])
def test_perspective_grids_equal_the_tree_walk(scenes, name, w, h, accel):
    """Camera rays through the perspective grid of the view and shadow rays through the cube of grids around the light (trace kernels
    ACCEL = 4, csrc/pgrid_build.cu) give the film, ids, frame and ray counts of the tree walk, bit for bit: every cell size, grids built at
    once or from the second launch of a view, full frames, a wrapping band, two samples per call (sample lanes), bounce rays."""
    s = scenes(name)
    calls = [(0, h, 1), (0, h, 1), (h - 20, 50, 1), (0, h, 2), (0, h, 1)]
    for rec in (0, 2):
        want = None
        for grid, after, light in ((0, 1, 0), (3, 0, 0), (3, 1, 8), (2, 0, 6), (4, 0, 9), (5, 1, 7)):
            t = rt.RayTracer.from_scene(s, rt.Config(w, h, recursions=rec, sub_spread=1, jitter_mode=rt.JITTER_HASHED, seed=11, accel=accel))
            t.set_tuning(22, grid)
            t.set_tuning(23, after)
            if light:
                t.set_tuning(24, light)
            else:
                t.set_tuning(24, 0)
            got = _render_digest(t, h, calls)
            t.close()
            if want is None:
                want = got
            assert got == want, (rec, grid, after, light)


def test_perspective_grid_follows_the_camera(scenes):
    """The grid belongs to one view: moves, turns and a view from inside the scene (triangles behind the eye, triangles cut by the eye
    plane) rebuild it; every frame equals the tree walk's frame of the same view."""
    w, h = 384, 216
    s = scenes("thai2")
    moves = [("move", (0.0, 0.0, 0.3)), ("y", 0.4), ("x", -0.2), ("move", (0.3, -0.2, 1.5)), ("y", 1.3), ("move", (0.0, 0.0, 2.0)), ("x", 0.9)]
    frames = {}
    for grid in (0, 3):
        t = tracer_for(s, w, h, jitter=rt.JITTER_HASHED, seed=2)
        t.set_tuning(22, grid)
        t.set_tuning(23, 0)
        out = []
        for kind, arg in moves:
            if kind == "move":
                t.camera.move_rel(*arg)
            elif kind == "y":
                t.camera.add_y_angle(arg)
            else:
                t.camera.add_x_angle(arg)
            t.film.clear()
            shadow = [t.trace_rows(0, h, 1)[1] for _ in range(2)]
            out.append((t.film.pixel_datas().tobytes(), t.get_primary_ids().tobytes(), tuple(shadow)))
        frames[grid] = out
        t.close()
    hit_counts = [int((np.frombuffer(f[1], np.uint32) != 0xFFFFFFFF).sum()) for f in frames[0]]
    assert max(hit_counts) > 1000  # the walk does look at the scene
    for k, (a, b) in enumerate(zip(frames[0], frames[3])):
        assert a == b, (k, moves[k])


def test_light_grid_with_several_lights_and_a_light_close_to_a_surface(scenes):
    """Three lights, one of them a hair away from a surface (its shadow rays can reach geometry lying beyond the light within their last
    hundredth: those rays walk the tree) and one far outside the scene: same film as with shadow rays walking the tree."""
    w, h = 320, 180
    base = scenes("ico2")
    import types

    v = base.vertices.reshape(-1, 3, 3)
    near = v[0].mean(0) + 1e-3 * np.cross(v[0][1] - v[0][0], v[0][2] - v[0][0])  # a hair above the first triangle
    lights = [(np.float32([10, 10, 10]), np.float32([10, 10, 10])), (near.astype(np.float32), np.float32([3, 2, 1])),
              (np.float32([-40, 25, 60]), np.float32([5, 5, 9]))]
    s = types.SimpleNamespace(**{k: getattr(base, k) for k in ("vertices", "tri_geom", "materials", "textures", "camera_orientation", "camera_fov_deg")},
                              lights=lights)
    want = None
    for light in (0, 8, 6):
        t = tracer_for(s, w, h, jitter=rt.JITTER_HASHED, seed=4)
        t.set_tuning(23, 0)
        t.set_tuning(24, light)
        got = _render_digest(t, h, [(0, h, 1), (0, h, 1)])
        t.close()
        if want is None:
            want = got
        assert got == want, light
