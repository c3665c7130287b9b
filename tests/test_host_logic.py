"""Host-side logic of the product (CPU only, handles created with RT_DEVICE_NONE): Collada loader, octree build,
BVH build, camera, Stats / BenchMark stand-ins — each against the oracle or an independent restatement."""
import os
import re
import time

import numpy as np
import pytest

import raytracer_rs_b200 as rt
from collada_ref import load_collada
from conftest import CONFIGS, DATA
from oracle_lib import JITTER_FIXED, Oracle
from raytracer_rs_b200.api import DEVICE_NONE

F = np.float32


def host_tracer(scene, w=64, h=36, **kw):
    return rt.RayTracer.from_scene(scene, rt.Config(w, h, device=DEVICE_NONE, recursions=0, **kw))


# ---- loader (colladaloader.rs) ---------------------------------------------------------------------------------


@pytest.mark.parametrize("name", list(CONFIGS))
def test_loader_bit_identical_to_numpy_restatement(scenes, name):
    """Every vertex, light, material, texel and camera float equals the independent restatement in collada_ref.py
    (Q13, Q14: geometry order = node order, v * (reflect_z * M^T * swap_yz) in f32)."""
    s, r = scenes(name), load_collada(os.path.join(DATA, CONFIGS[name][0]))
    assert np.array_equal(s.vertices.view(np.uint32), r.vertices.view(np.uint32))
    assert np.array_equal(s.tri_geom, r.tri_geom)
    assert len(s.materials) == len(r.materials)
    for a, b in zip(s.materials, r.materials):
        assert a[0] == b[0] and a[2] == b[2] and all(F(x) == F(y) for x, y in zip(a[1], b[1]))
    assert len(s.lights) == len(r.lights)
    for a, b in zip(s.lights, r.lights):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(s.camera_orientation, r.camera_orientation) and s.camera_fov_deg == r.camera_fov_deg
    assert len(s.textures) == len(r.textures)
    for a, b in zip(s.textures, r.textures):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_scene_inventory(scenes):
    """SURVEY.md section 8 scene table."""
    expect = {"ico2": (608, 5, 1, 0), "4boxes": (48, 4, 1, 0), "ico3_tex": (608, 5, 1, 1), "thai2": (20049, 2, 1, 0)}
    for name, (tris, geoms, lights, texs) in expect.items():
        s = scenes(name)
        assert (s.vertices.shape[0], len(s.materials), len(s.lights), len(s.textures)) == (tris, geoms, lights, texs)
    tex = scenes("ico3_tex")
    assert tex.textures[0][:2] == (640, 640) and tex.materials[1][0] == 1  # Cube_004 is textured
    assert tex.textures[0][2].max() <= 255 / 256  # texels are byte / 256 (texture.rs:42-44)
    assert tuple(scenes("thai2").lights[0][1]) == (1.0, 1.0, 1.0) and tuple(scenes("ico2").lights[0][1]) == (10.0, 10.0, 10.0)


def test_loader_errors_mirror_reference_categories(tmp_path):
    """SceneLoadError / ColladaError Display categories (colladaloader.rs:603-689, loaders/mod.rs:45-55)."""
    with pytest.raises(rt.RtError) as e:
        rt.load_scene(str(tmp_path / "missing.dae"))
    assert "No such file" in e.value.message and e.value.code == -2
    doc = open(os.path.join(DATA, "4boxes.dae")).read()

    def load(text):
        p = tmp_path / "x.dae"
        p.write_text(text)
        return rt.load_scene(str(p))

    with pytest.raises(rt.RtError, match="XmlDefinition error"):
        load(doc.replace('<?xml version="1.0" encoding="utf-8"?>', ""))
    with pytest.raises(rt.RtError, match="Not a collada doc"):
        load(doc.replace("<COLLADA", "<COLLADB").replace("</COLLADA>", "</COLLADB>"))
    with pytest.raises(rt.RtError, match="LibraryCamerasParsing error"):
        load(re.sub(r"<library_cameras>.*?</library_cameras>", "", doc, flags=re.S))  # fixed library order (:71-108)
    with pytest.raises(rt.RtError, match="RemainingData error"):
        load(doc + "<extra/>")
    with pytest.raises(rt.RtError, match="VisualSceneConversion error; unsupported node type"):
        load(doc.replace("<instance_camera", "<instance_controller"))
    # a geometry whose material is unknown falls back to Material::default() = (1000, 0, 1000) (Q13)
    s = load(doc.replace('material="Material_004-material"', 'material="nope"'))
    assert any(tuple(m[1]) == (1000.0, 0.0, 1000.0) for m in s.materials)
    with pytest.raises(rt.RtError) as e:  # texture file missing -> TextureLoadError text
        bad = open(os.path.join(DATA, "ico3_tex.dae")).read()
        load(bad)  # tmp dir has no png next to the .dae
    assert e.value.code == -2


def test_create_raytracer_from_str_and_file_errors():
    """lib.rs:15-27: Result<RayTracer, String> -> error code + message, no exception escapes the C ABI."""
    with pytest.raises(rt.RtError):
        rt.create_raytracer("not xml", 70, 64, 36)
    with pytest.raises(rt.RtError):
        rt.create_raytracer_from_file("/nonexistent/file.dae", 70, 64, 36)


# ---- octree build (oct_tree_intersector.rs:66-146, 274-330, 374-469) -----------------------------------------------


@pytest.mark.parametrize("name,tpl", [("4boxes", 70), ("ico2", 70), ("ico2", 5), ("thai2", 70), ("ico3_tex", 20)])
def test_octree_identical_to_oracle(scenes, name, tpl):
    """Node numbering, cubes (bit patterns), child links and leaf triangle lists equal the oracle's recursive build."""
    s = scenes(name)
    r = host_tracer(s, triangles_per_leaf=tpl)
    o = Oracle(s, 64, 36, tpl)
    assert r.octree_stats() == o.octree_stats()
    for mine, ref in zip(r.octree_export(), o.octree_export()):
        assert mine.dtype == ref.dtype and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_octree_depth_cap_and_leaf_order(scenes):
    """depth cap: a node is split only while recurse_level <= 8 (:108); leaf lists stay in ascending triangle order
    (:332-342), which is what makes 'first wins ties' well defined (Q7)."""
    r = host_tracer(scenes("thai2"), triangles_per_leaf=1)
    st = r.octree_stats()
    assert st["depth"] == 9
    cubes, first_child, leaf_offset, leaf_tris = r.octree_export()
    for i in np.nonzero(first_child < 0)[0][:500]:
        seg = leaf_tris[leaf_offset[i]:leaf_offset[i + 1]]
        assert (np.diff(seg.astype(np.int64)) > 0).all()


# ---- BVH build ------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["4boxes", "ico2", "thai2"])
def test_bvh_invariants(scenes, name):
    s = scenes(name)
    r = host_tracer(s)
    boxes, children, counts, order = r.bvh_export()
    n_tri = s.vertices.shape[0]
    assert sorted(order.tolist()) == list(range(n_tri))  # every triangle in exactly one leaf slot
    st = r.bvh_stats()
    assert st["max_leaf"] <= 4 and st["leaves"] == (children < 0).sum() - (counts[children < 0] == 0).sum()
    verts = s.vertices.reshape(n_tri, 3, 3)
    seen = np.zeros(n_tri, bool)

    def subtree_box(node):  # returns exact AABB of the subtree, checks stored (padded) child boxes contain it
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        for k in range(2):
            c, cnt = int(children[node, k]), int(counts[node, k])
            if c >= 0:
                lo, hi = subtree_box(c)
            else:
                first = ~c
                tri = order[first:first + cnt]
                assert not seen[tri].any()
                seen[tri] = True
                if cnt == 0:
                    continue
                assert (np.diff(tri.astype(np.int64)) > 0).all()  # ascending ids inside a leaf (tie-break order)
                lo, hi = verts[tri].reshape(-1, 3).min(0), verts[tri].reshape(-1, 3).max(0)
            assert (boxes[node, k, 0] < lo).all() and (boxes[node, k, 1] > hi).all()  # strictly padded outward
            lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
        return lo_all, hi_all

    import sys
    sys.setrecursionlimit(10000)
    subtree_box(0)
    assert seen.all()


@pytest.mark.parametrize("name", ["ico2", "thai2"])
def test_bvh_reinsertion_keeps_the_leaves_and_shrinks_the_inner_nodes(scenes, name, monkeypatch):
    """The insertion-based optimisation pass of the host BVH (csrc/bvh_build.cpp, Reinserter) moves subtrees, nothing else: the same leaves
    (as sets of triangle ids) as the builder's tree, every subtree still one contiguous run of tri_order (the wide builders rely on it), a
    smaller summed area of the inner nodes, a depth the traversal stacks hold."""
    s = scenes(name)

    def export(passes):
        monkeypatch.setenv("RT_BVH_REINSERT_PASSES", str(passes))
        r = host_tracer(s)
        out = r.bvh_export() + (r.bvh_stats(),)
        r.close()
        return out

    def leaves_and_area(boxes, children, counts, order):
        leaves, area = set(), 0.0
        runs = {}

        def walk(node):  # returns (first slot, count) of the subtree
            first, total = None, 0
            for k in range(2):
                c, cnt = int(children[node, k]), int(counts[node, k])
                if c >= 0:
                    d = boxes[node, k, 1].astype(np.float64) - boxes[node, k, 0]
                    nonlocal area
                    area += d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
                    f, n = walk(c)
                else:
                    f, n = ~c, cnt
                    leaves.add(tuple(order[f:f + n].tolist()))
                assert first is None or f == first + total  # child 1's run follows child 0's
                first = f if first is None else first
                total += n
            return first, total

        import sys
        sys.setrecursionlimit(10000)
        assert walk(0) == (0, len(order))
        return leaves, area

    b0, c0, n0, o0, st0 = export(0)
    b1, c1, n1, o1, st1 = export(8)
    leaves0, area0 = leaves_and_area(b0, c0, n0, o0)
    leaves1, area1 = leaves_and_area(b1, c1, n1, o1)
    assert leaves0 == leaves1 and st0["nodes"] == st1["nodes"] and st0["leaves"] == st1["leaves"]
    assert area1 < 0.99 * area0
    assert st1["depth"] + 2 <= 48  # kBvhStack


@pytest.mark.parametrize("name", ["4boxes", "ico2", "thai2"])
def test_bvh4_invariants(scenes, name):
    """4-wide BVH (csrc/bvh4_build.cpp): every triangle in exactly one leaf, stored child boxes strictly contain their
    subtrees, empty slots are the unhittable box lo = hi = +inf, leaves hold at most 4 ascending triangle ids."""
    s = scenes(name)
    r = host_tracer(s)
    boxes, children, counts, order = r.bvh4_export()
    st = r.bvh4_stats()
    n_tri = s.vertices.shape[0]
    assert sorted(order.tolist()) == list(range(n_tri))
    assert st["max_leaf"] <= 4 and 3 * st["depth"] + 5 <= 64
    assert st["nodes"] < r.bvh_stats()["nodes"]  # the collapse removes inner nodes
    verts = s.vertices.reshape(n_tri, 3, 3)
    seen = np.zeros(n_tri, bool)
    visited = np.zeros(st["nodes"], bool)

    def subtree_box(node):
        assert not visited[node]
        visited[node] = True
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        for k in range(4):
            c, cnt = int(children[node, k]), int(counts[node, k])
            if c < 0 and cnt == 0:
                assert np.isposinf(boxes[node, k]).all()
                continue
            if c >= 0:
                lo, hi = subtree_box(c)
            else:
                tri = order[~c:~c + cnt]
                assert not seen[tri].any()
                seen[tri] = True
                assert (np.diff(tri.astype(np.int64)) > 0).all()
                lo, hi = verts[tri].reshape(-1, 3).min(0), verts[tri].reshape(-1, 3).max(0)
            assert (boxes[node, k, 0] < lo).all() and (boxes[node, k, 1] > hi).all()
            lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
        return lo_all, hi_all

    import sys
    sys.setrecursionlimit(10000)
    subtree_box(0)
    assert seen.all() and visited.all()


@pytest.mark.parametrize("name", ["4boxes", "ico2", "thai2"])
def test_cwbvh_invariants(scenes, name):
    """Compressed 8-wide BVH (csrc/cwbvh_build.cpp): every triangle sits in exactly one leaf child, every quantised child
    box (decoded with the f32 expression the traversal uses, plane = p + q * 2^e) contains its subtree, inner
    children are numbered consecutively in slot order."""
    s = scenes(name)
    r = host_tracer(s)
    words, order = r.cwbvh_export()
    st = r.cwbvh_stats()
    n_tri = s.vertices.shape[0]
    assert sorted(order.tolist()) == list(range(n_tri))
    assert st["nodes"] == words.shape[0] and st["depth"] + 3 <= 32
    verts = s.vertices.reshape(n_tri, 3, 3)
    seen = np.zeros(n_tri, bool)
    visited = np.zeros(st["nodes"], bool)

    def bytes_of(w):
        return [(int(w) >> (8 * k)) & 255 for k in range(4)]

    def decode(node):
        w = words[node]
        p = w[0, :3].copy().view(np.float32)
        e = bytes_of(w[0, 3])
        scale = np.array([np.float32(2.0) ** np.float32(e[a] - 127) for a in range(3)], np.float32)
        imask = e[3]
        meta = bytes_of(w[1, 2]) + bytes_of(w[1, 3])
        q = [bytes_of(w[2, 0]) + bytes_of(w[2, 1]), bytes_of(w[2, 2]) + bytes_of(w[2, 3]), bytes_of(w[3, 0]) + bytes_of(w[3, 1]),
             bytes_of(w[3, 2]) + bytes_of(w[3, 3]), bytes_of(w[4, 0]) + bytes_of(w[4, 1]), bytes_of(w[4, 2]) + bytes_of(w[4, 3])]
        lo = np.array([[p[a] + np.float32(q[a][sl]) * scale[a] for a in range(3)] for sl in range(8)], np.float32)
        hi = np.array([[p[a] + np.float32(q[3 + a][sl]) * scale[a] for a in range(3)] for sl in range(8)], np.float32)
        return int(w[1, 0]), int(w[1, 1]), imask, meta, lo, hi

    def walk(node):
        assert not visited[node]
        visited[node] = True
        child_base, tri_base, imask, meta, lo, hi = decode(node)
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        rel = 0
        for sl in range(8):
            m = meta[sl]
            if m == 0:
                assert not (imask >> sl) & 1
                continue
            if (imask >> sl) & 1:
                assert m == (1 << 5) | (24 + sl)
                clo, chi = walk(child_base + rel)
                rel += 1
            else:
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]
                off = m & 31
                assert off + cnt <= 24
                tri = order[tri_base + off: tri_base + off + cnt]
                assert not seen[tri].any()
                seen[tri] = True
                clo, chi = verts[tri].reshape(-1, 3).min(0), verts[tri].reshape(-1, 3).max(0)
            assert (lo[sl] <= clo).all() and (hi[sl] >= chi).all(), (node, sl)
            lo_all, hi_all = np.minimum(lo_all, clo), np.maximum(hi_all, chi)
        return lo_all, hi_all

    import sys
    sys.setrecursionlimit(10000)
    walk(0)
    assert seen.all() and visited.all()


# ---- camera (camera.rs) ------------------------------------------------------------------------------------------------


def test_camera_matches_oracle_through_interactive_moves(scenes):
    s = scenes("thai2")
    r, o = host_tracer(s, 1920, 1080), Oracle(s, 1920, 1080)
    assert np.array_equal(r.camera.matrices().view(np.uint32), o.camera().view(np.uint32))
    moves = [("move_rel", (0.1, 0.0, 0.0)), ("add_y_angle", (0.01,)), ("add_x_angle", (-0.01,)), ("move_rel", (0.0, -0.1, 0.1)),
             ("add_y_angle", (0.01,)), ("add_x_angle", (0.5,))]  # the key handlers of raytracer/src/main.rs:124-162
    for fn, args in moves:
        getattr(r.camera, fn)(*args)
        getattr(o, fn)(*args)
        assert np.array_equal(r.camera.matrices().view(np.uint32), o.camera().view(np.uint32)), fn
    r.camera.set_state(0.25, -0.5, (1.0, 2.0, 3.0))
    o2 = Oracle(s, 1920, 1080)
    o2.add_x_angle(0.25), o2.add_y_angle(-0.5), o2.move_rel(1.0, 2.0, 3.0)
    assert np.array_equal(r.camera.matrices(), o2.camera())


@pytest.mark.parametrize("w,h", [(1920, 1080), (1024, 768), (300, 500)])
def test_camera_plane_matrix_inverts_get_ray(scenes, w, h):
    """The map behind the perspective grid of the camera rays (rt_get_camera_plane_matrix; csrc/pgrid_build.cu bins triangles with it): a point
    origin + t * dir of the ORACLE's camera ray (camera.rs:80-90 restated in oracle/rt_oracle.cpp) of pixel idx, with u = idx % W and
    v = idx / H (mod.rs:96) and any sub-pixel offsets, lands inside the pixel's unit square of the sample plane, and its Z is t. Checked
    for the start view and through moves and turns."""
    s = scenes("thai2")
    r, o = host_tracer(s, w, h), Oracle(s, w, h)
    rng = np.random.default_rng(5)
    moves = [None, ("move_rel", (0.3, -0.2, 1.5)), ("add_y_angle", (0.7,)), ("add_x_angle", (-0.4,)), ("move_rel", (-2.0, 0.5, 0.1)), ("add_y_angle", (2.9,))]
    for mv in moves:
        if mv:
            getattr(r.camera, mv[0])(*mv[1])
            getattr(o, mv[0])(*mv[1])
        A, origin = r.camera.plane_matrix()
        for idx in np.concatenate([[0, w * h - 1, w - 1, w * (h - 1)], rng.integers(0, w * h, 60)]):
            u, v = int(idx) % w, int(idx) // h
            xi1, xi2 = (0.5, 0.5) if idx % 3 == 0 else rng.random(2) * 0.998 + 0.001
            ray = o.get_ray(u, v, float(xi1), float(xi2)).astype(np.float64)
            pos, d = ray[:3], ray[3:]
            assert np.allclose(pos, origin, rtol=0, atol=1e-6 * (1 + np.abs(origin).max()))
            for t in (0.25, 3.0, 400.0):
                X, Y, Z = A @ (pos + t * d - origin)
                assert abs(Z - t) <= 2e-5 * t, (mv, idx, t, Z)
                # the f32 ray against the binary64 map: within a few thousandths of a pixel of the sample position
                assert abs(X / Z - (u + xi1)) < 5e-3 and abs(Y / Z - (v + xi2)) < 5e-3, (mv, idx, X / Z - u, Y / Z - v)
    r.close()


def test_render_calls_fail_loudly_without_a_device(scenes):
    """No CPU fallback: a host-side handle refuses to render."""
    r = host_tracer(scenes("4boxes"))
    for call in (r.trace_frame_additive, r.get_tonemapped_pixels, r.film.clear, lambda: r.trace_rows(0, 1)):
        with pytest.raises(rt.RtError) as e:
            call()
        assert e.value.code == -3


def test_rt_create_rejects_malformed_scene_descriptions():
    """The C ABI takes raw pointers: a non-zero count with a null array, a texture without texels or a triangle that
    names a geometry that does not exist must come back as RT_ERR_INVALID with a message, not as a crash."""
    import ctypes as C

    from raytracer_rs_b200 import api

    lib = rt.lib()
    cfg = rt.Config(16, 16, device=api.DEVICE_NONE).to_c()
    verts = (C.c_float * 9)(0, 0, 0, 1, 0, 0, 0, 1, 0)
    geom = (C.c_uint32 * 1)(0)
    mats = (api.CMaterial * 1)()
    lights = (api.CLight * 1)()
    texels = (C.c_float * 3)(0.5, 0.5, 0.5)

    def desc():
        d = api.CSceneDesc()
        d.num_triangles, d.vertices, d.tri_geom = 1, verts, geom
        d.num_geometries, d.materials = 1, mats
        d.num_lights, d.lights = 1, lights
        d.camera_orientation[:] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
        d.camera_fov_deg = 40.0
        return d

    def create(d):
        h, err = C.c_void_p(), C.create_string_buffer(256)
        rc = lib.rt_create(C.byref(d), C.byref(cfg), C.byref(h), err, len(err))
        if rc == 0:
            lib.rt_destroy(h)
        return rc, err.value.decode()

    assert create(desc())[0] == 0
    for field in ("vertices", "tri_geom", "materials", "lights"):
        d = desc()
        setattr(d, field, None)
        rc, msg = create(d)
        assert rc == -1 and "null array" in msg, field
    d = desc()
    d.num_textures, d.textures = 1, None
    assert create(d)[0] == -1
    d = desc()
    tex = (api.CTexture * 1)()
    tex[0].width, tex[0].height = 1, 1  # rgb stays null
    d.num_textures, d.textures = 1, tex
    rc, msg = create(d)
    assert rc == -1 and "texels" in msg
    tex[0].rgb = texels
    assert create(d)[0] == 0
    d = desc()
    geom_bad = (C.c_uint32 * 1)(3)
    d.tri_geom = geom_bad
    rc, msg = create(d)
    assert rc == -1 and "out of range" in msg


# ---- stats.rs / timing crate ---------------------------------------------------------------------------------------------


def test_stats_strings():
    st = rt.Stats()
    time.sleep(0.02)
    line = st.stats(96000)
    m = re.fullmatch(r"fps: ([0-9.]+)  primary rays/s: (\d+)", line)  # stats.rs:31
    assert m and 5 < float(m.group(1)) < 60 and 96000 * 5 < int(m.group(2)) < 96000 * 60
    time.sleep(0.01)
    st.stats(96000)
    assert re.fullmatch(r"mean fps: [0-9.]+  mean primary rays/s: [0-9.]+", st.mean_stats())  # stats.rs:35-39


def test_benchmark_report_shape():
    """timing/src/lib.rs:95-109 and the README example: sorted by total duration, '{name} total: {}ms, mean: {}ms, samples: {}'"""
    bm = rt.BenchMark()
    bm.start("foo")
    time.sleep(0.02)
    bm.stop("foo")
    for _ in range(3):
        with bm.time_scope("bar"):
            time.sleep(0.002)
    lines = str(bm).strip().split("\n")
    assert [l.split(" ")[0] for l in lines] == ["foo", "bar"]
    m = re.fullmatch(r"bar total: ([0-9.]+)ms, mean: ([0-9.]+)ms, samples: 3", lines[1])
    assert m and abs(float(m.group(1)) / 3 - float(m.group(2))) < 0.01
    with pytest.raises(KeyError):
        bm.stop("never started")


# ---- bench.py: the roofline's numerator ---------------------------------------------------------------------------


@pytest.mark.parametrize("workload", ["thai2_1080p", "ico2_1024x768", "4boxes_1080p", "ico3_tex_1080p"])
def test_bench_reference_work_constants_are_the_oracles_counters(scenes, workload):
    """bench.py's `roofline.achieved` = algorithmic bytes of the REFERENCE algorithm per pinned frame / kernel time, with
    the bytes computed from constants (rays, cube tests, triangle tests per frame). They must be what the oracle counts
    when it renders that frame at full size (24 B per cube test + 36 B per triangle test + 4 B per pixel)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(os.path.dirname(DATA), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fname, w, h, spp = bench.WORKLOADS[workload]
    assert spp == 1
    o = Oracle(scenes(fname.split(".")[0]), w, h)
    o.configure(recursions=0, jitter=JITTER_FIXED)
    o.trace_rows(0, h, 1, threads=0)
    c = o.counters()
    got = (c["rays"]["primary"], c["rays"]["shadow"], c["cube_tests"]["primary"] + c["cube_tests"]["shadow"],
           c["tri_tests"]["primary"] + c["tri_tests"]["shadow"])
    assert got == bench.REFERENCE_WORK[workload]
    assert bench.algorithmic_bytes(workload) == 24 * got[2] + 36 * got[3] + 4 * got[0]


def test_tuning_knobs_accept_their_documented_ranges(scenes):
    """rt_set_tuning (include/rt_b200.h): every knob takes its documented values and rejects the first value outside."""
    r = host_tracer(scenes("4boxes"))
    ranges = {0: (0, 2), 1: (0, 1), 2: (1, 32), 3: (0, 8), 4: (0, 32), 5: (0, 2), 6: (0, 1), 7: (0, 64), 8: (1, 64), 9: (0, 100), 10: (0, 1),
              12: (0, 3), 13: (0, 1), 14: (1, 32), 15: (0, 32), 16: (3, 5), 17: (0, 8), 18: (0, 1), 19: (0, 1), 20: (0, 3), 21: (0, 1024), 22: (2, 5), 23: (0, 1000), 24: (6, 9)}
    for key, (lo, hi) in ranges.items():
        r.set_tuning(key, lo)
        r.set_tuning(key, hi)
        for bad in (lo - 1, hi + 1):
            with pytest.raises(rt.RtError):
                r.set_tuning(key, bad)
    with pytest.raises(rt.RtError):
        r.set_tuning(11, 0)
