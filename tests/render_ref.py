"""Second, independent restatement of the reference's render path in numpy float32 — TEST INFRASTRUCTURE ONLY.

Written from the Rust sources (not from oracle/rt_oracle.cpp) so that a transcription error in either restatement shows
up as a difference between them (tests/test_oracle_golden.py::test_oracle_matches_independent_numpy_restatement). It is
slow (Python loops over rays) and only meant for frames of a few thousand pixels. Every numpy operation below is one
IEEE binary32 operation per element (separate ufunc calls are never contracted into FMAs), in the order the Rust
expressions evaluate.

Restated: octree build with the SAT triangle/cube test (oct_tree_intersector.rs:66-146, 315-458), octree traversal
(:148-199, 262-291, 348-371), Moller-Trumbore `intersect_late_out` (intersect.rs:62-98), camera (scene/camera.rs:22-98,
vecmath.rs:87-135, 200-211, 237-313), `compute_radiance` with RECURSIONS = 0, `calc_normal`, `shade`
(raytracer/mod.rs:132-261), nearest texel lookup (scene/texture.rs:22-28), film / tonemap / pack (film.rs:12-48,
tonemap.rs:4-10, color.rs:85-95), the pixel -> ray mapping of `trace_frame_additive` (mod.rs:88-112) with the sub-pixel
offset pinned to 0.5 (the reference draws it from an OS-seeded RNG).
Optional (configure): bounce rays — `compute_radiance` with RECURSIONS > 0 and `randomize_reflection_ray`
(mod.rs:132-196, sample_generator.rs:26-34) — and jittered sub-pixel offsets, both with the deterministic stand-in for
the reference's OS-seeded RNG that DESIGN.md section 3 defines (counter-based hash4; the table of unit vectors is data
and is passed in).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

# f32::powf is the C library's powf (Rust lowers it to the llvm.pow.f32 intrinsic, which calls libm); numpy's own float32
# power loop is a SIMD approximation that differs from it by an ulp now and then
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]

F = np.float32
F_MAX = np.finfo(np.float32).max
F_EPS = np.finfo(np.float32).eps
NO_HIT = 0xFFFFFFFF


def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]  # vecmath.rs:74-76


def _cross(a, b):  # vecmath.rs:79-85
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _normalized(v):  # vecmath.rs:23-26
    ln = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return np.array([v[0] / ln, v[1] / ln, v[2] / ln], F)


def _matmul(a, b):  # vecmath.rs:237-313: row-major 4x4, every element summed left to right
    out = np.zeros(16, F)
    for r in range(4):
        for c in range(4):
            out[4 * r + c] = a[4 * r] * b[c] + a[4 * r + 1] * b[4 + c] + a[4 * r + 2] * b[8 + c] + a[4 * r + 3] * b[12 + c]
    return out


def _mat_vec4(m, v):  # vecmath.rs:200-211
    return np.array([v[0] * m[k] + v[1] * m[4 + k] + v[2] * m[8 + k] + v[3] * m[12 + k] for k in range(4)], F)


def _ident():
    m = np.zeros(16, F)
    m[0] = m[5] = m[10] = m[15] = F(1)
    return m


def _mix32(h):
    h ^= h >> 16
    h = (h * 0x7FEB352D) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 0x846CA68B) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def hash4(a, b, c, d):
    h = _mix32((a + 0x9E3779B9) & 0xFFFFFFFF)
    h = _mix32(h ^ ((b + 0x85EBCA6B) & 0xFFFFFFFF))
    h = _mix32(h ^ ((c + 0xC2B2AE35) & 0xFFFFFFFF))
    return _mix32(h ^ ((d + 0x27D4EB2F) & 0xFFFFFFFF))


def _u01(h):
    return F(h >> 8) * F(2.0 ** -24)


class RenderRef:
    def __init__(self, scene, width: int, height: int, triangles_per_leaf: int = 70):
        self.recursions, self.sub_spread, self.jitter_hashed, self.seed, self.table = 0, 1, False, 0, None
        self.bounce_rays = 0
        self.W, self.H = width, height
        v = np.asarray(scene.vertices, F).reshape(-1, 3, 3)
        self.V0, self.V1, self.V2 = v[:, 0], v[:, 1], v[:, 2]
        self.E1, self.E2 = self.V1 - self.V0, self.V2 - self.V0  # intersect.rs:66-67 (same values for every ray)
        self.tri_geom = np.asarray(scene.tri_geom)
        self.materials, self.lights, self.textures = scene.materials, scene.lights, scene.textures
        self._camera(np.asarray(scene.camera_orientation, F), F(scene.camera_fov_deg))
        self._build_octree(triangles_per_leaf)
        n = width * height
        self.film_sum, self.film_sq, self.film_n = np.zeros((n, 3), F), np.zeros((n, 3), F), np.zeros(n, np.uint32)
        self.ids = np.full(n, NO_HIT, np.uint32)
        self.shadow_rays = 0

    def configure(self, recursions=0, sub_spread=1, jitter_hashed=False, seed=0, sample_table=None):
        self.recursions, self.sub_spread, self.jitter_hashed, self.seed = recursions, sub_spread, jitter_hashed, seed
        self.table = None if sample_table is None else np.asarray(sample_table, F)

    # ---- scene/camera.rs -------------------------------------------------------------------------------------------
    def _camera(self, orientation, fov_deg):
        rot = orientation.copy()  # camera.rs:28-39
        rot[3] = rot[7] = rot[11] = rot[12] = rot[13] = rot[14] = F(0)
        rot[15] = F(1)
        fov = fov_deg * F(math.pi) / F(180.0)
        half = F(0.5) * fov
        # f32::tan: the correctly rounded value (what glibc's tanf returns for these arguments)
        self.max_x = self.max_y = F(1.0) * F(math.tan(float(half)))
        zero = F(0)
        rx, ry = _ident(), _ident()  # vecmath.rs:112-127 with x_angle = y_angle = 0
        rx[5], rx[6], rx[9], rx[10] = np.cos(zero), -np.sin(zero), np.sin(zero), np.cos(zero)
        ry[0], ry[2], ry[8], ry[10] = np.cos(zero), np.sin(zero), -np.sin(zero), np.cos(zero)
        self.rotation = _matmul(_matmul(rx, ry), rot)  # camera.rs:92-98
        self.orientation = _matmul(_matmul(self.rotation, _ident()), orientation)  # translate(0, 0, 0) = identity

    def get_ray(self, u: int, v: int, xi1=F(0.5), xi2=F(0.5)):  # camera.rs:80-90, random_range(0.0..1.0) = xi
        dir_x = -self.max_x + F(2.0) * self.max_x * ((F(u) + xi1) / F(self.W))
        dir_y = -self.max_y + F(2.0) * self.max_y * ((F(v) + xi2) / F(self.H))
        d = _mat_vec4(self.rotation, np.array([dir_x, -dir_y, 1.0, 1.0], F))
        p = _mat_vec4(self.orientation, np.array([0.0, 0.0, 0.0, 1.0], F))
        return p[:3].copy(), d[:3].copy()

    # ---- oct_tree_intersector.rs: build ----------------------------------------------------------------------------
    def _build_octree(self, tpl):
        allv = np.concatenate([self.V0, self.V1, self.V2])
        cmin, cmax = np.full(3, F_MAX, F), np.full(3, -F_MAX, F)  # calc_extents :315-330
        for a in range(3):
            cmin[a], cmax[a] = np.fmin(cmin[a], allv[:, a].min()), np.fmax(cmax[a], allv[:, a].max())
        self.cubes = [(cmin, cmax)]
        self.nodes = [("leaf", np.arange(len(self.V0)))]  # all_triangle_indices :332-342: geometry order, triangle order
        self._split(0, tpl, 0)

    def _split(self, idx, tpl, level):  # split_node :91-146
        kind, tris = self.nodes[idx]
        if kind != "leaf" or len(tris) <= tpl or level > 8:
            return
        cmin, cmax = self.cubes[idx]
        mid = F(0.5) * (cmax + cmin)  # generate_child_cubes :265-313
        new = []
        for i in range(8):
            lo = np.array([mid[a] if (i >> a) & 1 else cmin[a] for a in range(3)], F)
            hi = np.array([cmax[a] if (i >> a) & 1 else mid[a] for a in range(3)], F)
            self.cubes.append((lo, hi))
            new.append(("leaf", tris[self._sat(lo, hi, tris)]))
        first = len(self.nodes)
        self.nodes[idx] = ("node", list(range(first, first + 8)))
        self.nodes.extend(new)
        for child in range(first, first + 8):
            self._split(child, tpl, level + 1)

    @staticmethod
    def _project(points, axis):  # project_points_on_axis :460-469; points (..., P, 3), axis (..., 3) -> min, max over P
        val = axis[..., None, 0] * points[..., 0] + axis[..., None, 1] * points[..., 1] + axis[..., None, 2] * points[..., 2]
        lo, hi = np.full(val.shape[:-1], F_MAX, F), np.full(val.shape[:-1], -F_MAX, F)
        for k in range(val.shape[-1]):
            lo, hi = np.fmin(lo, val[..., k]), np.fmax(hi, val[..., k])
        return lo, hi

    def _sat(self, lo, hi, tris):  # triangle_cube_intersection :393-458, vectorised over the candidate triangles
        if len(tris) == 0:
            return np.zeros(0, bool)
        tv = np.stack([self.V0[tris], self.V1[tris], self.V2[tris]], axis=1)  # (T, 3 vertices, 3)
        keep = np.ones(len(tris), bool)
        unit = np.eye(3, dtype=F)
        for a in range(3):
            tmin, tmax = self._project(tv, np.broadcast_to(unit[a], (len(tris), 3)))
            keep &= ~((tmax < lo[a]) | (tmin > hi[a]))
        cv = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [lo[0], hi[1], lo[2]], [lo[0], lo[1], hi[2]],
                       [lo[0], hi[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], lo[2]], [hi[0], hi[1], hi[2]]], F)
        cvb = np.broadcast_to(cv, (len(tris), 8, 3))
        e1, e2 = tv[:, 0] - tv[:, 1], tv[:, 1] - tv[:, 2]
        nrm = _cross(e1, e2)
        off = _dot(nrm, tv[:, 0])
        cmin, cmax = self._project(cvb, nrm)
        keep &= ~((cmax < off) | (cmin > off))
        e3 = tv[:, 2] - tv[:, 0]
        for e in (e1, e2, e3):
            for a in range(3):
                axis = _cross(e, np.broadcast_to(unit[a], e.shape))
                cmin, cmax = self._project(cvb, axis)
                tmin, tmax = self._project(tv, axis)
                keep &= ~((cmax < tmin) | (cmin > tmax))
        return keep

    def octree_export(self):
        """Same flat form as Oracle.octree_export: cubes, first_child (-1 = leaf), leaf_offset, leaf_tris."""
        n = len(self.nodes)
        cubes = np.array([np.concatenate(c) for c in self.cubes], F)
        first_child = np.array([nd[1][0] if nd[0] == "node" else -1 for nd in self.nodes], np.int32)
        counts = [len(nd[1]) if nd[0] == "leaf" else 0 for nd in self.nodes]
        leaf_offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
        leaf_tris = np.concatenate([nd[1] for nd in self.nodes if nd[0] == "leaf"]).astype(np.uint32) if n else np.zeros(0, np.uint32)
        return cubes, first_child, leaf_offset, leaf_tris

    # ---- oct_tree_intersector.rs: traversal ------------------------------------------------------------------------
    def intersect_ray(self, pos, d):  # :252-260
        with np.errstate(all="ignore"):
            inv = F(1.0) / d
            return self._node(pos, d, inv, 0)

    def _leaf_closest(self, pos, d, tris):  # intersect_leaf_triangles :262-291 + intersect_late_out intersect.rs:62-98
        if len(tris) == 0:
            return None
        e1, e2 = self.E1[tris], self.E2[tris]
        db = np.broadcast_to(d, e2.shape)
        pvec = _cross(db, e2)
        det = _dot(e1, pvec)
        inv_det = F(1.0) / det
        tvec = pos - self.V0[tris]
        u = _dot(tvec, pvec) * inv_det
        qvec = _cross(tvec, e1)
        v = _dot(db, qvec) * inv_det
        t = _dot(e2, qvec) * inv_det
        miss = (np.abs(det) < F_EPS) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1) | (t < 0)
        best = None
        for k in np.nonzero(~miss)[0]:  # in leaf order; strictly smaller t replaces
            if best is None or t[k] < best[0]:
                best = (t[k], u[k], v[k], int(tris[k]))
        return best

    def _node(self, pos, d, inv, idx):  # intersect_node :148-199
        kind, payload = self.nodes[idx]
        if kind == "leaf":
            hit = self._leaf_closest(pos, d, payload)
            if hit is None:
                return None
            hp = pos + d * hit[0]
            lo, hi = self.cubes[idx]
            inside = not (hp[0] < lo[0] or hp[0] > hi[0] or hp[1] < lo[1] or hp[1] > hi[1] or hp[2] < lo[2] or hp[2] > hi[2])
            return hit if inside else None
        lo = np.array([self.cubes[c][0] for c in payload], F)
        hi = np.array([self.cubes[c][1] for c in payload], F)
        t1, t2 = (lo - pos) * inv, (hi - pos) * inv  # intersect_cube_inverse_ray :348-371; f32::min/max ignore a NaN operand
        tmin, tmax = np.fmin(t1[:, 0], t2[:, 0]), np.fmax(t1[:, 0], t2[:, 0])
        for a in (1, 2):
            tmin = np.fmax(tmin, np.fmin(t1[:, a], t2[:, a]))
            tmax = np.fmin(tmax, np.fmax(t1[:, a], t2[:, a]))
        entered = np.nonzero((tmax >= tmin) & (tmax > 0))[0]
        for k in entered[np.argsort(tmin[entered], kind="stable")]:  # sort_by(partial_cmp) is stable
            hit = self._node(pos, d, inv, payload[k])
            if hit is not None:
                return hit
        return None

    # ---- raytracer/mod.rs ------------------------------------------------------------------------------------------
    def _radiance(self, pos, d, hit, rec, pixel, sample, path):  # compute_radiance :132-176
        tri = hit[3]
        normal = _normalized(_cross((self.V1[tri] - self.V0[tri])[None], (self.V2[tri] - self.V0[tri])[None])[0])  # calc_normal :198-205
        radiance = self._shade(pos, d, hit, normal)
        if rec < 1:
            return radiance
        n = self.sub_spread * rec
        total = np.zeros(3, F)
        for k in range(n):
            sub_path = (path * 31 + k + 1) & 0xFFFFFFFF
            # randomize_reflection_ray :178-196; normalized_vec_pseudo = table[random_range(0..NUM_SAMPLES - 1)], then
            # normalized_vec_lookup = table[(idx + 1) % SAMPLE_MAX] until the direction leaves the surface
            idx = (hash4(self.seed ^ 0xB0C0FFEE, pixel, sample, sub_path) * 65535) >> 32
            rd = self.table[idx]
            while _dot(rd, normal) <= 0:
                idx = (idx + 1) % 65535
                rd = self.table[idx]
            hp = pos + hit[0] * d
            hp = hp + F(0.00001) * rd
            self.bounce_rays += 1
            sub_hit = self.intersect_ray(hp, rd)
            x = np.zeros(3, F) if sub_hit is None else self._radiance(hp, rd, sub_hit, rec - 1, pixel, sample, sub_path)
            total = total + x
        return radiance + total * (F(1.0) / F(n))

    def _shade(self, pos, d, hit, normal):  # shade :207-261
        t, u, v, tri = hit
        accum = np.zeros(3, F)
        hit_point = pos + t * d
        for lpos, lcol in self.lights:
            to_light = np.asarray(lpos, F) - hit_point
            ndl = _dot(normal, _normalized(to_light))
            if ndl < 0:
                continue
            self.shadow_rays += 1
            blocker = self.intersect_ray(hit_point + to_light * F(0.01), to_light)
            if blocker is not None and blocker[0] > F(0.01) and blocker[0] < F(1.0):
                continue
            kind, rgb, tex_id = self.materials[int(self.tri_geom[tri])]
            if int(kind) == 1:  # Diffuse::TextureId, texture.rs:22-28: nearest texel, `as usize` truncates
                tw, th, texels = self.textures[int(tex_id)]
                x, y = int(u * F(tw)), int(v * F(th))
                # the reference panics when the index leaves the texture (v == 1.0); oracle and product read the last texel
                diffuse = np.asarray(texels, F).reshape(-1, 3)[min(y * tw + x, tw * th - 1)]
            else:
                diffuse = np.array(rgb, F)
            view = _normalized(d)
            reflected = F(2.0) * ndl * normal - _normalized(to_light)
            spec = F(_libm.powf(float(_dot(view, reflected)), 32.0))  # f32::powf
            accum = accum + (diffuse * ndl + F(1.0) * spec) * np.asarray(lcol, F)
        return accum

    def trace_rows(self, first_row: int, n_rows: int):
        """mod.rs:88-112 for `n_rows` rows starting at `first_row`, one sample per pixel."""
        with np.errstate(all="ignore"):
            for r in range(n_rows):
                row = (first_row + r) % self.H
                for i in range(self.W):
                    idx = row * self.W + i
                    sample = int(self.film_n[idx])
                    xi = (_u01(hash4(self.seed, idx, sample, 0)), _u01(hash4(self.seed, idx, sample, 1))) if self.jitter_hashed else (F(0.5), F(0.5))
                    pos, d = self.get_ray(idx % self.W, idx // self.H, *xi)  # [sic] mod.rs:96
                    hit = self.intersect_ray(pos, d)
                    color = np.zeros(3, F) if hit is None else self._radiance(pos, d, hit, self.recursions, idx, sample, 0)
                    self.ids[idx] = NO_HIT if hit is None else hit[3]
                    self.film_sum[idx] = self.film_sum[idx] + color  # film.rs:20-24
                    self.film_sq[idx] = self.film_sq[idx] + color * color
                    self.film_n[idx] += 1

    def get_tonemapped_pixels(self):  # mod.rs:120-128, film.rs:43-48, tonemap.rs:4-10, color.rs:85-95
        with np.errstate(all="ignore"):
            mean = self.film_sum * (F(1.0) / self.film_n.astype(F))[:, None]
            mapped = mean / (F(1.0) + mean)
            # f32::min / f32::max return the other operand for a NaN: NaN.min(1.0) = 1.0 (never-sampled pixels are white)
            ch = (np.fmax(np.fmin(mapped, F(1.0)), F(0.0)) * F(255.0)).astype(np.uint8).astype(np.uint32)
        return ch[:, 2] | (ch[:, 1] << 8) | (ch[:, 0] << 16) | np.uint32(0xFF000000)
