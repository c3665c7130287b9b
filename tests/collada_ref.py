"""Independent numpy/f32 restatement of the reference's Collada -> Scene flattening.

TEST INFRASTRUCTURE (not product code). It exists so that the product's C++ loader
(raytracer_rs_b200/csrc/collada_loader.cpp, reached through `rt_scene_load_file`) can be checked
bit for bit against a second implementation, and so that CPU-only tests can feed the oracle.

Follows /root/reference/raytracer_lib/src/scene/loaders/colladaloader.rs:
  Collada::parse :59-135, to_scene_flatten :137-273, to_cameras :276-319, to_lights :321-349,
  to_effects :351-467, to_images :469-486, to_materials :488-505, to_visual_scenes :507-548,
  convert_geometry :561-601, and colladaloader/collada_types.rs:76-90 (matrix conversion),
  scene/texture.rs:34-49 (texels = byte / 256), vecmath.rs:200-211,237-313 (operation order).
Decimal strings are converted with correctly rounded decimal->binary32 conversion (the reference's
`parseval` crate is not vendored; SURVEY.md section 8c: parity unpinned for parsed values).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np

F = np.float32
NS = "{http://www.collada.org/2005/11/COLLADASchema}"


def parse_f32(tok: str) -> np.float32:
    """Correctly rounded decimal -> binary32 (no double rounding)."""
    d = float(tok)
    f = np.float32(d)
    # double rounding can only go wrong when the double lies exactly on a binary32 midpoint
    if float(f) != d:
        lo = np.nextafter(f, np.float32(-np.inf)) if float(f) > d else f
        hi = np.nextafter(lo, np.float32(np.inf))
        mid = (Fraction(float(lo)) + Fraction(float(hi))) / 2
        if Fraction(d) == mid:  # exact tie in double: decide with the exact decimal value
            exact = Fraction(tok)
            if exact > mid:
                return hi
            if exact < mid:
                return lo
    return f


def array_f32(text: str) -> np.ndarray:
    return np.array([parse_f32(t) for t in text.split()], dtype=F)


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """vecmath.rs:237-313, f32, 4-term sums left to right."""
    r = np.zeros(16, dtype=F)
    for i in range(4):
        for j in range(4):
            s = F(a[4 * i + 0] * b[j])
            s = F(s + F(a[4 * i + 1] * b[4 + j]))
            s = F(s + F(a[4 * i + 2] * b[8 + j]))
            s = F(s + F(a[4 * i + 3] * b[12 + j]))
            r[4 * i + j] = s
    return r


def mat_transpose(m: np.ndarray) -> np.ndarray:
    return m.reshape(4, 4).T.reshape(16).copy()


SWAP_YZ = np.array([1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1], dtype=F)
REFLECT_Z = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1], dtype=F)


def collada_to_vecmath(elems: np.ndarray) -> np.ndarray:
    """collada_types.rs:76-90: reflect_z * transpose * swap_yz."""
    return mat_mul(mat_mul(REFLECT_Z, mat_transpose(elems[:16].astype(F))), SWAP_YZ)


def transform_points(m: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """row-vector * matrix with w = 1 (vecmath.rs:200-211), vectorised, f32 per operation."""
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    w = np.ones_like(x)
    out = np.empty_like(pts)
    for c in range(3):
        s = x * m[c]
        s = s + y * m[4 + c]
        s = s + z * m[8 + c]
        s = s + w * m[12 + c]
        out[:, c] = s
    return out


@dataclass
class FlatScene:
    vertices: np.ndarray  # [T, 9] f32
    tri_geom: np.ndarray  # [T] u32
    materials: list  # per geometry: (kind, (r,g,b), texture_id)
    lights: list  # (pos[3], color[3]) f32
    textures: list  # (w, h, rgb f32 [h*w*3])
    camera_orientation: np.ndarray  # [16] f32 (vecmath matrix of cameras[0])
    camera_fov_deg: np.float32 = F(0)
    geometry_ids: list = field(default_factory=list)


def load_texture(path: str):
    from PIL import Image

    im = Image.open(path).convert("RGB")  # image::open(..).to_rgb8(); gAMA/sRGB chunks ignored
    a = np.asarray(im, dtype=np.uint8)
    h, w, _ = a.shape
    return w, h, (a.astype(F) / F(256.0)).reshape(-1)


def load_collada(path: str) -> FlatScene:
    data_dir = os.path.dirname(path)
    root = ET.parse(path).getroot()
    if root.tag != NS + "COLLADA":
        raise ValueError("Not a collada doc")

    def lib(name):
        e = root.find(NS + name)
        if e is None:
            raise ValueError(f"{name} parsing error")
        return e

    cameras = []
    for cam in lib("library_cameras"):
        persp = cam.find(f"{NS}optics/{NS}technique_common/{NS}perspective")
        cameras.append((cam.get("id"), array_f32(persp.find(NS + "xfov").text)[0]))
    lights = []
    for li in lib("library_lights"):
        col = array_f32(li.find(f"{NS}technique_common/{NS}point/{NS}color").text)
        lights.append((li.get("id"), col[:3]))
    effects = {}
    for eff in lib("library_effects"):
        prof = eff.find(NS + "profile_COMMON")
        lam = prof.find(f"{NS}technique/{NS}lambert")
        diff = lam.find(NS + "diffuse")
        col = diff.find(NS + "color")
        if col is not None:
            effects[eff.get("id")] = ("color", array_f32(col.text)[:3])
        else:
            tex = diff.find(NS + "texture")
            sampler = tex.get("texture")
            surface = None
            for np_ in prof.findall(NS + "newparam"):
                if np_.get("sid") == sampler:
                    surface = np_.find(f"{NS}sampler2D/{NS}source").text
            image_id = None
            for np_ in prof.findall(NS + "newparam"):
                if np_.get("sid") == surface:
                    image_id = np_.find(f"{NS}surface/{NS}init_from").text
            effects[eff.get("id")] = ("texture", image_id)
    images = [(im.get("id"), im.find(NS + "init_from").text) for im in lib("library_images")]
    materials = {m.get("id"): m.find(NS + "instance_effect").get("url")[1:] for m in lib("library_materials")}
    geometries = []
    for g in lib("library_geometries"):
        gid = g.get("id")
        mesh = g.find(NS + "mesh")
        pos = None
        for src in mesh.findall(NS + "source"):
            if src.get("id") == f"{gid}-positions":
                pos = array_f32(src.find(NS + "float_array").text)
        tris = mesh.find(NS + "triangles")
        idx = np.array(tris.find(NS + "p").text.split(), dtype=np.uint32)
        idx = idx[: (len(idx) // 3) * 3].reshape(-1, 3)[:, 0]  # (pos, normal, texcoord) -> pos
        geometries.append((gid, pos.reshape(-1, 3), idx, tris.get("material")))

    textures = [load_texture(os.path.join(data_dir, fn)) for _, fn in images]

    out_vertices, out_geom, out_mats, out_lights, out_ids = [], [], [], [], []
    cam_orient, cam_fov = None, None
    for vs in lib("library_visual_scenes"):
        for node in vs:
            inst = node.find(NS + "instance_light")
            if inst is None:
                inst = node.find(NS + "instance_geometry")
            if inst is None:
                inst = node.find(NS + "instance_camera")
            if inst is None:
                raise ValueError("VisualSceneConversion error; unsupported node type")
            name = inst.get("url")[1:]
            m = collada_to_vecmath(array_f32(node.find(NS + "matrix").text))
            for cid, fov in cameras:
                if cid == name:
                    if cam_orient is None:
                        cam_orient, cam_fov = m, fov
                    break
            for lid, col in lights:
                if lid == name:
                    p = transform_points(m, np.zeros((1, 3), dtype=F))[0]
                    out_lights.append((p, col))
                    break
            for gid, pos, idx, mat_id in geometries:
                if gid != name:
                    continue
                soup = pos[idx.reshape(-1)]  # de-indexed positions, 3 per triangle
                soup = transform_points(m, soup.astype(F))
                g_index = len(out_mats)
                mat = (0, (F(1000.0), F(0.0), F(1000.0)), 0)  # Material::default (color.rs:37-41)
                eff_id = materials.get(mat_id)
                if eff_id is not None and eff_id in effects:
                    kind, val = effects[eff_id]
                    if kind == "color":
                        mat = (0, tuple(val), 0)
                    else:
                        pos_img = [i for i, (iid, _) in enumerate(images) if iid == val]
                        if not pos_img:
                            raise ValueError("MaterialsConversion error; can't find texture name")
                        mat = (1, (F(0), F(0), F(0)), pos_img[0])
                out_mats.append(mat)
                out_ids.append(gid)
                out_vertices.append(soup.reshape(-1, 9))
                out_geom.append(np.full(len(soup) // 3, g_index, dtype=np.uint32))
                break
    return FlatScene(
        vertices=np.concatenate(out_vertices).astype(F) if out_vertices else np.zeros((0, 9), F),
        tri_geom=np.concatenate(out_geom) if out_geom else np.zeros(0, np.uint32),
        materials=out_mats,
        lights=out_lights,
        textures=textures,
        camera_orientation=cam_orient,
        camera_fov_deg=cam_fov,
        geometry_ids=out_ids,
    )
