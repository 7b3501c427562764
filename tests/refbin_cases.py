"""Scenes that are run through the reference's own compiled rasterizer (docs/bonnie-32.wasm, interpreted by
oracle/wasm/) to pin the oracle.  Shared by the fixture generator (tests/golden/make_ref_wasm.py, build container
only) and by tests/test_ref_wasm.py (runs anywhere: compares the oracle with the committed fixtures).

The binary is crate version 0.1.8: `Face` has no `editor_alpha` there, so every scene is stripped to
editor_alpha = 255 (face blend modes survive: 0.1.8 takes them as a separate slice).
"""
from __future__ import annotations

import copy
import hashlib

import numpy as np

from bonnie32_b200 import abi, scenes
import cases
import fuzz

# oracle compat mask that reproduces the 0.1.8 binary (oracle/b32_oracle.cpp COMPAT_*, oracle/wasm/DRIFT.md)
COMPAT_0_1_8 = 1 | 2 | 4 | 8


def strip_editor_alpha(sc):
    s = copy.copy(sc)
    f = sc.faces.copy()
    f["flags"] = (f["flags"] & np.uint32(0x00FFFFFF)) | np.uint32(0xFF000000)
    s.faces = f
    return s


def has_spot_light(sc):
    return any(int(l.type) == abi.LIGHT_SPOT for l in sc.settings.lights)


def small_scenes():
    out = [scenes.scene_c1(), scenes.scene_c1(False), scenes.scene_c2(), cases._with(scenes.scene_c2(), "c2_1000_zbuffer", use_zbuffer=True),
           cases._with(scenes.scene_c2(), "c2_1000_float", use_fixed_point=False)]
    out += cases.feature_scenes()
    out.append(cases.big_triangle_scene())
    out += cases.wireframe_scenes()
    for seed in range(60):
        sc = fuzz.fuzz_scene(seed)
        if not has_spot_light(sc):
            out.append(sc)
    names = set()
    res = []
    for sc in out:
        assert sc.name not in names, sc.name
        names.add(sc.name)
        res.append(strip_editor_alpha(sc))
    return res


def small_scenes888():
    """The RGB888 sibling `render_mesh` (render.rs:1971-2259): feature scenes + fuzz."""
    out = list(cases.rgb888_scenes())
    for seed in range(40):
        sc = fuzz.fuzz_scene(seed, rgb888=True)
        if not has_spot_light(sc):
            out.append(sc)
    return [strip_editor_alpha(s) for s in out]


def spot_scenes():
    return [strip_editor_alpha(s) for s in cases.spot_scenes()]


def spot_scenes888():
    return [strip_editor_alpha(s) for s in cases.spot_scenes888()]


def big_scenes():
    out = [scenes.scene_c4(), cases._with(scenes.scene_c4(), "c4_100000_zbuffer", use_zbuffer=True),
           cases._with(scenes.scene_c4(), "c4_100000_float_nodither", use_fixed_point=False, dithering=False)]
    out += [scenes.scene_c5(k) for k in range(8)]
    return [strip_editor_alpha(s) for s in out]


def inputs_digest(sc):
    """sha256 over everything render_mesh_15 reads, so a fixture can tell 'scene generator changed' from 'oracle changed'."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(sc.vertices).tobytes())
    h.update(np.ascontiguousarray(sc.faces).tobytes())
    for t in (sc.textures8 or []):
        h.update(np.asarray([t.width, t.height, int(t.blend_mode)], np.uint32).tobytes())
        h.update(np.ascontiguousarray(t.pixels, dtype=np.uint8).tobytes())
    for t in sc.textures:
        e = scenes.expand_texture(t)
        h.update(np.asarray([e.width, e.height, int(e.blend_mode)], np.uint32).tobytes())
        h.update(np.ascontiguousarray(e.pixels, dtype="<u2").tobytes())
    c = sc.camera
    for a in (c.position, c.basis_x, c.basis_y, c.basis_z):
        h.update(np.asarray(a, "<f4").tobytes())
    s = sc.settings
    h.update(repr((bool(s.affine_textures), bool(s.use_zbuffer), int(s.shading), bool(s.backface_cull), bool(s.backface_wireframe),
                   float(np.float32(s.ambient)), bool(s.dithering), bool(s.wireframe_overlay),
                   None if s.ortho_projection is None else tuple(float(np.float32(x)) for x in s.ortho_projection),
                   bool(s.use_fixed_point), bool(s.xray_mode))).encode())
    for l in s.lights:
        h.update(repr((int(l.type), [float(np.float32(x)) for x in l.position], [float(np.float32(x)) for x in l.direction],
                       float(np.float32(l.radius)), float(np.float32(l.intensity)), tuple(int(x) for x in l.color), bool(l.enabled))).encode())
        if int(l.type) == abi.LIGHT_SPOT:                     # the cone angle only exists for Spot lights (older digests stay valid)
            h.update(repr(float(np.float32(l.angle))).encode())
    h.update(repr((None if sc.fog is None else (float(np.float32(sc.fog[0])), float(np.float32(sc.fog[1])), float(np.float32(sc.fog[2])),
                                                tuple(int(x) for x in sc.fog[3][:3])), sc.width, sc.height, tuple(sc.clear[:3]))).encode())
    return h.hexdigest()


def frame_digest(rgba, z):
    return (hashlib.sha256(np.ascontiguousarray(rgba, np.uint8).tobytes()).hexdigest(),
            hashlib.sha256(np.ascontiguousarray(z, "<f4").tobytes()).hexdigest())


def geometry_digest(pos, uv, normal, rgba, face_v, black_transparent, with_uv=True):
    """sha256 of a room's triangles without the texture ids (they depend on the resolver, not on the geometry code).
    with_uv=False also leaves the UVs out (their scale is 32 / texture width, a resolver result)."""
    h = hashlib.sha256()
    for a, dt in ((pos, "<f4"), (uv if with_uv else np.zeros(0), "<f4"), (normal, "<f4"), (rgba, np.uint8), (face_v, "<u4"), (black_transparent, np.uint8)):
        h.update(np.ascontiguousarray(a, dtype=dt).tobytes())
    return h.hexdigest()
