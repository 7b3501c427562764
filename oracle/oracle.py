"""ctypes wrapper of oracle/libb32oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
The C++ behind it restates /root/reference/src/rasterizer line by line (see b32_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libb32oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "b32_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libb32oracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.b32o_unr_table.restype = C.c_uint8
        _lib.b32o_unr_table.argtypes = [C.c_uint32]
        _lib.b32o_fixed_from_f32.restype = C.c_int32
        _lib.b32o_fixed_from_f32.argtypes = [C.c_float]
        _lib.b32o_fixed_mul.restype = C.c_int32
        _lib.b32o_fixed_mul.argtypes = [C.c_int32, C.c_int32]
        _lib.b32o_div_unr.restype = C.c_int32
        _lib.b32o_div_unr.argtypes = [C.c_int32, C.c_int32]
        _lib.b32o_texture_sample.restype = C.c_uint16
        _lib.b32o_render_mesh_15.restype = C.c_int
        _lib.b32o_render_mesh.restype = C.c_int
    return _lib


def _abi():
    import bonnie32_b200.abi as abi   # registered by __graft_entry__.load_package()
    return abi


def render_scene(scene, fb_rgba=None, fb_z=None, want_order=False):
    """Run the oracle's render_mesh_15 on a `scenes.Scene`. Returns (rgba[h,w,4], z[h,w], timings, rc[, order])."""
    abi = _abi()
    from bonnie32_b200.raster import fog_to_abi, tex_descs
    w, h = scene.width, scene.height
    if fb_rgba is None:
        fb_rgba = np.empty((h, w, 4), dtype=np.uint8)
        fb_z = np.empty((h, w), dtype=np.float32)
        r, g, b = scene.clear[:3]
        lib().b32o_fb_clear(fb_rgba.ctypes.data_as(C.c_void_p), fb_z.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h),
                            C.c_uint8(r), C.c_uint8(g), C.c_uint8(b), C.c_uint8(255))
    rc, tm, order = render_mesh_15(fb_rgba, fb_z, scene.vertices, scene.faces, scene.textures, scene.camera,
                                   scene.settings, scene.fog, want_order=True)
    if want_order:
        return fb_rgba, fb_z, tm, rc, order
    return fb_rgba, fb_z, tm, rc


def render_mesh_15(fb_rgba, fb_z, vertices, faces, textures, camera, settings, fog=None, want_order=False):
    """Oracle render_mesh_15 into caller-owned numpy framebuffer arrays (modified in place)."""
    abi = _abi()
    from bonnie32_b200.raster import fog_to_abi, tex_descs
    h, w = fb_z.shape
    v = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
    f = np.ascontiguousarray(faces, dtype=abi.FACE_DTYPE)
    tex, keep_t = tex_descs(textures)
    cam = camera.to_abi()
    s, keep_s = settings.to_abi()
    fg = fog_to_abi(fog)
    tm = abi.Timings()
    cap = len(f) if want_order else 0
    order = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_uint32(0)
    rc = lib().b32o_render_mesh_15(
        C.c_void_p(fb_rgba.ctypes.data), C.c_void_p(fb_z.ctypes.data), C.c_uint32(w), C.c_uint32(h),
        C.c_void_p(v.ctypes.data), C.c_uint32(len(v)), C.c_void_p(f.ctypes.data), C.c_uint32(len(f)),
        tex, C.c_uint32(len(textures)), C.byref(cam), C.byref(s), C.byref(fg) if fg is not None else None,
        C.byref(tm), C.c_void_p(order.ctypes.data) if want_order else None, C.c_uint32(cap), C.byref(n))
    del keep_t, keep_s
    return rc, tm.as_dict(), order[: n.value] if want_order else None


def render_mesh(fb_rgba, fb_z, vertices, faces, textures8, camera, settings, want_order=False):
    """Oracle render_mesh (RGB888, render.rs:1971-2259) into caller-owned numpy framebuffer arrays."""
    abi = _abi()
    from bonnie32_b200.raster import tex8_descs
    h, w = fb_z.shape
    v = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
    f = np.ascontiguousarray(faces, dtype=abi.FACE_DTYPE)
    tex, keep_t = tex8_descs(textures8)
    cam = camera.to_abi()
    s, keep_s = settings.to_abi()
    tm = abi.Timings()
    cap = len(f) if want_order else 0
    order = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_uint32(0)
    rc = lib().b32o_render_mesh(
        C.c_void_p(fb_rgba.ctypes.data), C.c_void_p(fb_z.ctypes.data), C.c_uint32(w), C.c_uint32(h),
        C.c_void_p(v.ctypes.data), C.c_uint32(len(v)), C.c_void_p(f.ctypes.data), C.c_uint32(len(f)),
        tex, C.c_uint32(len(textures8)), C.byref(cam), C.byref(s),
        C.byref(tm), C.c_void_p(order.ctypes.data) if want_order else None, C.c_uint32(cap), C.byref(n))
    del keep_t, keep_s
    return rc, tm.as_dict(), order[: n.value] if want_order else None


def render_scene888(scene, want_order=False):
    """Oracle render_mesh on a Scene with `textures8`. Returns (rgba, z, timings, rc[, order])."""
    w, h = scene.width, scene.height
    fb_rgba = np.empty((h, w, 4), dtype=np.uint8)
    fb_z = np.empty((h, w), dtype=np.float32)
    r, g, b = scene.clear[:3]
    lib().b32o_fb_clear(fb_rgba.ctypes.data_as(C.c_void_p), fb_z.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h),
                        C.c_uint8(r), C.c_uint8(g), C.c_uint8(b), C.c_uint8(255))
    rc, tm, order = render_mesh(fb_rgba, fb_z, scene.vertices, scene.faces, scene.textures8, scene.camera, scene.settings, want_order=True)
    if want_order:
        return fb_rgba, fb_z, tm, rc, order
    return fb_rgba, fb_z, tm, rc


def render_skybox_mesh(fb_rgba, sky_vertices, faces, camera):
    """Oracle sphere pass of Framebuffer::render_skybox into a caller-owned u8[h,w,4] array."""
    abi = _abi()
    h, w = fb_rgba.shape[:2]
    v = np.ascontiguousarray(sky_vertices, dtype=abi.SKY_VERTEX_DTYPE)
    f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
    cam = camera.to_abi()
    lib().b32o_render_skybox_mesh.restype = C.c_int
    return lib().b32o_render_skybox_mesh(C.c_void_p(fb_rgba.ctypes.data), C.c_uint32(w), C.c_uint32(h), C.c_void_p(v.ctypes.data),
                                         C.c_uint32(len(v)), C.c_void_p(f.ctypes.data), C.c_uint32(len(f) // 3), C.byref(cam))


def render_stars(fb_rgba, stars, camera, size):
    """Star pass of Framebuffer::render_skybox into a caller-owned u8[h,w,4] array; stars = abi.STAR_DTYPE records."""
    abi = _abi()
    h, w = fb_rgba.shape[:2]
    st = np.ascontiguousarray(stars, dtype=abi.STAR_DTYPE)
    cam = camera.to_abi()
    lib().b32o_render_stars.restype = C.c_int
    return lib().b32o_render_stars(C.c_void_p(fb_rgba.ctypes.data), C.c_uint32(w), C.c_uint32(h), C.c_void_p(st.ctypes.data),
                                   C.c_uint32(len(st)), C.byref(cam), C.c_float(size))


def place_vertices(vertices, facing, cos_f, sin_f, world_pos):
    """render_asset_parts' per-object transform (scene.rs:121-160) of an abi.VERTEX_DTYPE array."""
    abi = _abi()
    v = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
    out = np.empty_like(v)
    wp = (C.c_float * 3)(*[float(x) for x in world_pos])
    lib().b32o_place_vertices.restype = C.c_int
    lib().b32o_place_vertices(C.c_void_p(v.ctypes.data), C.c_uint32(len(v)), C.c_float(facing), C.c_float(cos_f), C.c_float(sin_f), wp,
                              C.c_void_p(out.ctypes.data))
    return out


def fb_clear_gradient(fb_rgba, fb_z, top, bottom):
    """Framebuffer::clear_gradient on caller-owned arrays; top/bottom = (r, g, b[, blend])."""
    abi = _abi()
    h, w = fb_rgba.shape[:2]
    a = 0 if (len(top) > 3 and top[3] == abi.BLEND_ERASE) else 255
    t = (C.c_uint8 * 3)(*top[:3]); b = (C.c_uint8 * 3)(*bottom[:3])
    lib().b32o_fb_clear_gradient.restype = None
    lib().b32o_fb_clear_gradient(C.c_void_p(fb_rgba.ctypes.data), C.c_void_p(fb_z.ctypes.data), C.c_uint32(w), C.c_uint32(h), t, b, C.c_uint8(a))


def draw_lines(fb_rgba, fb_z, lines):
    """Framebuffer::draw_line* for every entry of `lines` (abi.LINE_DTYPE), in order."""
    abi = _abi()
    h, w = fb_rgba.shape[:2]
    ln = np.ascontiguousarray(lines, dtype=abi.LINE_DTYPE)
    lib().b32o_draw_lines.restype = C.c_int
    return lib().b32o_draw_lines(C.c_void_p(fb_rgba.ctypes.data), C.c_void_p(fb_z.ctypes.data), C.c_uint32(w), C.c_uint32(h),
                                 C.c_void_p(ln.ctypes.data), C.c_uint32(len(ln)))


def transform(vertices, camera, settings, w, h):
    abi = _abi()
    v = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
    scr = np.empty((len(v), 3), dtype=np.float32)
    cam = np.empty((len(v), 3), dtype=np.float32)
    c = camera.to_abi()
    s, keep = settings.to_abi()
    lib().b32o_transform(C.c_void_p(v.ctypes.data), C.c_uint32(len(v)), C.byref(c), C.byref(s), C.c_uint32(w), C.c_uint32(h),
                         C.c_void_p(scr.ctypes.data), C.c_void_p(cam.ctypes.data))
    return scr, cam
