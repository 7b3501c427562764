"""Host-side mirror of the reference rasterizer interface over the C ABI.

Same names, argument meaning and error behaviour as /root/reference/src/rasterizer:
`Framebuffer` (render.rs:10-45), `Camera` (camera.rs:9-91), `RasterSettings` (types.rs:1392-1495),
`Light` (types.rs:1307-1373), `Texture15` (types.rs:532-539), `render_mesh_15` (render.rs:2302-2310),
and the RGB888 siblings `Texture` (types.rs:1058-1066) / `render_mesh` (render.rs:1971-1978).
Where the reference panics (bad vertex index, NaN sort key) the render calls raise `B32Error`.

Vertices and faces are numpy record arrays (`abi.VERTEX_DTYPE`, `abi.FACE_DTYPE`) — the POD layout
the Rust shim marshals `&[Vertex]` / `&[Face]` into.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import abi
from .abi import (BLEND_OPAQUE, SHADE_GOURAUD, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_SPOT,
                  TEX_RGB555, TEX_IDX8, TEX_IDX4, B32Error)

F32 = np.float32


def _v3(x, y, z):
    return np.array([x, y, z], dtype=F32)


def _normalize(v):
    """Vec3::normalize, math.rs:39-49 (sqrt and / are IEEE-exact, so host == reference)."""
    l = np.sqrt(F32(F32(F32(v[0] * v[0]) + F32(v[1] * v[1])) + F32(v[2] * v[2])))
    if l == 0:
        return _v3(0, 0, 0)
    return np.array([v[0] / l, v[1] / l, v[2] / l], dtype=F32)


def _cross(a, b):
    """Vec3::cross, math.rs:27-33."""
    return np.array([F32(a[1] * b[2]) - F32(a[2] * b[1]),
                     F32(a[2] * b[0]) - F32(a[0] * b[2]),
                     F32(a[0] * b[1]) - F32(a[1] * b[0])], dtype=F32)


class Camera:
    """camera.rs:9-91. Basis vectors are computed on the host (libm sin/cos) and passed as data."""

    def __init__(self):
        self.position = _v3(0, 0, 0)
        self.rotation_x = F32(0.0)
        self.rotation_y = F32(0.0)
        self.basis_x = _v3(1, 0, 0)
        self.basis_y = _v3(0, 1, 0)
        self.basis_z = _v3(0, 0, 1)
        self.update_basis()

    def update_basis(self):  # camera.rs:76-91
        upward = _v3(0.0, -1.0, 0.0)
        rx, ry = F32(self.rotation_x), F32(self.rotation_y)
        self.basis_z = np.array([F32(np.cos(rx) * np.sin(ry)), F32(-np.sin(rx)), F32(np.cos(rx) * np.cos(ry))], dtype=F32)
        self.basis_x = _normalize(_cross(upward, self.basis_z))
        self.basis_y = _cross(self.basis_z, self.basis_x)

    def to_abi(self) -> abi.Camera:
        c = abi.Camera()
        for name in ("position", "basis_x", "basis_y", "basis_z"):
            getattr(c, name)[:] = [float(x) for x in getattr(self, name)]
        return c


@dataclass
class Light:
    """types.rs:1307-1373."""
    type: int = LIGHT_DIRECTIONAL
    position: np.ndarray = field(default_factory=lambda: _v3(0, 0, 0))
    direction: np.ndarray = field(default_factory=lambda: _v3(0, 0, 0))
    radius: float = 0.0
    angle: float = 0.0
    intensity: float = 1.0
    color: tuple = (255, 255, 255)
    enabled: bool = True

    @staticmethod
    def directional(direction, intensity):           # types.rs:1317-1325
        return Light(type=LIGHT_DIRECTIONAL, direction=_normalize(np.asarray(direction, dtype=F32)), intensity=intensity)

    @staticmethod
    def point(position, radius, intensity):          # types.rs:1328-1336
        return Light(type=LIGHT_POINT, position=np.asarray(position, dtype=F32), radius=radius, intensity=intensity)

    @staticmethod
    def point_colored(position, radius, intensity, r, g, b):   # types.rs:1339-1352 (`as u8` saturates)
        col = tuple(int(min(max(np.trunc(F32(F32(c) * F32(255.0))), 0), 255)) for c in (r, g, b))
        return Light(type=LIGHT_POINT, position=np.asarray(position, dtype=F32), radius=radius, intensity=intensity, color=col)

    @staticmethod
    def spot(position, direction, angle, radius, intensity):   # types.rs:1355-1368
        return Light(type=LIGHT_SPOT, position=np.asarray(position, dtype=F32),
                     direction=_normalize(np.asarray(direction, dtype=F32)), angle=angle, radius=radius, intensity=intensity)

    def to_abi(self) -> abi.Light:
        l = abi.Light()
        l.type = self.type
        l.position[:] = [float(x) for x in self.position]
        l.direction[:] = [float(x) for x in self.direction]
        l.radius, l.angle, l.intensity = self.radius, self.angle, self.intensity
        l.r, l.g, l.b = self.color
        l.enabled = 1 if self.enabled else 0
        return l


@dataclass
class RasterSettings:
    """types.rs:1392-1428, defaults :1475-1495."""
    affine_textures: bool = True
    use_zbuffer: bool = True
    shading: int = SHADE_GOURAUD
    backface_cull: bool = True
    backface_wireframe: bool = True
    lights: list = field(default_factory=lambda: [Light.directional((-1.0, -1.0, -1.0), 0.7)])
    ambient: float = 0.3
    dithering: bool = True
    wireframe_overlay: bool = False
    ortho_projection: Optional[tuple] = None     # (zoom, center_x, center_y)
    use_rgb555: bool = True
    use_fixed_point: bool = True
    xray_mode: bool = False

    @staticmethod
    def game():                                   # types.rs:1455-1460
        return RasterSettings(backface_wireframe=False)

    @staticmethod
    def modeler():                                # types.rs:1465-1472
        return RasterSettings(backface_wireframe=False, lights=[], ambient=0.7)

    def to_abi(self):
        """Returns (abi.Settings, keepalive)."""
        s = abi.Settings()
        s.affine_textures = self.affine_textures
        s.use_zbuffer = self.use_zbuffer
        s.shading = self.shading
        s.backface_cull = self.backface_cull
        s.backface_wireframe = self.backface_wireframe
        s.dithering = self.dithering
        s.wireframe_overlay = self.wireframe_overlay
        s.use_rgb555 = self.use_rgb555
        s.use_fixed_point = self.use_fixed_point
        s.xray_mode = self.xray_mode
        s.ortho_enabled = self.ortho_projection is not None
        if self.ortho_projection is not None:
            s.ortho_zoom, s.ortho_center_x, s.ortho_center_y = self.ortho_projection
        s.ambient = self.ambient
        arr = (abi.Light * max(1, len(self.lights)))(*[l.to_abi() for l in self.lights])
        s.n_lights = len(self.lights)
        s.lights = C.cast(arr, C.POINTER(abi.Light))
        return s, arr


@dataclass
class Texture15:
    """types.rs:532-539, or an indexed texture + CLUT (types.rs:438; mesh_editor.rs:669-682)."""
    width: int
    height: int
    pixels: np.ndarray                      # u16[h*w] | u8[h*w] | u8[(h*w+1)//2]
    blend_mode: int = BLEND_OPAQUE
    format: int = TEX_RGB555
    clut: Optional[np.ndarray] = None       # u16[clut_len]

    def to_abi(self):
        d = abi.TexDesc()
        d.width, d.height, d.format, d.blend_mode = self.width, self.height, self.format, self.blend_mode
        want = np.uint16 if self.format == TEX_RGB555 else np.uint8
        px = np.ascontiguousarray(self.pixels, dtype=want)
        d.pixels = px.ctypes.data
        keep = [px]
        if self.clut is not None:
            cl = np.ascontiguousarray(self.clut, dtype=np.uint16)
            d.clut = cl.ctypes.data_as(C.POINTER(C.c_uint16))
            d.clut_len = cl.size
            keep.append(cl)
        return d, keep


def tex_descs(textures: Sequence[Texture15]):
    keep = []
    arr = (abi.TexDesc * max(1, len(textures)))()
    for i, t in enumerate(textures):
        d, k = t.to_abi()
        arr[i] = d
        keep.append(k)
    return arr, keep


@dataclass
class Texture:
    """RGB888 texture, types.rs:1058-1066: one Color (r, g, b, blend) per texel, blend == Erase = transparent."""
    width: int
    height: int
    pixels: np.ndarray                      # u8[h*w*4]
    blend_mode: int = BLEND_OPAQUE
    name: str = ""

    def to_abi(self):
        d = abi.Tex8Desc()
        d.width, d.height, d.blend_mode = self.width, self.height, self.blend_mode
        px = np.ascontiguousarray(self.pixels, dtype=np.uint8)
        assert px.size == self.width * self.height * 4
        d.pixels = px.ctypes.data
        return d, [px]


def tex8_descs(textures: Sequence[Texture]):
    keep = []
    arr = (abi.Tex8Desc * max(1, len(textures)))()
    for i, t in enumerate(textures):
        d, k = t.to_abi()
        arr[i] = d
        keep.append(k)
    return arr, keep


def fog_to_abi(fog):
    """fog: None or (start, falloff, cull_distance, (r, g, b[, blend]))."""
    if fog is None:
        return None
    f = abi.Fog()
    f.start, f.falloff, f.cull_distance = fog[0], fog[1], fog[2]
    col = tuple(fog[3])
    f.r, f.g, f.b = col[:3]
    f.blend = col[3] if len(col) > 3 else BLEND_OPAQUE
    return f


class Context:
    """One GPU, one stream, one device-resident framebuffer (b32_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        rc = self.lib.b32_ctx_create(device, C.byref(h))
        if rc != abi.B32_OK:
            raise B32Error(rc, "b32_ctx_create failed (no CPU fallback exists)")
        self.h = h
        self._tex_ref = None      # the list last uploaded (kept alive so identity stays meaningful)
        self._tex8_ref = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.b32_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != abi.B32_OK:
            msg = self.lib.b32_last_error(self.h)
            raise B32Error(rc, msg.decode() if msg else "")

    @property
    def stream(self) -> int:
        return int(self.lib.b32_ctx_stream(self.h) or 0)

    def sync(self):
        self.check(self.lib.b32_sync(self.h))

    def kernel_launches(self) -> int:
        return int(self.lib.b32_kernel_launches(self.h))

    def graph_launches(self) -> int:
        return int(self.lib.b32_graph_launches(self.h))

    def set_textures(self, textures: Sequence[Texture15]):
        arr, keep = tex_descs(textures)
        self.check(self.lib.b32_textures_set(self.h, arr, len(textures)))
        self._tex_ref = textures


    def set_textures_rgb888(self, textures: Sequence[Texture]):
        arr, keep = tex8_descs(textures)
        self.check(self.lib.b32_textures_set_rgb888(self.h, arr, len(textures)))
        self._tex8_ref = textures


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Framebuffer:
    """render.rs:10-45. `pixels` (RGBA8) and `zbuffer` (f32) live on the device; the numpy views
    returned by `.pixels` / `.zbuffer` are downloads (the frame-end sync point)."""

    def __init__(self, width: int, height: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.width, self.height = 0, 0
        self.resize(width, height)
        self.clear_transparent()                    # Framebuffer::new: pixels 0, zbuffer f32::MAX

    def resize(self, width: int, height: int):      # render.rs:27-34 (no-op when the size is unchanged)
        self.ctx.check(self.ctx.lib.b32_fb_resize(self.ctx.h, width, height))
        self.width, self.height = width, height

    def clear(self, color):                         # render.rs:36-45; color = (r, g, b[, blend])
        r, g, b = color[:3]
        a = 0 if (len(color) > 3 and color[3] == abi.BLEND_ERASE) else 255   # Color::to_bytes
        self.ctx.check(self.ctx.lib.b32_fb_clear(self.ctx.h, r, g, b, a))

    def clear_transparent(self):                    # render.rs:48-56
        self.ctx.check(self.ctx.lib.b32_fb_clear(self.ctx.h, 0, 0, 0, 0))

    def clear_gradient(self, top_color, bottom_color):   # render.rs:60-77
        a = 0 if (len(top_color) > 3 and top_color[3] == abi.BLEND_ERASE) else 255
        self.ctx.check(self.ctx.lib.b32_fb_clear_gradient(self.ctx.h, *top_color[:3], *bottom_color[:3], a))

    # ---- overlay lines (render.rs:684-872).  One list = one device pass with the result of the calls made in order;
    # the single-line methods below are the reference's signatures.
    def draw_lines(self, lines: np.ndarray):
        ln = np.ascontiguousarray(lines, dtype=abi.LINE_DTYPE)
        self.ctx.check(self.ctx.lib.b32_draw_lines(self.ctx.h, ln.ctypes.data, len(ln)))

    def draw_line(self, x0, y0, x1, y1, color):
        self.draw_lines(make_lines([line_entry(abi.LINE_2D, x0, y0, x1, y1, color)]))

    def draw_line_blended(self, x0, y0, x1, y1, color, mode):
        self.draw_lines(make_lines([line_entry(abi.LINE_2D, x0, y0, x1, y1, color, mode=mode)]))

    def draw_line_alpha(self, x0, y0, x1, y1, color, alpha):
        self.draw_lines(make_lines([line_entry(abi.LINE_2D_ALPHA, x0, y0, x1, y1, color, alpha=alpha)]))

    def draw_line_3d(self, x0, y0, z0, x1, y1, z1, color):
        self.draw_lines(make_lines([line_entry(abi.LINE_3D, x0, y0, x1, y1, color, z0, z1)]))

    def draw_line_3d_overlay(self, x0, y0, z0, x1, y1, z1, color):
        self.draw_lines(make_lines([line_entry(abi.LINE_3D_OVERLAY, x0, y0, x1, y1, color, z0, z1)]))

    def draw_line_3d_alpha(self, x0, y0, z0, x1, y1, z1, color, alpha):
        self.draw_lines(make_lines([line_entry(abi.LINE_3D_ALPHA, x0, y0, x1, y1, color, z0, z1, alpha=alpha)]))

    def draw_circle(self, cx, cy, radius, color):                            # render.rs:631-644
        self.draw_lines(make_lines([line_entry(abi.LINE_CIRCLE, cx, cy, radius, 0, color)]))

    def draw_circle_alpha(self, cx, cy, radius, color, alpha):               # :670-682
        self.draw_lines(make_lines([line_entry(abi.LINE_CIRCLE_ALPHA, cx, cy, radius, 0, color, alpha=alpha)]))

    def draw_thick_line(self, x0, y0, x1, y1, thickness, color):             # :875-938
        self.draw_lines(make_lines([line_entry(abi.LINE_THICK, x0, y0, x1, y1, color, z0=float(thickness))]))

    def draw_rect(self, x0, y0, x1, y1, color):                              # :941-951: four draw_line calls
        self.draw_lines(make_lines(rect_entries(x0, y0, x1, y1, color)))

    def draw_filled_rect(self, x0, y0, x1, y1, color):                       # :954-972
        self.draw_lines(make_lines([line_entry(abi.LINE_FILLED_RECT, x0, y0, x1, y1, color)]))

    def upload(self, pixels: np.ndarray, zbuffer: Optional[np.ndarray] = None):
        px = np.ascontiguousarray(pixels, dtype=np.uint8)
        assert px.size == self.width * self.height * 4
        zp = None
        if zbuffer is not None:
            zb = np.ascontiguousarray(zbuffer, dtype=np.float32)
            assert zb.size == self.width * self.height
            zp = zb.ctypes.data
        self.ctx.check(self.ctx.lib.b32_fb_upload(self.ctx.h, px.ctypes.data, zp))

    def download(self, want_z: bool = True):
        px = np.empty((self.height, self.width, 4), dtype=np.uint8)
        zb = np.empty((self.height, self.width), dtype=np.float32) if want_z else None
        self.ctx.check(self.ctx.lib.b32_fb_download(self.ctx.h, px.ctypes.data, zb.ctypes.data if want_z else None))
        return px, zb

    def download_view(self, want_z: bool = False):
        """download() into pinned buffers owned by this Framebuffer (direct DMA, no allocation per frame): what a game
        loop hands to `Texture2D::from_rgba8`.  The returned arrays are views, valid until the next download_view /
        resize of this object."""
        n = self.width * self.height
        if getattr(self, "_pin_n", 0) != n:
            self._free_pinned()
            self._pin = self.ctx.lib.b32_host_alloc(n * 8)
            if not self._pin:
                raise B32Error(abi.B32_ERR_CUDA, "b32_host_alloc failed")
            self._pin_n = n
        self.ctx.check(self.ctx.lib.b32_fb_download(self.ctx.h, self._pin, (self._pin + n * 4) if want_z else None))
        px = np.ctypeslib.as_array(C.cast(self._pin, C.POINTER(C.c_uint8)), shape=(self.height, self.width, 4))
        zb = np.ctypeslib.as_array(C.cast(self._pin + n * 4, C.POINTER(C.c_float)), shape=(self.height, self.width)) if want_z else None
        return px, zb

    def _free_pinned(self):
        if getattr(self, "_pin", None):
            self.ctx.lib.b32_host_free(self._pin)
        self._pin, self._pin_n = None, 0

    def __del__(self):
        try:
            self._free_pinned()
        except Exception:
            pass

    def render_skybox_mesh(self, sky_vertices: np.ndarray, faces: np.ndarray, camera: "Camera"):
        """Sphere pass of Framebuffer::render_skybox (render.rs:81-139): `sky_vertices` (abi.SKY_VERTEX_DTYPE) and
        `faces` (int[nf,3]) are what Skybox::generate_mesh returns; stars stay a host pass."""
        v = np.ascontiguousarray(sky_vertices, dtype=abi.SKY_VERTEX_DTYPE)
        f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
        cam = camera.to_abi()
        self.ctx.check(self.ctx.lib.b32_render_skybox_mesh(self.ctx.h, v.ctypes.data, len(v), f.ctypes.data, len(f) // 3, C.byref(cam)))

    def render_stars(self, stars: np.ndarray, camera: "Camera", size: float):
        """Star pass of Framebuffer::render_skybox (render.rs:149-235): `stars` (abi.STAR_DTYPE) holds, per star of the
        reference's loop, the direction and the brightness-scaled colour the host computed with libm + the LCG."""
        st = np.ascontiguousarray(stars, dtype=abi.STAR_DTYPE)
        cam = camera.to_abi()
        self.ctx.check(self.ctx.lib.b32_render_stars(self.ctx.h, st.ctypes.data, len(st), C.byref(cam), float(size)))

    @property
    def pixels(self) -> np.ndarray:
        return self.download(False)[0]

    @property
    def zbuffer(self) -> np.ndarray:
        return self.download(True)[1]


_LIBM = None


def _libm():
    global _LIBM
    if _LIBM is None:
        _LIBM = C.CDLL("libm.so.6")
        for f in (_LIBM.cosf, _LIBM.sinf):
            f.restype, f.argtypes = C.c_float, [C.c_float]
    return _LIBM


def libm_cosf(x: float) -> float:
    """f32::cos as Rust evaluates it on Linux (libm cosf), for host-side values that feed the device."""
    return float(_libm().cosf(float(x)))


def libm_sinf(x: float) -> float:
    return float(_libm().sinf(float(x)))


def line_entry(kind, x0, y0, x1, y1, color, z0=0.0, z1=0.0, mode=abi.BLEND_OPAQUE, alpha=255):
    """One b32_line: `color` = (r, g, b[, blend]) as everywhere in this module."""
    blend = color[3] if len(color) > 3 else abi.BLEND_OPAQUE
    return (x0, y0, x1, y1, z0, z1, tuple(color[:3]), blend, kind, mode, alpha, 0)


def rect_entries(x0, y0, x1, y1, color):
    """draw_rect (render.rs:941-951): top, right, bottom, left as B32_LINE_2D entries."""
    min_x, max_x = (x0, x1) if x0 < x1 else (x1, x0)
    min_y, max_y = (y0, y1) if y0 < y1 else (y1, y0)
    return [line_entry(abi.LINE_2D, min_x, min_y, max_x, min_y, color), line_entry(abi.LINE_2D, max_x, min_y, max_x, max_y, color),
            line_entry(abi.LINE_2D, max_x, max_y, min_x, max_y, color), line_entry(abi.LINE_2D, min_x, max_y, min_x, min_y, color)]


def make_lines(entries) -> np.ndarray:
    return np.array(list(entries), dtype=abi.LINE_DTYPE)


def _check_geometry(vertices, faces):
    v = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
    f = np.ascontiguousarray(faces, dtype=abi.FACE_DTYPE)
    return v, f


def render_mesh_15(fb: Framebuffer, vertices: np.ndarray, faces: np.ndarray,
                   textures: Sequence[Texture15], camera: Camera, settings: RasterSettings,
                   fog=None) -> dict:
    """render.rs:2302-2310. Returns RasterTimings as a dict (types.rs:1499-1514)."""
    ctx = fb.ctx
    v, f = _check_geometry(vertices, faces)
    if ctx._tex_ref is not textures:     # cf. textures_15_cache_generation, src/editor/viewport_3d.rs:3459
        ctx.set_textures(textures)
    cam = camera.to_abi()
    s, keep = settings.to_abi()
    fg = fog_to_abi(fog)
    tm = abi.Timings()
    rc = ctx.lib.b32_render_mesh_15(ctx.h, v.ctypes.data, len(v), f.ctypes.data, len(f),
                                    C.byref(cam), C.byref(s), C.byref(fg) if fg is not None else None, C.byref(tm))
    del keep
    ctx.check(rc)
    return tm.as_dict()


def render_mesh(fb: Framebuffer, vertices: np.ndarray, faces: np.ndarray, textures: Sequence[Texture],
                camera: Camera, settings: RasterSettings) -> dict:
    """render.rs:1971-1978, the RGB888 sibling (RasterSettings.use_rgb555 == false)."""
    ctx = fb.ctx
    v, f = _check_geometry(vertices, faces)
    if ctx._tex8_ref is not textures:
        ctx.set_textures_rgb888(textures)
    cam = camera.to_abi()
    s, keep = settings.to_abi()
    tm = abi.Timings()
    rc = ctx.lib.b32_render_mesh(ctx.h, v.ctypes.data, len(v), f.ctypes.data, len(f), C.byref(cam), C.byref(s), C.byref(tm))
    del keep
    ctx.check(rc)
    return tm.as_dict()


class Mesh:
    """Device-resident geometry (b32_mesh): upload once, render many times."""

    def __init__(self, ctx: Context, vertices: np.ndarray, faces: np.ndarray):
        self.ctx = ctx
        v, f = _check_geometry(vertices, faces)
        h = C.c_void_p()
        ctx.check(ctx.lib.b32_mesh_upload(ctx.h, v.ctypes.data, len(v), f.ctypes.data, len(f), C.byref(h)))
        self.h = h
        self.nv, self.nf = len(v), len(f)

    def free(self):
        if self.h:
            self.ctx.lib.b32_mesh_free(self.ctx.h, self.h)
            self.h = None

    def render(self, camera: Camera, settings: RasterSettings, fog=None, enqueue_only=False):
        cam = camera.to_abi()
        s, keep = settings.to_abi()
        fg = fog_to_abi(fog)
        fgp = C.byref(fg) if fg is not None else None
        if enqueue_only:
            self.ctx.check(self.ctx.lib.b32_render_mesh_15_enqueue(self.ctx.h, self.h, C.byref(cam), C.byref(s), fgp))
            return None
        tm = abi.Timings()
        self.ctx.check(self.ctx.lib.b32_render_mesh_15_resident(self.ctx.h, self.h, C.byref(cam), C.byref(s), fgp, C.byref(tm)))
        return tm.as_dict()

    def frame_enqueue(self, clear, camera: Camera, settings: RasterSettings, fog=None):
        """b32_frame_15_enqueue: Framebuffer::clear(clear) (None = keep) + render_mesh_15 as one enqueued frame
        (a CUDA graph launch from the third frame of this mesh on).  Errors surface at sync / download."""
        cam = camera.to_abi()
        s, keep = settings.to_abi()
        fg = fog_to_abi(fog)
        col = None
        if clear is not None:
            a = 0 if (len(clear) > 3 and clear[3] == abi.BLEND_ERASE) else 255
            col = (C.c_uint8 * 4)(clear[0], clear[1], clear[2], a)
        self.ctx.check(self.ctx.lib.b32_frame_15_enqueue(self.ctx.h, col, self.h, C.byref(cam), C.byref(s),
                                                         C.byref(fg) if fg is not None else None))

    def render_placed(self, camera: Camera, settings: RasterSettings, facing: float, world_pos, fog=None, rgb888=False,
                      enqueue_only=False):
        """One part of render_asset_parts (src/scene.rs:109-169): this resident mesh rotated about Y by `facing` and moved
        to `world_pos` on the device, then render_mesh_15 / render_mesh.  cos/sin are libm's f32 functions, as in Rust."""
        pl = abi.Placement(facing, libm_cosf(facing), libm_sinf(facing), (C.c_float * 3)(*[float(x) for x in world_pos]))
        cam = camera.to_abi()
        s, keep = settings.to_abi()
        fg = fog_to_abi(fog)
        tm = abi.Timings()
        self.ctx.check(self.ctx.lib.b32_render_mesh_placed(self.ctx.h, self.h, C.byref(pl), C.byref(cam), C.byref(s),
                                                           C.byref(fg) if fg is not None else None, int(rgb888),
                                                           abi.RENDER_ASYNC if enqueue_only else 0, C.byref(tm)))
        return None if enqueue_only else tm.as_dict()

    def render_rgb888(self, camera: Camera, settings: RasterSettings):
        """render_mesh (RGB888) on the resident geometry; textures come from Context.set_textures_rgb888."""
        cam = camera.to_abi()
        s, keep = settings.to_abi()
        tm = abi.Timings()
        self.ctx.check(self.ctx.lib.b32_render_mesh_resident(self.ctx.h, self.h, C.byref(cam), C.byref(s), C.byref(tm)))
        return tm.as_dict()
