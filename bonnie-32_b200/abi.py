"""ctypes binding of include/b32_raster.h (the C ABI of the CUDA rasterizer).

The product path has NO CPU fallback: `load_library()` raises if libb32raster.so is missing, and
`b32_ctx_create` returns B32_ERR_NO_DEVICE when there is no GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B32_LIB") or os.path.join(HERE, "libb32raster.so")   # B32_LIB: experiment builds only

# ---- return codes -------------------------------------------------------------------------
B32_OK, B32_ERR_INVALID, B32_ERR_OOB_INDEX, B32_ERR_NAN_DEPTH = 0, 1, 2, 3
B32_ERR_UNSUPPORTED, B32_ERR_CUDA, B32_ERR_NO_DEVICE = 4, 5, 6
ERR_NAMES = {0: "B32_OK", 1: "B32_ERR_INVALID", 2: "B32_ERR_OOB_INDEX", 3: "B32_ERR_NAN_DEPTH",
             4: "B32_ERR_UNSUPPORTED", 5: "B32_ERR_CUDA", 6: "B32_ERR_NO_DEVICE"}

# ---- enums (src/rasterizer/types.rs:1378-1388, 1288-1293, 1296-1304) -----------------------
BLEND_OPAQUE, BLEND_AVERAGE, BLEND_ADD, BLEND_SUBTRACT, BLEND_ADD_QUARTER, BLEND_ERASE = range(6)
SHADE_NONE, SHADE_FLAT, SHADE_GOURAUD = range(3)
LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_SPOT = range(3)
TEX_RGB555, TEX_IDX8, TEX_IDX4 = range(3)
FACE_TEX_NONE = 0xFFFF
RENDER_ASYNC, RENDER_ALL_OPAQUE, VTX_NO_NORMAL, FACES_IMPLICIT, FACES_UNIFORM = 1, 2, 4, 8, 16

# ---- POD records as numpy dtypes (b32_vertex 36 B, b32_face 16 B) ---------------------------
VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("uv", "<f4", 2), ("normal", "<f4", 3), ("rgba", "u1", 4)])
FACE_DTYPE = np.dtype([("v", "<u4", 3), ("flags", "<u4")])
SKY_VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("rgb", "u1", 3), ("_pad", "u1")])     # b32_sky_vertex
VERTEX_NN_DTYPE = np.dtype([("pos", "<f4", 3), ("uv", "<f4", 2), ("rgba", "u1", 4)])                            # b32_vertex_nn
assert VERTEX_DTYPE.itemsize == 36 and FACE_DTYPE.itemsize == 16 and SKY_VERTEX_DTYPE.itemsize == 16 and VERTEX_NN_DTYPE.itemsize == 24


def compact_buffers(vertices, faces, shading_none):
    """What a marshalling shim sends through b32_render_mesh_15_ex for this mesh: (vertex array, face array, flags).
    Normals are dropped when nothing reads them; an unindexed soup sends only its flags words — or, when they are all
    equal (one texture, one blend mode), a single word that travels with the kernel parameters."""
    flags = 0
    v, f = vertices, faces
    if shading_none:
        nn = np.empty(len(vertices), VERTEX_NN_DTYPE)
        nn["pos"], nn["uv"], nn["rgba"] = vertices["pos"], vertices["uv"], vertices["rgba"]
        v, flags = nn, flags | VTX_NO_NORMAL
    if len(faces) * 3 <= len(vertices) and np.array_equal(faces["v"].reshape(-1), np.arange(len(faces) * 3, dtype=np.uint32)):
        f, flags = np.ascontiguousarray(faces["flags"], dtype=np.uint32), flags | FACES_IMPLICIT
        if len(f) and (f == f[0]).all():
            f, flags = f[:1].copy(), (flags & ~FACES_IMPLICIT) | FACES_UNIFORM
    return np.ascontiguousarray(v), np.ascontiguousarray(f), flags
# b32_line (overlay lines, Framebuffer::draw_line*)
LINE_DTYPE = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"), ("z0", "<f4"), ("z1", "<f4"),
                       ("rgb", "u1", 3), ("blend", "u1"), ("kind", "u1"), ("mode", "u1"), ("alpha", "u1"), ("_pad", "u1")])
assert LINE_DTYPE.itemsize == 32
STAR_DTYPE = np.dtype([("dir", "<f4", 3), ("rgb", "u1", 3), ("_pad", "u1")])                # b32_star
assert STAR_DTYPE.itemsize == 16
LINE_2D, LINE_2D_ALPHA, LINE_3D, LINE_3D_OVERLAY, LINE_3D_ALPHA = 0, 1, 2, 3, 4
LINE_CIRCLE, LINE_CIRCLE_ALPHA, LINE_FILLED_RECT, LINE_THICK = 5, 6, 7, 8
LINE_MAX_COORD = 1 << 20


def face_flags(tex_id=FACE_TEX_NONE, blend=BLEND_OPAQUE, black_transparent=True, editor_alpha=255):
    """B32_FACE_FLAGS of include/b32_raster.h (works on scalars and numpy arrays)."""
    return ((np.asarray(tex_id, dtype=np.uint32) & 0xFFFF)
            | ((np.asarray(blend, dtype=np.uint32) & 7) << 16)
            | (np.asarray(black_transparent, dtype=np.uint32) << 19)
            | ((np.asarray(editor_alpha, dtype=np.uint32) & 0xFF) << 24)).astype(np.uint32)


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("basis_x", C.c_float * 3),
                ("basis_y", C.c_float * 3), ("basis_z", C.c_float * 3)]


class Light(C.Structure):
    _fields_ = [("type", C.c_uint32), ("position", C.c_float * 3), ("direction", C.c_float * 3),
                ("radius", C.c_float), ("angle", C.c_float), ("intensity", C.c_float),
                ("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("enabled", C.c_uint8)]


class Settings(C.Structure):
    _fields_ = [("affine_textures", C.c_uint8), ("use_zbuffer", C.c_uint8), ("shading", C.c_uint8),
                ("backface_cull", C.c_uint8), ("backface_wireframe", C.c_uint8), ("dithering", C.c_uint8),
                ("wireframe_overlay", C.c_uint8), ("use_rgb555", C.c_uint8), ("use_fixed_point", C.c_uint8),
                ("xray_mode", C.c_uint8), ("ortho_enabled", C.c_uint8), ("_pad", C.c_uint8),
                ("ambient", C.c_float), ("ortho_zoom", C.c_float), ("ortho_center_x", C.c_float),
                ("ortho_center_y", C.c_float), ("n_lights", C.c_uint32), ("lights", C.POINTER(Light))]


class Fog(C.Structure):
    _fields_ = [("start", C.c_float), ("falloff", C.c_float), ("cull_distance", C.c_float),
                ("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("blend", C.c_uint8)]


class Timings(C.Structure):
    _fields_ = [("transform_ms", C.c_float), ("fog_ms", C.c_float), ("cull_ms", C.c_float),
                ("sort_ms", C.c_float), ("draw_ms", C.c_float), ("wireframe_ms", C.c_float),
                ("triangles_drawn", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TexDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                ("blend_mode", C.c_uint32), ("pixels", C.c_void_p), ("clut", C.POINTER(C.c_uint16)),
                ("clut_len", C.c_uint32)]


class Tex8Desc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("blend_mode", C.c_uint32), ("_pad", C.c_uint32),
                ("pixels", C.c_void_p)]


class Placement(C.Structure):            # b32_placement
    _fields_ = [("facing", C.c_float), ("cos_f", C.c_float), ("sin_f", C.c_float), ("world_pos", C.c_float * 3)]


# Every symbol include/b32_raster.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "b32_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "b32_ctx_destroy": (None, [_P]),
    "b32_last_error": (C.c_char_p, [_P]),
    "b32_ctx_stream": (_P, [_P]),
    "b32_sync": (C.c_int, [_P]),
    "b32_kernel_launches": (C.c_uint64, [_P]),
    "b32_fb_resize": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "b32_fb_clear": (C.c_int, [_P, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8]),
    "b32_fb_clear_gradient": (C.c_int, [_P] + [C.c_uint8] * 7),
    "b32_draw_lines": (C.c_int, [_P, _P, C.c_uint32]),
    "b32_render_mesh_placed": (C.c_int, [_P, _P, C.POINTER(Placement), C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Fog),
                                         C.c_int, C.c_uint32, C.POINTER(Timings)]),
    "b32_render_stars": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(Camera), C.c_float]),
    "b32_fb_upload": (C.c_int, [_P, _P, _P]),
    "b32_fb_download": (C.c_int, [_P, _P, _P]),
    "b32_fb_size": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b32_textures_set": (C.c_int, [_P, C.POINTER(TexDesc), C.c_uint32]),
    "b32_render_mesh_15": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(Camera),
                                     C.POINTER(Settings), C.POINTER(Fog), C.POINTER(Timings)]),
    "b32_render_mesh_15_ex": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(Camera),
                                        C.POINTER(Settings), C.POINTER(Fog), C.c_uint32, C.POINTER(Timings)]),
    "b32_fb_download_async": (C.c_int, [_P, _P, _P]),
    "b32_mesh_upload": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(_P)]),
    "b32_mesh_free": (None, [_P, _P]),
    "b32_render_mesh_15_resident": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Settings),
                                              C.POINTER(Fog), C.POINTER(Timings)]),
    "b32_render_mesh_15_enqueue": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Fog)]),
    "b32_textures_set_rgb888": (C.c_int, [_P, C.POINTER(Tex8Desc), C.c_uint32]),
    "b32_render_mesh": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Timings)]),
    "b32_render_mesh_resident": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Timings)]),
    "b32_frame_15_enqueue": (C.c_int, [_P, _P, _P, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Fog)]),
    "b32_graph_launches": (C.c_uint64, [_P]),
    "b32_ctx_frame_timings": (C.c_int, [_P, C.c_int]),
    "b32_frame_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "b32_render_skybox_mesh": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(Camera)]),
    "b32_host_alloc": (_P, [C.c_size_t]),
    "b32_host_free": (None, [_P]),
    "b32_debug_transform": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(Camera), C.POINTER(Settings), _P, _P]),
    "b32_debug_draw_order": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "b32_debug_kernel_times": (C.c_int, [_P, C.POINTER(C.c_float), C.c_uint32]),
    "b32_debug_timing_ring": (C.c_int, [_P, C.c_uint32]),
    "b32_debug_prefix_hint": (C.c_int, [_P]),
    "b32_debug_timing_read": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32]),
}

_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libb32raster.so and bind every declared symbol. Raises if the library is missing:
    there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The rasterizer has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B32Error(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}{': ' + detail if detail else ''}")
