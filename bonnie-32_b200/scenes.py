"""Synthetic scenes of BASELINE.json `configs` (made concrete in SURVEY.md §8d).

C1  single flat-shaded triangle (+ its reversed winding, which must be culled)
C2  1 000 random affine-textured triangles, 64x64 8-bit atlas, painter's sort
C4  100 000-triangle stress scene, 256x256 4-bit atlas
C5  8 independent C4 frames (seed 0xB3200500 + k)

PRNG: SplitMix64, u01 = (z >> 40) * 2^-24 (exact in f32). One stream per scene; draws are consumed
in this order: per triangle [cz, cx, cy], then per vertex [dx, dy, dz, u, v, r, g, b]; after all
triangles the CLUT entries 1.., then the atlas indices row-major.  Coordinates are computed in
float64 and rounded once to f32.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi
from .raster import Camera, RasterSettings, Texture15

SEED_C2 = 0xB3200002
SEED_C4 = 0xB3200004
SEED_C5 = 0xB3200500
CLEAR_COLOR = (20, 22, 28)      # src/game/renderer.rs:95 style clear, not 5-bit representable

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64_u01(seed: int, n: int) -> np.ndarray:
    """n draws of u01 as float64 (each exactly representable in f32)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        x = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float64) * (2.0 ** -24)


@dataclass
class Scene:
    name: str
    vertices: np.ndarray
    faces: np.ndarray
    textures: List[Texture15]
    camera: Camera
    settings: RasterSettings
    fog: Optional[tuple] = None
    width: int = 320
    height: int = 240
    clear: tuple = CLEAR_COLOR
    textures8: Optional[list] = None        # RGB888 scenes: raster.Texture list for render_mesh

    @property
    def algorithmic_bytes(self) -> int:
        """nv*36 + nf*16 + atlas + clut + 48 + W*H*4."""
        tex = 0
        for t in self.textures:
            tex += t.pixels.nbytes + (t.clut.nbytes if t.clut is not None else 0)
        return len(self.vertices) * 36 + len(self.faces) * 16 + tex + 48 + self.width * self.height * 4


def common_settings(**kw) -> RasterSettings:
    """SURVEY.md §8d 'Common settings for C1-C5'."""
    s = RasterSettings(affine_textures=True, use_zbuffer=False, shading=abi.SHADE_NONE, backface_cull=True,
                       backface_wireframe=False, lights=[], ambient=0.3, dithering=True, wireframe_overlay=False,
                       ortho_projection=None, use_rgb555=True, use_fixed_point=True, xray_mode=False)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def make_vertices(pos, uv=None, normal=None, rgba=None) -> np.ndarray:
    n = len(pos)
    v = np.zeros(n, dtype=abi.VERTEX_DTYPE)
    v["pos"] = np.asarray(pos, dtype=np.float32)
    if uv is not None:
        v["uv"] = np.asarray(uv, dtype=np.float32)
    if normal is not None:
        v["normal"] = np.asarray(normal, dtype=np.float32)
    if rgba is None:
        v["rgba"] = np.array([128, 128, 128, abi.BLEND_OPAQUE], dtype=np.uint8)   # Color::NEUTRAL
    else:
        v["rgba"] = np.asarray(rgba, dtype=np.uint8)
    return v


def make_faces(idx, tex_id=abi.FACE_TEX_NONE, blend=abi.BLEND_OPAQUE, black_transparent=True, editor_alpha=255) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.uint32).reshape(-1, 3)
    f = np.zeros(len(idx), dtype=abi.FACE_DTYPE)
    f["v"] = idx
    f["flags"] = abi.face_flags(tex_id, blend, black_transparent, editor_alpha)
    return f


def scene_c1(use_fixed_point: bool = True) -> Scene:
    """Single flat-shaded triangle + reversed copy (culled). SURVEY.md §8d C1."""
    pos = [(-1, -1, 3), (1, -1, 3), (0, 1, 3)]
    v = make_vertices(pos, normal=[(0, 0, -1)] * 3, rgba=[(200, 100, 50, abi.BLEND_OPAQUE)] * 3)
    f = make_faces([(0, 1, 2), (0, 2, 1)])
    s = common_settings(shading=abi.SHADE_FLAT, ambient=1.0, use_fixed_point=use_fixed_point)
    return Scene("c1_single_triangle" + ("" if use_fixed_point else "_float"), v, f, [], Camera(), s)


def _random_triangles(seed: int, n_tris: int):
    per_tri = 3 + 3 * 8
    u = splitmix64_u01(seed, n_tris * per_tri).reshape(n_tris, per_tri)
    cz = 2.0 + 58.0 * u[:, 0]
    cx = (2.0 * u[:, 1] - 1.0) * 0.5 * (cz + 5.0)
    cy = (2.0 * u[:, 2] - 1.0) * 0.4 * (cz + 5.0)
    r = 0.04 * (cz + 5.0)
    pv = u[:, 3:].reshape(n_tris, 3, 8)
    pos = np.stack([cx[:, None] + r[:, None] * (2.0 * pv[:, :, 0] - 1.0),
                    cy[:, None] + r[:, None] * (2.0 * pv[:, :, 1] - 1.0),
                    cz[:, None] + r[:, None] * (2.0 * pv[:, :, 2] - 1.0)], axis=-1)
    uv = 2.0 * pv[:, :, 3:5]
    col = 64 + np.floor(128.0 * pv[:, :, 5:8]).astype(np.int64)
    rgba = np.concatenate([col, np.zeros((n_tris, 3, 1), dtype=np.int64)], axis=-1)
    v = make_vertices(pos.reshape(-1, 3), uv=uv.reshape(-1, 2), normal=np.tile([0.0, 0.0, -1.0], (n_tris * 3, 1)),
                      rgba=rgba.reshape(-1, 4))
    f = make_faces(np.arange(n_tris * 3).reshape(n_tris, 3), tex_id=0)
    return v, f, n_tris * per_tri


def _atlas(seed: int, consumed: int, size: int, bits: int) -> Texture15:
    ncol = 1 << bits
    n = (ncol - 1) + size * size
    u = splitmix64_u01(seed, consumed + n)[consumed:]
    clut = np.zeros(ncol, dtype=np.uint16)
    clut[1:] = np.floor(32768.0 * u[: ncol - 1]).astype(np.uint16)          # bit15 clear
    idx = np.floor(float(ncol) * u[ncol - 1:]).astype(np.uint8)
    if bits == 8:
        return Texture15(size, size, idx, format=abi.TEX_IDX8, clut=clut)
    packed = (idx[0::2] | (idx[1::2] << 4)).astype(np.uint8)                 # low nibble = even x
    return Texture15(size, size, packed, format=abi.TEX_IDX4, clut=clut)


def scene_random(seed: int, n_tris: int, atlas_size: int, atlas_bits: int, name: str, **settings_kw) -> Scene:
    v, f, consumed = _random_triangles(seed, n_tris)
    tex = _atlas(seed, consumed, atlas_size, atlas_bits)
    return Scene(name, v, f, [tex], Camera(), common_settings(**settings_kw))


def scene_c2(n_tris: int = 1000, seed: int = SEED_C2, **kw) -> Scene:
    """1 000 random affine-textured triangles, 64x64 8-bit atlas, painter's. SURVEY.md §8d C2."""
    return scene_random(seed, n_tris, 64, 8, f"c2_{n_tris}_tris_64x64_idx8", **kw)


def scene_c4(n_tris: int = 100_000, seed: int = SEED_C4, **kw) -> Scene:
    """100 000-triangle stress scene, 256x256 4-bit atlas. SURVEY.md §8d C4."""
    return scene_random(seed, n_tris, 256, 4, f"c4_{n_tris}_tris_256x256_idx4", **kw)


def scene_c5(k: int, n_tris: int = 100_000, **kw) -> Scene:
    """Frame k of the 8 independent C4 frames. SURVEY.md §8d C5."""
    return scene_random(SEED_C5 + k, n_tris, 256, 4, f"c5_frame{k}_{n_tris}_tris", **kw)


def expand_texture(t: Texture15) -> Texture15:
    """IndexedAtlas::to_texture15 (mesh_editor.rs:669-682) on the host: what callers of the
    reference do before render_mesh_15. Used to check indexed == expanded."""
    if t.format == abi.TEX_RGB555:
        return t
    if t.format == abi.TEX_IDX8:
        idx = np.asarray(t.pixels, dtype=np.uint8)
    else:
        p = np.asarray(t.pixels, dtype=np.uint8)
        idx = np.empty(p.size * 2, dtype=np.uint8)
        idx[0::2] = p & 0xF
        idx[1::2] = p >> 4
        idx = idx[: t.width * t.height]
    clut = np.zeros(256, dtype=np.uint16)
    clut[: t.clut.size] = t.clut
    return Texture15(t.width, t.height, clut[idx], blend_mode=t.blend_mode)
