"""Host-side scene assembly for BASELINE config 3 (sample world scenes): the step *before* the hot
path.

Mirrors, in Python, the parts of the reference that turn a level file into the arguments of
`render_mesh_15`:
  * level files: brotli-compressed RON            src/world/level.rs:242-308
  * sector grid -> triangles                      src/world/geometry.rs:2839-3352
        Room::to_render_data_with_textures, add_horizontal_face_to_render_data,
        add_wall_to_render_data, add_diagonal_wall_to_render_data
  * texture packs: sorted packs / sorted PNGs     src/editor/texture_pack.rs:52-69, src/rasterizer/types.rs:1123-1168
        PNG -> Color (alpha 0 = Erase) -> Texture15   types.rs:1080-1111, 1267-1275
  * the game tab's texture resolver               src/game/renderer.rs:104-112
  * per-room ambient + fog, one call per room     src/scene.rs:180-261, 263-276
  * camera from the level's saved orbit           src/editor/state.rs:1129-1145

  * placed asset meshes and asset lights          src/scene.rs:32-107 (collect_scene_lights), :109-169 (render_asset_parts),
        :219-259 (the object loop of render_scene); AssetInstance::world_position geometry.rs:2353-2364;
        EditableMesh::to_render_data_textured + EditFace::triangulate src/modeler/mesh_editor.rs:1623-1653, :99-112;
        resolve_part_texture scene.rs:73-101, IndexedAtlas::to_texture15 mesh_editor.rs:669-682, checkerboard_clut :201-211

All float arithmetic is done in numpy float32 in the reference's operation order.  The room triangles are pinned bit
for bit against the reference's own compiled geometry code (tests/test_ref_wasm.py::test_sample_level_geometry_...).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import abi
from .raster import Camera, RasterSettings, Texture15

F = np.float32
SECTOR_SIZE = F(1024.0)            # src/world/geometry.rs (SECTOR_SIZE)


# ------------------------------------------------------------------------------------------------
# brotli + RON
# ------------------------------------------------------------------------------------------------
def brotli_decompress(data: bytes) -> bytes:
    lib = ctypes.CDLL(ctypes.util.find_library("brotlidec") or "libbrotlidec.so.1")
    lib.BrotliDecoderDecompress.restype = ctypes.c_int
    for mult in (64, 512, 4096):
        size = ctypes.c_size_t(len(data) * mult + (1 << 20))
        out = ctypes.create_string_buffer(size.value)
        if lib.BrotliDecoderDecompress(ctypes.c_size_t(len(data)), data, ctypes.byref(size), out) == 1:
            return out.raw[: size.value]
    raise ValueError("brotli decode failed")


class _Ron:
    """Minimal RON reader for the Level subset: structs `(a: 1, ...)`, tuples `(1, 2)`, lists, strings,
    numbers, booleans, `Some(x)` / `None`, bare enum identifiers and `Name(...)` variants."""

    def __init__(self, text: str):
        self.s = text
        self.i = 0

    def ws(self):
        s, n = self.s, len(self.s)
        while self.i < n:
            c = s[self.i]
            if c in " \t\r\n,":
                self.i += 1
            elif s.startswith("//", self.i):
                while self.i < n and s[self.i] != "\n":
                    self.i += 1
            else:
                break

    def ident(self) -> str:
        j = self.i
        while j < len(self.s) and (self.s[j].isalnum() or self.s[j] == "_"):
            j += 1
        out = self.s[self.i:j]
        self.i = j
        return out

    def value(self):
        self.ws()
        c = self.s[self.i]
        if c == "(":
            return self.paren()
        if c == "[":
            self.i += 1
            out = []
            while True:
                self.ws()
                if self.s[self.i] == "]":
                    self.i += 1
                    return out
                out.append(self.value())
        if c == '"':
            j = self.i + 1
            buf = []
            while self.s[j] != '"':
                if self.s[j] == "\\":
                    j += 1
                buf.append(self.s[j])
                j += 1
            self.i = j + 1
            return "".join(buf)
        if c.isdigit() or c in "-+.":
            j = self.i + 1
            while j < len(self.s) and (self.s[j].isalnum() or self.s[j] in ".-+_"):
                j += 1
            tok = self.s[self.i:j]
            self.i = j
            if any(ch in tok for ch in ".eE") or tok in ("inf", "-inf", "NaN"):
                return F(float(tok))          # Rust str::parse::<f32> rounds to nearest, as does float()->float32
            return int(tok)
        name = self.ident()
        if not name:
            raise ValueError(f"RON parse error at {self.i}: {self.s[self.i:self.i + 40]!r}")
        if name == "true":
            return True
        if name == "false":
            return False
        if name == "None":
            return None
        self.ws()
        if self.i < len(self.s) and self.s[self.i] == "(":
            inner = self.paren()
            if name == "Some":
                return inner[0] if isinstance(inner, tuple) and len(inner) == 1 else inner
            return {"__variant__": name, "value": inner}
        return name                                # bare enum variant

    def paren(self):
        assert self.s[self.i] == "("
        self.i += 1
        self.ws()
        # struct if the first token is `ident :`
        j = self.i
        while j < len(self.s) and (self.s[j].isalnum() or self.s[j] == "_"):
            j += 1
        k = j
        while k < len(self.s) and self.s[k] in " \t\r\n":
            k += 1
        if j > self.i and k < len(self.s) and self.s[k] == ":" and not self.s.startswith("::", k):
            out = {}
            while True:
                self.ws()
                if self.s[self.i] == ")":
                    self.i += 1
                    return out
                key = self.ident()
                self.ws()
                assert self.s[self.i] == ":", (key, self.s[self.i:self.i + 30])
                self.i += 1
                out[key] = self.value()
        items = []
        while True:
            self.ws()
            if self.s[self.i] == ")":
                self.i += 1
                return tuple(items)
            items.append(self.value())


def parse_ron(text: str):
    return _Ron(text).value()


def load_level_file(path: str) -> dict:
    """src/world/level.rs:242-271: plain RON if it starts with '(' / whitespace, else brotli."""
    raw = open(path, "rb").read()
    if raw[:1] not in (b"(", b" ", b"\n", b"\t", b"\r"):
        raw = brotli_decompress(raw)
    return parse_ron(raw.decode("utf-8"))


# ------------------------------------------------------------------------------------------------
# textures
# ------------------------------------------------------------------------------------------------
@dataclass
class PackTexture:
    name: str
    width: int
    height: int
    pixels15: np.ndarray        # u16[h*w]


def load_texture_packs(packs_dir: str) -> List[PackTexture]:
    """All packs (sorted by name), PNGs sorted by path, flattened: texture id = index (src/main.rs:495-499)."""
    from PIL import Image
    out: List[PackTexture] = []
    for pack in sorted(d for d in os.listdir(packs_dir) if os.path.isdir(os.path.join(packs_dir, d))):
        pdir = os.path.join(packs_dir, pack)
        for fn in sorted(f for f in os.listdir(pdir) if f.lower().endswith(".png")):
            try:
                img = Image.open(os.path.join(pdir, fn)).convert("RGBA")
            except Exception:
                continue
            a = np.asarray(img, dtype=np.uint8)
            h, w = a.shape[:2]
            r, g, b, al = (a[..., k].astype(np.uint16) for k in range(4))
            # Texture::from_file (alpha 0 -> Erase) then Texture::to_15 (types.rs:1267-1275)
            c15 = ((r >> 3) << 10) | ((g >> 3) << 5) | (b >> 3)
            c15 = np.where(al == 0, 0, c15).astype(np.uint16)
            out.append(PackTexture(os.path.splitext(fn)[0], w, h, c15.reshape(-1)))
    return out


# ------------------------------------------------------------------------------------------------
# geometry (src/world/geometry.rs:2839-3352)
# ------------------------------------------------------------------------------------------------
_BLEND = {"Opaque": 0, "Average": 1, "Add": 2, "Subtract": 3, "AddQuarter": 4, "Erase": 5}
_NEUTRAL = {"r": 128, "g": 128, "b": 128, "blend": "Opaque"}


def _color(c) -> tuple:
    return (int(c["r"]), int(c["g"]), int(c["b"]), _BLEND[c.get("blend", "Opaque")])


def _v2(t) -> np.ndarray:
    return np.array([F(t["x"]), F(t["y"])], dtype=F)


class _Builder:
    def __init__(self):
        self.pos, self.uv, self.nrm, self.col, self.faces = [], [], [], [], []

    def vertex(self, p, uv, n, c):
        self.pos.append(p); self.uv.append(uv); self.nrm.append(n); self.col.append(c)
        return len(self.pos) - 1

    def face(self, a, b, c, tex, black_tr, blend):
        self.faces.append((a, b, c, tex, black_tr, blend))

    def arrays(self):
        n = len(self.pos)
        v = np.zeros(n, dtype=abi.VERTEX_DTYPE)
        if n:
            v["pos"] = np.asarray(self.pos, dtype=F)
            v["uv"] = np.asarray(self.uv, dtype=F)
            v["normal"] = np.asarray(self.nrm, dtype=F)
            v["rgba"] = np.asarray(self.col, dtype=np.uint8)
        f = np.zeros(len(self.faces), dtype=abi.FACE_DTYPE)
        if self.faces:
            fa = np.asarray(self.faces, dtype=np.int64)
            f["v"] = fa[:, :3]
            f["flags"] = abi.face_flags(fa[:, 3], fa[:, 5], fa[:, 4].astype(bool), 255)
        return v, f


def _normalize(v):
    l = np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2])))
    if l == 0:
        return np.zeros(3, dtype=F)
    return np.array([v[0] / l, v[1] / l, v[2] / l], dtype=F)


def _cross(a, b):
    return np.array([F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0])], dtype=F)


def _vec(x, y, z):
    return np.array([F(x), F(y), F(z)], dtype=F)


def _horizontal(bld: _Builder, room_y, face: dict, base_x, base_z, gx, gz, is_floor, resolve):
    """add_horizontal_face_to_render_data, geometry.rs:2907-3052."""
    h1 = [F(x) for x in face["heights"]]
    h2 = [F(x) for x in face["heights_2"]] if face.get("heights_2") is not None else h1
    def corners(h):
        return [_vec(base_x, room_y + h[0], base_z), _vec(base_x + SECTOR_SIZE, room_y + h[1], base_z),
                _vec(base_x + SECTOR_SIZE, room_y + h[2], base_z + SECTOR_SIZE), _vec(base_x, room_y + h[3], base_z + SECTOR_SIZE)]
    c1, c2 = corners(h1), corners(h2)
    tex1 = face["texture"]
    tex2 = face.get("texture_2") or tex1
    tid1, tw1 = resolve(tex1) or (0, 64)
    tid2, tw2 = resolve(tex2) or (0, 64)
    s1, s2 = F(32.0) / F(tw1), F(32.0) / F(tw2)

    def default_uvs(s):
        uo, vo = F(gx) * s, F(gz) * s
        return [np.array([uo, vo], dtype=F), np.array([uo + s, vo], dtype=F), np.array([uo + s, vo + s], dtype=F), np.array([uo, vo + s], dtype=F)]
    uvs1 = [_v2(t) for t in face["uv"]] if face.get("uv") is not None else default_uvs(s1)
    uv2_src = face.get("uv_2") if face.get("uv_2") is not None else face.get("uv")
    if uv2_src is not None:
        uvs2 = [_v2(t) for t in uv2_src]
    else:
        uvs2 = uvs1 if tw1 == tw2 else default_uvs(s2)
    cols1 = [_color(c) for c in face.get("colors", (_NEUTRAL,) * 4)]
    cols2 = [_color(c) for c in face["colors_2"]] if face.get("colors_2") is not None else cols1
    mode = face.get("normal_mode", "Front")
    render_front, render_back = mode != "Back", mode != "Front"
    split = face.get("split_direction", "NwSe")
    t1c = (0, 1, 2) if split == "NwSe" else (0, 1, 3)
    t2c = (0, 2, 3) if split == "NwSe" else (1, 2, 3)

    def front_normal(c):
        e1, e2 = c[1] - c[0], c[3] - c[0]
        return _normalize(_cross(e2, e1)) if is_floor else _normalize(_cross(e1, e2))
    fn1, fn2 = front_normal(c1), front_normal(c2)
    black_tr = bool(face.get("black_transparent", True))
    blend = _BLEND[face.get("blend_mode", "Opaque")]

    def add_triangle(c, idx, uvs, cols, normal, tid, flip):
        i0 = bld.vertex(c[idx[0]], uvs[idx[0]], normal, cols[idx[0]])
        bld.vertex(c[idx[1]], uvs[idx[1]], normal, cols[idx[1]])
        bld.vertex(c[idx[2]], uvs[idx[2]], normal, cols[idx[2]])
        if flip:
            bld.face(i0, i0 + 2, i0 + 1, tid, black_tr, blend)
        else:
            bld.face(i0, i0 + 1, i0 + 2, tid, black_tr, blend)
    if render_front:
        add_triangle(c1, t1c, uvs1, cols1, fn1, tid1, not is_floor)
    if render_back:
        add_triangle(c1, t1c, uvs1, cols1, fn1 * F(-1.0), tid1, is_floor)
    if render_front:
        add_triangle(c2, t2c, uvs2, cols2, fn2, tid2, not is_floor)
    if render_back:
        add_triangle(c2, t2c, uvs2, cols2, fn2 * F(-1.0), tid2, is_floor)


def _quad(bld: _Builder, corners, front_normal, wall: dict, u_left, uv_scale, y_offset, resolve_tid):
    """Shared tail of add_wall_to_render_data / add_diagonal_wall_to_render_data (geometry.rs:3155-3223)."""
    u_right = u_left + uv_scale
    corner_u = [u_left, u_right, u_right, u_left]
    default = [np.array([corner_u[0], uv_scale], dtype=F), np.array([corner_u[1], uv_scale], dtype=F),
               np.array([corner_u[2], F(0.0)], dtype=F), np.array([corner_u[3], F(0.0)], dtype=F)]
    base_uvs = [_v2(t) for t in wall["uv"]] if wall.get("uv") is not None else default
    if wall.get("uv_projection", "Default") == "Projected":
        wh = [y_offset + F(h) for h in wall["heights"]]
        uvs = [np.array([base_uvs[i][0], -wh[i] / SECTOR_SIZE * uv_scale], dtype=F) for i in range(4)]
    else:
        uvs = base_uvs
    cols = [_color(c) for c in wall.get("colors", (_NEUTRAL,) * 4)]
    mode = wall.get("normal_mode", "Front")
    black_tr = bool(wall.get("black_transparent", True))
    blend = _BLEND[wall.get("blend_mode", "Opaque")]
    if mode != "Back":
        i0 = bld.vertex(corners[0], uvs[0], front_normal, cols[0])
        for i in range(1, 4):
            bld.vertex(corners[i], uvs[i], front_normal, cols[i])
        bld.face(i0, i0 + 2, i0 + 1, resolve_tid, black_tr, blend)
        bld.face(i0, i0 + 3, i0 + 2, resolve_tid, black_tr, blend)
    if mode != "Front":
        bn = front_normal * F(-1.0)
        i0 = bld.vertex(corners[0], uvs[0], bn, cols[0])
        for i in range(1, 4):
            bld.vertex(corners[i], uvs[i], bn, cols[i])
        bld.face(i0, i0 + 1, i0 + 2, resolve_tid, black_tr, blend)
        bld.face(i0, i0 + 2, i0 + 3, resolve_tid, black_tr, blend)


def _wall(bld, room_y, wall, bx, bz, gx, gz, direction, resolve):
    """add_wall_to_render_data, geometry.rs:3054-3224 (cardinal directions)."""
    h = [F(x) for x in wall["heights"]]
    y = room_y
    S = SECTOR_SIZE
    if direction == "North":
        corners = [_vec(bx, y + h[0], bz), _vec(bx + S, y + h[1], bz), _vec(bx + S, y + h[2], bz), _vec(bx, y + h[3], bz)]
        n = _vec(0.0, 0.0, 1.0)
    elif direction == "East":
        corners = [_vec(bx + S, y + h[0], bz), _vec(bx + S, y + h[1], bz + S), _vec(bx + S, y + h[2], bz + S), _vec(bx + S, y + h[3], bz)]
        n = _vec(-1.0, 0.0, 0.0)
    elif direction == "South":
        corners = [_vec(bx + S, y + h[0], bz + S), _vec(bx, y + h[1], bz + S), _vec(bx, y + h[2], bz + S), _vec(bx + S, y + h[3], bz + S)]
        n = _vec(0.0, 0.0, -1.0)
    else:  # West
        corners = [_vec(bx, y + h[0], bz + S), _vec(bx, y + h[1], bz), _vec(bx, y + h[2], bz), _vec(bx, y + h[3], bz + S)]
        n = _vec(1.0, 0.0, 0.0)
    tid, tw = resolve(wall["texture"]) or (0, 64)
    s = F(32.0) / F(tw)
    u_left = (F(gx) if direction in ("North", "South") else F(gz)) * s
    _quad(bld, corners, n, wall, u_left, s, y, tid)


def _diagonal(bld, room_y, wall, bx, bz, gx, is_nwse, resolve):
    """add_diagonal_wall_to_render_data, geometry.rs:3228-3352."""
    h = [F(x) for x in wall["heights"]]
    y = room_y
    S = SECTOR_SIZE
    nn = F(1.0) / np.sqrt(F(2.0))
    if is_nwse:
        corners = [_vec(bx + S, y + h[1], bz + S), _vec(bx, y + h[0], bz), _vec(bx, y + h[3], bz), _vec(bx + S, y + h[2], bz + S)]
        n = np.array([nn, F(0.0), -nn], dtype=F)
    else:
        corners = [_vec(bx, y + h[1], bz + S), _vec(bx + S, y + h[0], bz), _vec(bx + S, y + h[3], bz), _vec(bx, y + h[2], bz + S)]
        n = np.array([nn, F(0.0), nn], dtype=F)
    tid, tw = resolve(wall["texture"]) or (0, 64)
    s = F(32.0) / F(tw)
    _quad(bld, corners, n, wall, F(gx) * s, s, y, tid)


def room_to_render_data(room: dict, resolve):
    """Room::to_render_data_with_textures, geometry.rs:2839-2904."""
    bld = _Builder()
    px, py, pz = F(room["position"]["x"]), F(room["position"]["y"]), F(room["position"]["z"])
    for gx, col in enumerate(room["sectors"]):
        for gz, sector in enumerate(col):
            if sector is None:
                continue
            bx = px + F(gx) * SECTOR_SIZE
            bz = pz + F(gz) * SECTOR_SIZE
            if sector.get("floor") is not None:
                _horizontal(bld, py, sector["floor"], bx, bz, gx, gz, True, resolve)
            if sector.get("ceiling") is not None:
                _horizontal(bld, py, sector["ceiling"], bx, bz, gx, gz, False, resolve)
            for d, key in (("North", "walls_north"), ("East", "walls_east"), ("South", "walls_south"), ("West", "walls_west")):
                for w in sector.get(key, []) or []:
                    _wall(bld, py, w, bx, bz, gx, gz, d, resolve)
            for w in sector.get("walls_nwse", []) or []:
                _diagonal(bld, py, w, bx, bz, gx, True, resolve)
            for w in sector.get("walls_nesw", []) or []:
                _diagonal(bld, py, w, bx, bz, gx, False, resolve)
    return bld.arrays()


def build_room_fog(room: dict):
    """scene.rs:263-276. Returns None or (start, falloff, cull_distance, (r,g,b))."""
    fog = room.get("fog")
    if not fog or not fog.get("enabled", False):
        return None
    r, g, b = (F(c) for c in fog["color"])
    to_u8 = lambda c: int(min(max(np.trunc(F(c * F(255.0))), 0), 255))
    start = F(fog["start"])
    falloff = F(fog.get("falloff", fog.get("end", 30000.0)))
    cull = start + falloff + F(fog.get("cull_offset", 0.0))
    return (float(start), float(falloff), float(cull), (to_u8(r), to_u8(g), to_u8(b)))


def orbit_camera(level: dict) -> Camera:
    """EditorState::sync_camera_from_orbit, src/editor/state.rs:1129-1145, from the level's editor_layout."""
    lay = level.get("editor_layout") or {}
    tgt = _vec(lay.get("orbit_target_x", 512.0), lay.get("orbit_target_y", 512.0), lay.get("orbit_target_z", 512.0))
    dist = F(lay.get("orbit_distance", 4000.0))
    yaw = F(lay.get("orbit_azimuth", 0.8))
    pitch = F(lay.get("orbit_elevation", 0.4))
    fwd = np.array([F(np.cos(pitch) * np.sin(yaw)), F(-np.sin(pitch)), F(np.cos(pitch) * np.cos(yaw))], dtype=F)
    cam = Camera()
    cam.position = (tgt - fwd * dist).astype(F)
    cam.rotation_x, cam.rotation_y = pitch, yaw
    cam.update_basis()
    return cam


@dataclass
class RoomCall:
    """The arguments of one render_mesh_15 call of render_scene (scene.rs:196-217)."""
    vertices: np.ndarray
    faces: np.ndarray
    ambient: float
    fog: Optional[tuple]


@dataclass
class PartCall:
    """The arguments of one render_asset_parts iteration (scene.rs:131-168): one visible mesh part of a placed asset.
    vertices are LOCAL (the per-object rotation about Y by `facing` + translation to `world_pos` is part of the call)."""
    vertices: np.ndarray
    faces: np.ndarray             # texture ids already point into LevelScene.textures
    facing: float
    world_pos: tuple
    double_sided: bool
    ambient: float
    fog: Optional[tuple]


@dataclass
class LevelScene:
    name: str
    rooms: List[RoomCall]
    textures: List[Texture15]
    camera: Camera
    width: int = 320
    height: int = 240
    clear: tuple = (20, 22, 28)
    parts: List[PartCall] = field(default_factory=list)      # placed asset parts, drawn after all rooms (scene.rs:219-259)
    lights: list = field(default_factory=list)               # collect_scene_lights (scene.rs:32-70)

    def settings(self, ambient: float, **kw) -> RasterSettings:
        """RasterSettings::default() with backface_wireframe off; render_scene replaces lights (the point lights of
        placed assets) and ambient (room.ambient)."""
        s = RasterSettings(backface_wireframe=False, lights=list(self.lights), ambient=ambient)
        for k, v in kw.items():
            setattr(s, k, v)
        return s

    def part_settings(self, pc: "PartCall", **kw) -> RasterSettings:
        """render_asset_parts' per-part settings (scene.rs:138-143): double-sided parts are not culled."""
        s = self.settings(pc.ambient, **kw)
        s.backface_cull = (not pc.double_sided) and s.backface_cull
        s.backface_wireframe = (not pc.double_sided) and s.backface_wireframe
        return s


# ------------------------------------------------------------------------------------------------
# placed assets (src/scene.rs:32-169)
# ------------------------------------------------------------------------------------------------
def load_ron_dir(path: str) -> list:
    """Every brotli-compressed (or plain) .ron file of a directory, parsed, in sorted order."""
    out = []
    if not path or not os.path.isdir(path):
        return out
    for fn in sorted(os.listdir(path)):
        if not fn.endswith(".ron"):
            continue
        raw = open(os.path.join(path, fn), "rb").read()
        try:
            txt = brotli_decompress(raw).decode("utf-8")
        except ValueError:
            txt = raw.decode("utf-8")
        out.append(parse_ron(txt))
    return out


def _variant(c, name):
    return isinstance(c, dict) and c.get("__variant__") == name


def object_world_position(obj: dict, room: dict):
    """AssetInstance::world_position, geometry.rs:2353-2364 (the floor's average height is NOT offset by room.position.y
    there: restated as it is)."""
    pos = room["position"]
    half = SECTOR_SIZE * F(0.5)
    bx = F(pos["x"]) + F(obj["sector_x"]) * SECTOR_SIZE + half
    bz = F(pos["z"]) + F(obj["sector_z"]) * SECTOR_SIZE + half
    by = F(pos["y"])
    cols = room["sectors"]
    sx, sz = int(obj["sector_x"]), int(obj["sector_z"])
    if sx < len(cols) and sz < len(cols[sx]) and cols[sx][sz] is not None and cols[sx][sz].get("floor") is not None:
        h = [F(x) for x in cols[sx][sz]["floor"]["heights"]]
        by = (h[0] + h[1] + h[2] + h[3]) / F(4.0)                                      # HorizontalFace::avg_height :1262
    return (bx, by + F(obj.get("height", 0.0)), bz)


def collect_scene_lights(level: dict, assets: Dict[int, dict]) -> list:
    """scene.rs:32-70: the first Light component of every enabled placed asset, with per-instance overrides."""
    from .raster import Light
    out = []
    for room in level["rooms"]:
        for obj in room.get("objects", []) or []:
            if not obj.get("enabled", True):
                continue
            asset = assets.get(int(obj.get("asset_id", 0)))
            if asset is None:
                continue
            for comp in asset.get("components", []):
                if not _variant(comp, "Light"):
                    continue
                c = comp["value"]
                ov = (obj.get("overrides") or {}).get("light") or {}
                pick = lambda k: ov[k] if ov.get(k) is not None else c[k]
                color, intensity, radius, offset = pick("color"), pick("intensity"), pick("radius"), pick("offset")
                bp = object_world_position(obj, room)
                lp = (bp[0] + F(offset[0]), bp[1] + F(offset[1]), bp[2] + F(offset[2]))
                r, g, b = (F(int(color[k])) / F(255.0) for k in range(3))
                out.append(Light.point_colored(lp, float(F(radius)), float(F(intensity)), r, g, b))
                break
    return out


def _clut_lookup(palette, idx):
    """Clut::lookup, types.rs:390-397: out-of-range index -> 0x0000."""
    pal = np.zeros(256, np.uint16)
    pal[: min(len(palette), 256)] = np.asarray(palette[:256], np.uint16)
    valid = np.asarray(idx, np.int64) < len(palette)
    return np.where(valid, pal[np.asarray(idx, np.int64) & 0xFF], 0).astype(np.uint16)


def part_texture15(part: dict, user_textures: Dict[int, dict]) -> Texture15:
    """resolve_part_texture (scene.rs:73-101) + IndexedAtlas::to_texture15 (mesh_editor.rs:669-682)."""
    ref = part.get("texture_ref")
    if _variant(ref, "Id"):
        v = ref["value"]
        tid = int(v[0] if isinstance(v, tuple) else v)
        tex = user_textures.get(tid)
        if tex is not None:
            px = _clut_lookup(tex["palette"], tex["indices"])
            return Texture15(int(tex["width"]), int(tex["height"]), px)
    atlas = part.get("atlas") or {}
    w, h, idx = int(atlas.get("width", 0)), int(atlas.get("height", 0)), atlas.get("indices", [])
    checker = [(v << 10) | (v << 5) | v for v in (2 * i for i in range(16))]             # checkerboard_clut, mesh_editor.rs:201-211
    return Texture15(w, h, _clut_lookup(checker, idx) if len(idx) else np.zeros(0, np.uint16))


def part_render_data(part: dict, tex_index: int):
    """EditableMesh::to_render_data_textured (mesh_editor.rs:1623-1653): vertices as they are; n-gon faces as a fan from
    their first vertex (EditFace::triangulate :99-112); texture_id = the face's or Some(0) = the part's atlas, which is
    texture `tex_index` of the scene's table here (the reference passes a one-element table; other ids are out of range
    there = untextured)."""
    mesh = part["mesh"]
    mv = mesh["vertices"]
    v = np.zeros(len(mv), dtype=abi.VERTEX_DTYPE)
    blends = {"Opaque": 0, "Average": 1, "Add": 2, "Subtract": 3, "AddQuarter": 4, "Erase": 5}
    for i, e in enumerate(mv):
        v["pos"][i] = (F(e["pos"]["x"]), F(e["pos"]["y"]), F(e["pos"]["z"]))
        v["uv"][i] = (F(e["uv"]["x"]), F(e["uv"]["y"]))
        v["normal"][i] = (F(e["normal"]["x"]), F(e["normal"]["y"]), F(e["normal"]["z"]))
        c = e.get("color") or {"r": 128, "g": 128, "b": 128, "blend": "Opaque"}
        v["rgba"][i] = (int(c["r"]), int(c["g"]), int(c["b"]), blends.get(c.get("blend", "Opaque"), 0))
    tris, flags = [], []
    for f in mesh["faces"]:
        ids = [int(x) for x in f["vertices"]]
        n = len(ids)
        if n < 3:
            continue
        fan = [(ids[0], ids[1], ids[2])] if n == 3 else [(ids[0], ids[i], ids[i + 1]) for i in range(1, n - 1)]
        tid = f.get("texture_id")
        tid = tex_index if tid is None or int(tid) == 0 else abi.FACE_TEX_NONE
        fl = abi.face_flags(tid, blends.get(f.get("blend_mode", "Opaque"), 0), bool(f.get("black_transparent", True)), 255)
        for t in fan:
            tris.append(t)
            flags.append(fl)
    faces = np.zeros(len(tris), dtype=abi.FACE_DTYPE)
    if tris:
        faces["v"] = np.asarray(tris, np.uint32)
        faces["flags"] = np.asarray(flags, np.uint32)
    return v, faces


def assemble_parts(level: dict, assets: Dict[int, dict], user_textures: Dict[int, dict], first_tex: int):
    """The object loop of render_scene (scene.rs:219-259) -> [PartCall], [their textures]."""
    parts, texs = [], []
    for room in level["rooms"]:
        fog = build_room_fog(room)
        for obj in room.get("objects", []) or []:
            if not obj.get("enabled", True):
                continue
            asset = assets.get(int(obj.get("asset_id", 0)))
            if asset is None:
                continue
            mesh = next((c["value"] for c in asset.get("components", []) if _variant(c, "Mesh")), None)
            if mesh is None:
                continue
            wp = object_world_position(obj, room)
            for part in mesh.get("parts", []):
                if not part.get("visible", True):
                    continue
                v, f = part_render_data(part, first_tex + len(texs))
                if len(v) == 0:
                    continue
                texs.append(part_texture15(part, user_textures))
                parts.append(PartCall(v, f, float(F(obj.get("facing", 0.0))), tuple(float(x) for x in wp), bool(part.get("double_sided", False)),
                                      float(F(room.get("ambient", 0.5))), fog))
    return parts, texs


def assemble_level(level_path: str, all_textures: List[PackTexture], compact: bool = True, assets_dir: Optional[str] = None,
                   user_textures_dir: Optional[str] = None) -> LevelScene:
    """Level file -> per-room render_mesh_15 arguments, with the game tab's resolver
    (src/game/renderer.rs:104-112): invalid ref -> (0, 64); first texture whose name matches; miss -> None
    -> (0, 64) in the geometry code.  compact=True renumbers the textures actually used (ids are only
    indices; the rendered bytes do not depend on them) so fixtures stay small."""
    level = load_level_file(level_path)
    by_name: Dict[str, int] = {}
    for i, t in enumerate(all_textures):
        by_name.setdefault(t.name, i)

    def resolve(ref):
        if not ref or not ref.get("pack") or not ref.get("name"):
            return (0, 64)
        i = by_name.get(ref["name"])
        return None if i is None else (i, all_textures[i].width)

    rooms = []
    for room in level["rooms"]:
        v, f = room_to_render_data(room, resolve)
        if len(v) == 0:
            continue
        rooms.append(RoomCall(v, f, float(F(room.get("ambient", 0.5))), build_room_fog(room)))
    used = sorted({int(t) for rc in rooms for t in (rc.faces["flags"] & 0xFFFF)})
    if compact:
        remap = {t: k for k, t in enumerate(used)}
        for rc in rooms:
            ids = np.array([remap[int(t)] for t in (rc.faces["flags"] & 0xFFFF)], dtype=np.uint32)
            rc.faces["flags"] = (rc.faces["flags"] & ~np.uint32(0xFFFF)) | ids
        texs = [all_textures[t] for t in used]
    else:
        texs = all_textures
    textures = [Texture15(t.width, t.height, t.pixels15) for t in texs]
    # legacy objects (no asset_id) are dropped on load, as Level::load does (geometry.rs:2304-2306 "filtered out on load")
    for room in level["rooms"]:
        room["objects"] = [o for o in (room.get("objects") or []) if int(o.get("asset_id", 0)) != 0]
    assets = {int(a["id"]): a for a in load_ron_dir(assets_dir)}
    utex = {int(t["id"]): t for t in load_ron_dir(user_textures_dir)}
    parts, part_texs = assemble_parts(level, assets, utex, len(textures))
    scene = LevelScene(os.path.splitext(os.path.basename(level_path))[0], rooms, textures + part_texs, orbit_camera(level))
    scene.parts = parts
    scene.lights = collect_scene_lights(level, assets)
    return scene


class LevelRenderer:
    """render_scene's room loop (src/scene.rs:180-261) on device-resident geometry.

    The reference regenerates `room.to_render_data_with_textures()` and re-marshals it on every frame
    (scene.rs:199-203).  Here a room's triangles are uploaded once per level *generation* (bump `generation` when the
    level is edited; cf. `textures_15_cache_generation`, src/editor/viewport_3d.rs:3459) and every frame only enqueues
    one frame-graph launch per room (both passes: rooms with water / glass stay enqueued too) and one enqueued placed
    render per asset part — no host round trip until the framebuffer is downloaded.
    """

    def __init__(self, ctx, scene: LevelScene):
        from .raster import Mesh
        self.ctx = ctx
        self.scene = scene
        self.generation = 0
        self._uploaded = -1
        self._meshes: list = []
        self._parts: list = []
        self._Mesh = Mesh

    def _sync_geometry(self):
        if self._uploaded == self.generation:
            return
        for m in self._meshes + self._parts:
            m.free()
        self.ctx.set_textures(self.scene.textures)
        self._meshes = [self._Mesh(self.ctx, rc.vertices, rc.faces) for rc in self.scene.rooms]
        self._parts = [self._Mesh(self.ctx, pc.vertices, pc.faces) for pc in self.scene.parts]
        self._uploaded = self.generation

    def render(self, fb, camera: Optional[Camera] = None, clear=True, **settings_kw):
        """fb.clear + one render_mesh_15 per room, enqueued; the caller reads the frame with fb.download()."""
        self._sync_geometry()
        cam = camera or self.scene.camera
        for i, (mesh, rc) in enumerate(zip(self._meshes, self.scene.rooms)):
            mesh.frame_enqueue(self.scene.clear if (clear and i == 0) else None, cam, self.scene.settings(rc.ambient, **settings_kw), rc.fog)
        if clear and not self._meshes:
            fb.clear(self.scene.clear)
        # placed asset parts (scene.rs:219-259): resident, transformed per object on the device, enqueued behind the rooms
        for mesh, pc in zip(self._parts, self.scene.parts):
            mesh.render_placed(cam, self.scene.part_settings(pc, **settings_kw), pc.facing, pc.world_pos, pc.fog, enqueue_only=True)

    def close(self):
        for m in self._meshes + self._parts:
            m.free()
        self._parts = []
        self._meshes = []
        self._uploaded = -1
