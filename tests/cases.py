"""Named parity scenes shared by the CPU tests (oracle vs pymodel vs golden) and the GPU tests
(CUDA vs oracle).  Each covers one feature of render_mesh_15 / rasterize_triangle_15."""
from __future__ import annotations

import copy

import numpy as np

from bonnie32_b200 import abi, scenes
from bonnie32_b200.raster import Camera, Light, RasterSettings, Texture, Texture15


def _rng_texture(seed, w, h, blend=abi.BLEND_OPAQUE, semi_fraction=0.0, zero_fraction=0.05):
    u = scenes.splitmix64_u01(seed, w * h * 3)
    px = np.floor(u[: w * h] * 32768.0).astype(np.uint16)
    px[u[w * h: 2 * w * h] < zero_fraction] = 0
    semi = u[2 * w * h:] < semi_fraction
    px[semi] |= 0x8000
    return Texture15(w, h, px, blend_mode=blend)


def _with(scene, name, **kw):
    s = copy.copy(scene)
    s.settings = copy.copy(scene.settings)
    s.name = name
    for k, v in kw.items():
        if hasattr(s.settings, k):
            setattr(s.settings, k, v)
        else:
            setattr(s, k, v)
    return s


def _rotated_camera(rx, ry, pos):
    c = Camera()
    c.rotation_x = np.float32(rx)
    c.rotation_y = np.float32(ry)
    c.update_basis()
    c.position = np.asarray(pos, dtype=np.float32)
    return c


def _mixed_faces(scene, seed):
    """Give the faces a mix of blend modes / black_transparent / editor_alpha / texture ids."""
    n = len(scene.faces)
    u = scenes.splitmix64_u01(seed, n * 4).reshape(n, 4)
    blend = np.where(u[:, 0] < 0.5, 0, np.floor(u[:, 0] * 10).astype(np.int64) - 4)   # 0 or 1..5
    blend = np.clip(blend, 0, 5)
    black_tr = u[:, 1] < 0.6
    ea = np.where(u[:, 2] < 0.7, 255, np.floor(u[:, 2] * 256).astype(np.int64))
    ea = np.where(u[:, 2] > 0.98, 0, ea)
    tex = np.floor(u[:, 3] * 4).astype(np.int64)          # 0..3 ; 3 = out of range -> untextured
    tex = np.where(u[:, 3] > 0.9, abi.FACE_TEX_NONE, tex)
    f = scene.faces.copy()
    f["flags"] = abi.face_flags(tex, blend, black_tr, ea)
    return f


def feature_scenes(n_tris=160):
    """A list of small scenes, one per feature; all 320x240 unless stated."""
    out = []
    base = scenes.scene_c2(n_tris=n_tris)
    out.append(_with(base, "painter_idx8"))
    out.append(_with(base, "zbuffer_idx8", use_zbuffer=True))
    out.append(_with(base, "nodither", dithering=False))
    out.append(_with(base, "float_projection", use_fixed_point=False))
    out.append(_with(base, "perspective_correct", affine_textures=False))
    out.append(_with(base, "nocull_painter", backface_cull=False))
    out.append(_with(base, "nocull_zbuffer", backface_cull=False, use_zbuffer=True))
    out.append(_with(base, "xray", xray_mode=True, use_zbuffer=True))
    out.append(_with(base, "ortho", ortho_projection=(6.0, 0.5, -0.25), use_zbuffer=True))
    out.append(_with(base, "fb_200x150", width=200, height=150))
    out.append(_with(base, "fb_640x480_zbuffer", width=640, height=480, use_zbuffer=True))

    lights = [Light.directional((-1.0, -1.0, -1.0), 0.7),
              Light.point_colored((0.5, 0.5, 4.0), 30.0, 1.5, 1.0, 0.5, 0.25),
              Light.point((-3.0, 2.0, 20.0), 25.0, 0.9)]
    off = Light.point((0.0, 0.0, 10.0), 50.0, 2.0)
    off.enabled = False
    # random (non-unit) normals so that lighting varies per vertex
    g = copy.copy(base)
    g.vertices = base.vertices.copy()
    un = scenes.splitmix64_u01(77, len(g.vertices) * 3).reshape(-1, 3)
    g.vertices["normal"] = (2.0 * un - 1.0).astype(np.float32)
    out.append(_with(g, "gouraud_lights", shading=abi.SHADE_GOURAUD, lights=lights + [off], ambient=0.3, use_zbuffer=True))
    out.append(_with(g, "flat_lights", shading=abi.SHADE_FLAT, lights=lights, ambient=0.2))
    out.append(_with(g, "gouraud_default_nocull", shading=abi.SHADE_GOURAUD,
                     lights=[Light.directional((-1.0, -1.0, -1.0), 0.7)], ambient=0.3, backface_cull=False, use_zbuffer=True))

    out.append(_with(base, "fog", fog=(10.0, 30.0, 50.0, (90, 110, 130)), use_zbuffer=True))
    out.append(_with(base, "fog_nofalloff", fog=(20.0, 0.0, 45.0, (10, 200, 30, abi.BLEND_ADD))))

    # several textures, mixed face flags: transparent pass, all blend modes, editor alpha
    m = copy.copy(base)
    m.textures = [_rng_texture(11, 64, 64, semi_fraction=0.5),
                  _rng_texture(12, 32, 64, blend=abi.BLEND_AVERAGE, semi_fraction=0.5),
                  _rng_texture(13, 16, 8, blend=abi.BLEND_ADD, semi_fraction=0.9, zero_fraction=0.3)]
    m.faces = _mixed_faces(base, 99)
    out.append(_with(m, "mixed_painter"))
    out.append(_with(m, "mixed_zbuffer", use_zbuffer=True))
    out.append(_with(m, "mixed_zbuffer_nocull_gouraud", use_zbuffer=True, backface_cull=False,
                     shading=abi.SHADE_GOURAUD, lights=lights, ambient=0.4))
    out.append(_with(m, "mixed_xray", xray_mode=True))

    # untextured vertex-coloured triangles (needs_dither from colour inequality)
    ut = copy.copy(base)
    ut.faces = base.faces.copy()
    ut.faces["flags"] = abi.face_flags(abi.FACE_TEX_NONE)
    out.append(_with(ut, "untextured_vertex_colours"))

    # camera inside the cloud, rotated: near-plane rejects, huge off-screen coordinates (slow edge path)
    big = scenes.scene_c2(n_tris=n_tris, seed=0xB32000AA)
    big.vertices = big.vertices.copy()
    big.vertices["pos"] *= np.float32(40.0)
    out.append(_with(big, "rotated_camera_large_world", camera=_rotated_camera(0.3, 0.7, (10.0, -20.0, 300.0)), use_zbuffer=True))
    out.append(_with(big, "rotated_camera_large_world_float", camera=_rotated_camera(-0.2, 2.4, (-30.0, 15.0, 900.0)),
                     use_fixed_point=False))
    return out


def big_triangle_scene():
    """Few very large triangles that cover the whole screen with vertices far off-screen."""
    pos = [(-4000.0, -3000.0, 60.0), (5000.0, -2500.0, 30.0), (100.0, 4000.0, 2.0),
           (-30.0, -20.0, 8.0), (40.0, -25.0, 9.0), (5.0, 35.0, 1.0),
           (-900.5, 700.25, 3.0), (800.75, 650.5, 2.5), (10.25, -1200.125, 40.0)]
    uv = [(0, 0), (3, 0), (0, 3)] * 3
    rgba = [(255, 128, 64, 0), (64, 255, 128, 0), (128, 64, 255, 0)] * 3
    v = scenes.make_vertices(pos, uv=uv, normal=[(0, 0, -1)] * 9, rgba=rgba)
    f = scenes.make_faces([(0, 1, 2), (0, 2, 1), (3, 4, 5), (3, 5, 4), (6, 7, 8), (6, 8, 7)], tex_id=0)
    tex = _rng_texture(5, 64, 64)
    s = scenes.common_settings(use_zbuffer=True, backface_cull=False)
    return scenes.Scene("big_triangles", v, f, [tex], Camera(), s)


def nan_depth_vertices(scene, oracle):
    """A copy of scene.vertices where one drawn face got a NaN depth that survives culling, so the
    reference's `partial_cmp().unwrap()` (render.rs:2531) panics in painter's mode."""
    import dataclasses
    from bonnie32_b200 import abi
    for fi in range(len(scene.faces)):
        v = scene.vertices.copy()
        v["pos"][scene.faces["v"][fi, 0], 2] = np.nan
        if oracle.render_scene(dataclasses.replace(scene, vertices=v))[3] == abi.B32_ERR_NAN_DEPTH:
            return v
    raise AssertionError("no face produces a NaN sort key")


def grid_mesh_scene(nx=40, ny=30, name="wire_grid_shared_edges", flip=True, **kw):
    """A wavy quad grid with shared vertices seen from behind (every triangle is a back face): each interior edge occurs
    in two triangles, and many snap to the same integer end points — the case the first-occurrence de-duplication of the
    wireframe phase (render.rs:2587-2591) exists for."""
    xs, ys = np.meshgrid(np.linspace(-6.0, 6.0, nx + 1), np.linspace(-4.5, 4.5, ny + 1))
    zs = 14.0 + 1.5 * np.sin(xs * 0.9) * np.cos(ys * 1.1)
    pos = np.stack([xs, ys, zs], axis=-1).reshape(-1, 3)
    u = scenes.splitmix64_u01(4242, len(pos) * 3).reshape(-1, 3)
    v = scenes.make_vertices(pos, uv=pos[:, :2] * 0.2, normal=np.tile([0.0, 0.0, -1.0], (len(pos), 1)),
                             rgba=np.concatenate([64 + np.floor(128 * u), np.zeros((len(pos), 1))], axis=1))
    idx = []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = j * (nx + 1) + i, j * (nx + 1) + i + 1, (j + 1) * (nx + 1) + i, (j + 1) * (nx + 1) + i + 1
            idx += [(a, c, b), (b, c, d)] if flip else [(a, b, c), (b, d, c)]
    base = scenes.scene_c2(n_tris=8)
    f = scenes.make_faces(idx, tex_id=0)
    st = scenes.common_settings(use_zbuffer=True, backface_wireframe=True, **kw)
    return scenes.Scene(name, v, f, base.textures, Camera(), st)


def wireframe_scenes(n_tris=160):
    """Editor wireframe phase (render.rs:2574-2635): back-face edges with depth test, front-face overlay."""
    base = scenes.scene_c2(n_tris=n_tris)
    big = scenes.scene_c2(n_tris=n_tris, seed=0xB32000AA)
    big.vertices = big.vertices.copy()
    big.vertices["pos"] *= np.float32(40.0)
    cam = _rotated_camera(0.3, 0.7, (10.0, -20.0, 300.0))
    return [
        _with(base, "wire_backface_zbuffer", backface_wireframe=True, use_zbuffer=True),
        _with(base, "wire_backface_painter", backface_wireframe=True),
        _with(base, "wire_overlay", wireframe_overlay=True),
        _with(base, "wire_overlay_and_backface", wireframe_overlay=True, backface_wireframe=True, use_zbuffer=True),
        _with(base, "wire_backface_nocull_is_off", backface_wireframe=True, backface_cull=False, use_zbuffer=True),
        _with(base, "wire_backface_xray", backface_wireframe=True, xray_mode=True),
        _with(big, "wire_backface_large_world", camera=cam, backface_wireframe=True, use_zbuffer=True),
        grid_mesh_scene(),
        grid_mesh_scene(name="wire_grid_overlay_nocull", flip=False, wireframe_overlay=True, backface_cull=False),
    ]


# ---- RGB888 sibling: render_mesh / rasterize_triangle (render.rs:1971-2259, 1202-1433) --------------------
def _rng_texture8(seed, w, h, erase_fraction=0.1, blend_fraction=0.0, name=""):
    """Random Color texels; blend tag Opaque, Erase (transparent) or one of the four PS1 blend modes."""
    u = scenes.splitmix64_u01(seed, w * h * 5).reshape(5, w * h)
    px = np.zeros((w * h, 4), dtype=np.uint8)
    px[:, :3] = np.floor(u[:3].T * 256.0).astype(np.uint8)
    tag = np.zeros(w * h, dtype=np.uint8)
    tag[u[3] < erase_fraction] = abi.BLEND_ERASE
    blended = u[3] > 1.0 - blend_fraction
    tag[blended] = (1 + np.floor(u[4][blended] * 4.0)).astype(np.uint8)      # Average, Add, Subtract, AddQuarter
    px[:, 3] = tag
    return Texture(w, h, px.reshape(-1), name=name)


def rgb888_scenes(n_tris=160):
    """Scenes for render_mesh: opaque-only texels (order-free on the device) and per-texel blend tags /
    editor alpha (strict draw-order replay), painter's and z-buffer, lights, x-ray, ortho, no-cull."""
    out = []
    base = scenes.scene_c2(n_tris=n_tris)
    base.settings = copy.copy(base.settings)
    base.settings.use_rgb555 = False
    opaque_tex = [_rng_texture8(21, 64, 64)]
    blend_tex = [_rng_texture8(22, 64, 64, blend_fraction=0.4), _rng_texture8(23, 32, 16, erase_fraction=0.3, blend_fraction=0.6),
                 _rng_texture8(24, 8, 8, erase_fraction=0.0)]
    o = _with(base, "rgb888_opaque_painter", textures8=opaque_tex)
    out.append(o)
    out.append(_with(o, "rgb888_opaque_zbuffer", use_zbuffer=True))
    out.append(_with(o, "rgb888_opaque_nodither", dithering=False))
    out.append(_with(o, "rgb888_opaque_float_perspective", use_fixed_point=False, affine_textures=False, use_zbuffer=True))
    out.append(_with(o, "rgb888_opaque_nocull_xray_zbuffer", backface_cull=False, xray_mode=True, use_zbuffer=True))
    out.append(_with(o, "rgb888_opaque_xray_painter", xray_mode=True))
    out.append(_with(o, "rgb888_opaque_ortho", ortho_projection=(6.0, 0.5, -0.25), use_zbuffer=True))
    out.append(_with(o, "rgb888_opaque_fb_200x150", width=200, height=150))
    lights = [Light.directional((-1.0, -1.0, -1.0), 0.7), Light.point_colored((0.5, 0.5, 4.0), 30.0, 1.5, 1.0, 0.5, 0.25),
              Light.point((-3.0, 2.0, 20.0), 25.0, 2.9)]
    g = copy.copy(o)
    g.vertices = base.vertices.copy()
    un = scenes.splitmix64_u01(78, len(g.vertices) * 3).reshape(-1, 3)
    g.vertices["normal"] = (2.0 * un - 1.0).astype(np.float32)
    out.append(_with(g, "rgb888_gouraud_lights", shading=abi.SHADE_GOURAUD, lights=lights, ambient=0.3, use_zbuffer=True))
    out.append(_with(g, "rgb888_flat_lights", shading=abi.SHADE_FLAT, lights=lights, ambient=0.2))
    ut = copy.copy(o)
    ut.faces = base.faces.copy()
    ut.faces["flags"] = abi.face_flags(abi.FACE_TEX_NONE)
    out.append(_with(ut, "rgb888_untextured_vertex_colours"))
    # ordered replay: blended texels, editor alpha, several textures (id 3 is out of range -> untextured)
    m = copy.copy(base)
    m.faces = _mixed_faces(base, 101)
    m = _with(m, "rgb888_mixed_painter", textures8=blend_tex)
    out.append(m)
    out.append(_with(m, "rgb888_mixed_zbuffer", use_zbuffer=True))
    out.append(_with(m, "rgb888_mixed_zbuffer_xray_nocull", use_zbuffer=True, xray_mode=True, backface_cull=False))
    out.append(_with(m, "rgb888_mixed_gouraud", shading=abi.SHADE_GOURAUD, lights=lights, ambient=0.4, use_zbuffer=True))
    ea = copy.copy(o)                      # opaque texels, but some faces have editor alpha < 255
    ea.faces = _mixed_faces(base, 102)
    ea.faces["flags"] = (ea.faces["flags"] & ~np.uint32(0xFFFF)) | np.uint32(0)
    out.append(_with(ea, "rgb888_editor_alpha_zbuffer", use_zbuffer=True))
    out.append(_with(m, "rgb888_wire_backface", backface_wireframe=True, use_zbuffer=True))
    return out


def spot_lights():
    """Spot lights (render.rs:1038-1059) over the C2 volume: a torch at the camera, a coloured one from the side with a
    wide cone, a narrow one whose direction is not normalised (|dot| > 1 at the axis: acos = NaN, which the reference
    lets through its `spot_angle > angle` test), one with a zero cone, one disabled."""
    torch = Light.spot((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 0.6, 70.0, 1.8)
    side = Light.spot((-30.0, 10.0, 30.0), (0.8, -0.2, 0.1), 1.3, 80.0, 2.5)
    side.color = (255, 120, 40)
    unnorm = Light.spot((5.0, -20.0, 10.0), (-0.2, 1.1, 1.3), 2.0, 90.0, 0.7)
    unnorm.direction = np.asarray((-0.2, 1.1, 1.3), dtype=np.float32)          # Light::spot normalises; the field is public
    zero_cone = Light.spot((0.0, 0.0, 5.0), (0.0, 0.0, 1.0), 0.0, 50.0, 3.0)
    off = Light.spot((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 3.0, 500.0, 9.0)
    off.enabled = False
    return [torch, side, unnorm, zero_cone, off]


def spot_scenes(n_tris=160):
    """Scenes lit by Spot lights, alone and mixed with Directional / Point ones: Gouraud and flat, painter's and z-buffer,
    both colour depths, no-cull (flipped normals), a mixed-blend mesh.  Kept apart from feature_scenes() so that the
    fixtures made from that list stay what they are."""
    import fuzz
    out = []
    base = scenes.scene_c2(n_tris=n_tris)
    g = copy.copy(base)
    g.vertices = base.vertices.copy()
    un = scenes.splitmix64_u01(79, len(g.vertices) * 3).reshape(-1, 3)
    g.vertices["normal"] = (2.0 * un - 1.0).astype(np.float32)
    sp = spot_lights()
    mixed = [Light.directional((-1.0, -1.0, -1.0), 0.4), sp[0], Light.point_colored((0.5, 0.5, 4.0), 30.0, 1.5, 1.0, 0.5, 0.25), sp[1], sp[4]]
    out.append(_with(g, "spot_gouraud_torch", shading=abi.SHADE_GOURAUD, lights=[sp[0]], ambient=0.15, use_zbuffer=True))
    out.append(_with(g, "spot_gouraud_all", shading=abi.SHADE_GOURAUD, lights=sp, ambient=0.1))
    out.append(_with(g, "spot_flat_all", shading=abi.SHADE_FLAT, lights=sp, ambient=0.2, use_zbuffer=True))
    out.append(_with(g, "spot_gouraud_mixed_nocull", shading=abi.SHADE_GOURAUD, lights=mixed, ambient=0.25, backface_cull=False, use_zbuffer=True))
    out.append(_with(g, "spot_flat_mixed_float_nodither", shading=abi.SHADE_FLAT, lights=mixed, ambient=0.0, use_fixed_point=False, dithering=False))
    out.append(_with(g, "spot_shading_none_is_ignored", shading=abi.SHADE_NONE, lights=sp, ambient=0.3))
    m = copy.copy(g)
    m.textures = [_rng_texture(11, 64, 64, semi_fraction=0.5), _rng_texture(12, 32, 64, blend=abi.BLEND_AVERAGE, semi_fraction=0.5),
                  _rng_texture(13, 16, 8, blend=abi.BLEND_ADD, semi_fraction=0.9, zero_fraction=0.3)]
    m.faces = _mixed_faces(base, 99)
    out.append(_with(m, "spot_mixed_blend_gouraud", shading=abi.SHADE_GOURAUD, lights=mixed, ambient=0.3, use_zbuffer=True))
    for seed in (3, 5, 8, 11, 17, 23, 30, 36, 41, 52):                     # fuzz scenes, their own lights replaced
        sc = fuzz.fuzz_scene(seed)
        rng = np.random.default_rng(880000 + seed)
        ls = list(sc.settings.lights)
        for _ in range(int(rng.integers(1, 4))):
            l = Light.spot(rng.normal(size=3) * 8 + np.array([0, 0, 10.0]), rng.normal(size=3),
                           float(rng.random() * 2.5), float(10 + rng.random() * 80), float(rng.random() * 3))
            if rng.random() < 0.3:
                l.direction = (l.direction * np.float32(1.0 + rng.random() * 1e-6)).astype(np.float32)   # a hair over unit length
            l.color = tuple(int(x) for x in rng.integers(0, 256, size=3))
            ls.insert(int(rng.integers(0, len(ls) + 1)), l)
        out.append(_with(sc, f"spot_fuzz_{seed}", lights=ls, shading=int(abi.SHADE_GOURAUD if seed % 2 else abi.SHADE_FLAT)))
    return out


def spot_scenes888(n_tris=160):
    """The same lights through the RGB888 sibling `render_mesh`."""
    by = {s.name: s for s in rgb888_scenes(n_tris)}
    sp = spot_lights()
    mixed = [Light.point((-3.0, 2.0, 20.0), 25.0, 2.9), sp[1], sp[0], sp[3]]
    return [_with(by["rgb888_gouraud_lights"], "rgb888_spot_gouraud", lights=sp, ambient=0.2),
            _with(by["rgb888_flat_lights"], "rgb888_spot_flat_mixed", lights=mixed, ambient=0.1, use_zbuffer=True),
            _with(by["rgb888_mixed_gouraud"], "rgb888_spot_mixed_blend", lights=mixed, ambient=0.3)]


# ---- skybox sphere pass (Framebuffer::render_skybox step 1, render.rs:81-139) ------------------------------
def sky_mesh(center, seed=5, h_segments=48, v_segments=32, radius=10000.0, n_mountains=40):
    """A mesh shaped like Skybox::generate_mesh's (src/world/geometry.rs:529-): a vertex-coloured sphere around `center`
    wound for inside viewing, plus a ring of peaked 'mountain' triangles slightly inside it, drawn after the sphere."""
    u = scenes.splitmix64_u01(seed, (v_segments + 1) * (h_segments + 1) * 3 + n_mountains * 16)
    k = 0
    verts = []
    for v in range(v_segments + 1):
        phi = np.pi * v / v_segments
        for h in range(h_segments + 1):
            theta = 2.0 * np.pi * h / h_segments
            d = (np.sin(phi) * np.cos(theta), np.cos(phi), np.sin(phi) * np.sin(theta))
            col = tuple(int(u[k + c] * 256.0) for c in range(3)); k += 3
            verts.append((tuple(center[c] + d[c] * radius for c in range(3)), col))
    faces = []
    rw = h_segments + 1
    for v in range(v_segments):
        for h in range(h_segments):
            i0, i1, i2, i3 = v * rw + h, v * rw + h + 1, (v + 1) * rw + h, (v + 1) * rw + h + 1
            faces += [(i0, i2, i1), (i1, i2, i3)]
    r2 = radius * 0.97
    for m in range(n_mountains):
        t0 = 2.0 * np.pi * (m + u[k]) / n_mountains; wdt = 0.05 + 0.1 * u[k + 1]; hgt = 0.05 + 0.25 * u[k + 2]
        base_phi = np.pi * 0.5 + 0.08
        pts = [(t0 - wdt, base_phi), (t0 + wdt, base_phi), (t0, base_phi - hgt)]
        b = len(verts)
        for j, (th, ph) in enumerate(pts):
            d = (np.sin(ph) * np.cos(th), np.cos(ph), np.sin(ph) * np.sin(th))
            col = tuple(int(u[k + 3 + 3 * j + c] * 200.0) for c in range(3))
            verts.append((tuple(center[c] + d[c] * r2 for c in range(3)), col))
        k += 16
        faces += [(b, b + 2, b + 1), (b, b + 1, b + 2)]          # both windings: one of them faces inward
    sv = np.zeros(len(verts), dtype=abi.SKY_VERTEX_DTYPE)
    sv["pos"] = np.array([p for p, _ in verts], dtype=np.float32)
    sv["rgb"] = np.array([c for _, c in verts], dtype=np.uint8)
    return sv, np.array(faces, dtype=np.uint32)


def sky_cases():
    """(name, width, height, camera) for the skybox pass; the mesh is centred on the camera as in the reference."""
    out = []
    for name, w, h, rx, ry, pos in [("sky_level", 320, 240, 0.0, 0.0, (0.0, 0.0, 0.0)), ("sky_look_up", 320, 240, -0.9, 0.4, (3.0, 2.0, -1.0)),
                                    ("sky_look_down_turned", 320, 240, 0.7, 2.5, (100.0, -30.0, 250.0)), ("sky_640x480", 640, 480, 0.2, -1.3, (0.0, 5.0, 0.0)),
                                    ("sky_straight_up", 200, 150, -1.5707964, 0.0, (0.0, 0.0, 0.0))]:
        out.append((name, w, h, _rotated_camera(rx, ry, pos)))
    return out


# ---- overlay lines (Framebuffer::draw_line*, render.rs:684-872) ---------------------------------------------
def line_background(w, h, seed):
    """A framebuffer to draw over: random colours and a z-buffer of a few depth planes with f32::MAX holes, so that
    3D lines hit `<`, `==` and `>` against it."""
    rng = np.random.default_rng(seed)
    rgba = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    rgba[..., 3] = 255
    z = np.full((h, w), np.finfo(np.float32).max, dtype=np.float32)
    for _ in range(6):
        x0, x1 = sorted(rng.integers(0, w + 1, 2)); y0, y1 = sorted(rng.integers(0, h + 1, 2))
        z[y0:y1, x0:x1] = np.float32(rng.choice([4.0, 8.0, 8.0, 12.5, 20.0]))
    return rgba, z


def random_lines(w, h, n, seed, kinds=(0, 1, 2, 3, 4), spread=1.3):
    """n lines with end points in a box `spread` x the screen (so many leave it), every kind / mode / alpha, depths
    on and around the planes of line_background."""
    from bonnie32_b200 import abi
    rng = np.random.default_rng(seed)
    ln = np.zeros(n, dtype=abi.LINE_DTYPE)
    cx, cy = w / 2, h / 2
    for f, c, e in (("x0", cx, w), ("x1", cx, w), ("y0", cy, h), ("y1", cy, h)):
        ln[f] = (c + (rng.random(n) - 0.5) * e * spread).astype(np.int32)
    flat = rng.random(n) < 0.15                       # horizontals, verticals, points
    ln["y1"][flat] = ln["y0"][flat]
    vert = rng.random(n) < 0.1
    ln["x1"][vert] = ln["x0"][vert]
    ln["z0"] = rng.choice(np.array([2.0, 4.0, 8.0, 8.0, 12.5, 30.0], dtype=np.float32), n)
    ln["z1"] = np.where(rng.random(n) < 0.5, ln["z0"], rng.choice(np.array([3.0, 8.0, 16.0], dtype=np.float32), n))
    ln["rgb"] = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    ln["blend"] = rng.choice(np.array([0, 0, 0, 1, 5], dtype=np.uint8), n)
    ln["kind"] = rng.choice(np.array(kinds, dtype=np.uint8), n)
    ln["mode"] = rng.integers(0, 6, n).astype(np.uint8)
    ln["alpha"] = rng.choice(np.array([0, 1, 64, 128, 200, 254, 255], dtype=np.uint8), n)
    return ln


def star_lines(w, h, n, kind, mode=0, alpha=128):
    """n lines through the screen centre: the deepest stack of operations on one pixel."""
    from bonnie32_b200 import abi
    ln = np.zeros(n, dtype=abi.LINE_DTYPE)
    ang = np.arange(n) * (np.pi / n)
    r = max(w, h)
    ln["x0"] = (w // 2 + np.round(np.cos(ang) * r)).astype(np.int32); ln["x1"] = (w // 2 - np.round(np.cos(ang) * r)).astype(np.int32)
    ln["y0"] = (h // 2 + np.round(np.sin(ang) * r)).astype(np.int32); ln["y1"] = (h // 2 - np.round(np.sin(ang) * r)).astype(np.int32)
    ln["z0"] = 8.0; ln["z1"] = 8.0
    ln["rgb"] = (np.arange(n)[:, None] * np.array([37, 91, 13]) % 256).astype(np.uint8)
    ln["kind"] = kind; ln["mode"] = mode; ln["alpha"] = alpha
    return ln


def line_cases():
    """(name, width, height, background seed, lines)."""
    out = [("lines_opaque_2d", 320, 240, 1, random_lines(320, 240, 400, 11, kinds=(0,))),
           ("lines_3d_strict_and_overlay", 320, 240, 2, random_lines(320, 240, 400, 12, kinds=(2, 3))),
           ("lines_all_kinds", 320, 240, 3, random_lines(320, 240, 600, 13)),
           ("lines_all_kinds_dense", 64, 48, 4, random_lines(64, 48, 500, 14, spread=1.1)),
           ("lines_odd_size", 37, 23, 5, random_lines(37, 23, 200, 15, spread=3.0)),
           ("lines_640x480", 640, 480, 6, random_lines(640, 480, 1500, 16)),
           ("lines_star_alpha", 160, 120, 7, star_lines(160, 120, 40, 1)),
           ("lines_star_average", 160, 120, 8, star_lines(160, 120, 9, 0, mode=1)),
           ("lines_star_3d_alpha", 160, 120, 9, star_lines(160, 120, 12, 4, alpha=77))]
    ln = random_lines(96, 64, 300, 17)
    ln[::7]["kind"] = 0
    out.append(("lines_far_endpoints", 96, 64, 10, ln))
    ln = out[-1][4]
    ln["x1"][::5] = 200000; ln["y0"][::9] = -150000                      # long off-screen runs (still below the coordinate cap)
    return out


# ---- star field (render_stars, render.rs:149-199): the host part that builds the b32_star list ----------------
def star_list(camera, w, h, time, seed=42, count=300, horizon=0.5, twinkle_speed=1.5, color=(255, 250, 220)):
    """What a shim computes per star: the LCG draws (theta, phi, and the twinkle phase of VISIBLE stars only), the
    libm direction, and the brightness-scaled colour.  f32 arithmetic through numpy scalars; sin/cos are libm's
    cosf/sinf (what Rust's f32::cos/sin call on Linux)."""
    import ctypes as C
    from bonnie32_b200 import abi
    libm = C.CDLL("libm.so.6")
    libm.cosf.restype = libm.sinf.restype = C.c_float
    libm.cosf.argtypes = libm.sinf.argtypes = [C.c_float]
    F = np.float32
    cosf = lambda v: F(libm.cosf(float(v)))
    sinf = lambda v: F(libm.sinf(float(v)))
    PI = F(np.pi)
    state = [seed & 0xFFFFFFFFFFFFFFFF]
    def next_rand():
        state[0] = (state[0] * 1103515245 + 12345) & 0xFFFFFFFFFFFFFFFF
        return F(state[0] >> 16) / F(65536.0)
    bx, by, bz = (np.asarray(b, dtype=F) for b in (camera.basis_x, camera.basis_y, camera.basis_z))
    out = np.zeros(count, dtype=abi.STAR_DTYPE)
    for i in range(count):
        theta = next_rand() * F(2.0) * PI
        phi = next_rand() * (F(horizon) * PI)
        y = cosf(phi); ring = sinf(phi)
        d = np.array([ring * cosf(theta), y, ring * sinf(theta)], dtype=F)
        v = d * F(10000.0)
        cz = v[0] * bz[0] + v[1] * bz[1] + v[2] * bz[2]
        brightness = F(1.0)
        if cz > F(0.1) and twinkle_speed > 0.0:
            phase = next_rand() * F(2.0) * PI
            brightness = F(0.5) + F(0.5) * sinf(F(time) * F(twinkle_speed) + phase)
        out["dir"][i] = d
        out["rgb"][i] = [min(max(int(F(c) * brightness), 0), 255) for c in color]
    return out


def random_prims(w, h, n, seed):
    """A list mixing the five line kinds with circles, alpha circles, filled rectangles and thick lines."""
    from bonnie32_b200 import abi
    rng = np.random.default_rng(seed)
    ln = random_lines(w, h, n, seed + 1)
    kind = rng.integers(0, 9, n).astype(np.uint8)
    ln["kind"] = kind
    circ = (kind == abi.LINE_CIRCLE) | (kind == abi.LINE_CIRCLE_ALPHA)
    ln["x1"][circ] = rng.choice(np.array([-3, 0, 1, 2, 5, 9, 17, 40, 300], dtype=np.int32), int(circ.sum()))      # radius
    ln["y1"][circ] = 0
    thick = kind == abi.LINE_THICK
    ln["z0"][thick] = rng.choice(np.array([-2, 0, 1, 2, 3, 4, 7, 16], dtype=np.float32), int(thick.sum()))          # thickness
    short = thick & (rng.random(n) < 0.2)                                       # zero-length thick lines draw nothing
    ln["x1"][short] = ln["x0"][short]; ln["y1"][short] = ln["y0"][short]
    big = (kind == abi.LINE_FILLED_RECT) & (rng.random(n) < 0.5)                # keep half of the rectangles small
    for f in ("x1", "y1"):
        ln[f][(kind == abi.LINE_FILLED_RECT) & ~big] = ln[f.replace("1", "0")][(kind == abi.LINE_FILLED_RECT) & ~big] + rng.integers(-12, 13, int(((kind == abi.LINE_FILLED_RECT) & ~big).sum()))
    return ln


def prim_cases():
    """(name, width, height, background seed, list)."""
    return [("prims_320x240", 320, 240, 21, random_prims(320, 240, 300, 31)),
            ("prims_dense_64x48", 64, 48, 22, random_prims(64, 48, 250, 32)),
            ("prims_odd_size", 37, 23, 23, random_prims(37, 23, 150, 33)),
            ("prims_640x480", 640, 480, 24, random_prims(640, 480, 400, 34))]
