"""Second, independently written restatement of the reference rasterizer (numpy, vectorised per
triangle) — TEST INFRASTRUCTURE ONLY.

Purpose: the Rust reference cannot be executed here and render.rs has no tests, so the C++ oracle
(oracle/b32_oracle.cpp, scalar loops) is cross-checked against this model, which was written
straight from the Rust source in a different style (whole-array float32 operations, integer
arithmetic in int64/uint64 with explicit wrapping).  Both must agree bit for bit; the golden
fixtures in tests/golden/ were produced by THIS model (tests/golden/make_golden.py).

numpy float32 arrays round after every operator and never fuse, which is exactly the reference's
f32 semantics (Rust never contracts a*b+c).  `np.add.accumulate` is a sequential running sum, which
restates the reference's incremental edge stepping (render.rs:1526-1532, 1706-1712).

Citations are path:line under /root/reference/src/rasterizer/.
"""
from __future__ import annotations

import numpy as np

F = np.float32
I32 = np.int32
I64 = np.int64
U64 = np.uint64

OPAQUE, AVERAGE, ADD, SUBTRACT, ADD_QUARTER, ERASE = range(6)
SH_NONE, SH_FLAT, SH_GOURAUD = range(3)
L_DIR, L_POINT, L_SPOT = range(3)


class ReferencePanic(Exception):
    """Raised where the Rust reference would panic."""


# ------------------------------------------------------------------------------------------
# Rust `as` casts
# ------------------------------------------------------------------------------------------
def as_i32(f):
    f = np.asarray(f, dtype=F)
    with np.errstate(invalid="ignore"):
        g = np.where(np.isnan(f), F(0), f).astype(np.float64)
        g = np.clip(g, -2147483648.0, 2147483647.0)
        return np.trunc(g).astype(I64).astype(I32)


def as_usize(f):
    f = np.asarray(f, dtype=F)
    g = np.where(np.isnan(f), F(0), f).astype(np.float64)
    g = np.clip(g, 0.0, 2.0 ** 62)          # saturation above 2^62 is irrelevant: always .min()'d
    return np.trunc(g).astype(I64)


def as_u8(f):
    f = np.asarray(f, dtype=F)
    g = np.where(np.isnan(f), F(0), f).astype(np.float64)
    return np.trunc(np.clip(g, 0.0, 255.0)).astype(I64)


def fmin(a, b):
    """f32::min (NaN operand -> the other one)."""
    a = np.asarray(a, dtype=F); b = np.asarray(b, dtype=F)
    return np.fmin(a, b)


def fmax(a, b):
    a = np.asarray(a, dtype=F); b = np.asarray(b, dtype=F)
    return np.fmax(a, b)


def wrap32(x):
    """i64 -> i32 two's-complement truncation."""
    return np.asarray(x, dtype=I64).astype(I32)


# ------------------------------------------------------------------------------------------
# fixed.rs
# ------------------------------------------------------------------------------------------
def unr_table():                                              # fixed.rs:20-31
    t = np.zeros(257, dtype=I64)
    for i in range(257):
        val = ((0x40000 // (i + 0x100)) + 1) // 2 - 0x101
        t[i] = val if val > 0 else 0
    return t


UNR_TABLE = unr_table()


def fx_from_f32(f):                                           # fixed.rs:125-127
    return as_i32(np.asarray(f, dtype=F) * F(4096.0))


def fx_mul(a, b):                                             # fixed.rs:161-165
    return wrap32((np.asarray(a, dtype=I64) * np.asarray(b, dtype=I64)) >> 12)


def fx_add(a, b):                                             # fixed.rs:236-238
    return wrap32(np.asarray(a, dtype=I64) + np.asarray(b, dtype=I64))


def fx_sub(a, b):
    return wrap32(np.asarray(a, dtype=I64) - np.asarray(b, dtype=I64))


def div_unr(num_i32, den_i32):                                # fixed.rs:178-230
    num_i32 = np.atleast_1d(np.asarray(num_i32, dtype=I64))
    den_i32 = np.atleast_1d(np.asarray(den_i32, dtype=I64))
    num_i32, den_i32 = np.broadcast_arrays(num_i32, den_i32)
    neg = (num_i32 < 0) != (den_i32 < 0)
    num = np.abs(num_i32).astype(U64)
    den = np.abs(den_i32).astype(U64)                          # 1 .. 2^31
    zero = den == 0
    den_safe = np.where(zero, U64(1), den)
    # leading zeros of a u32
    nbits = np.floor(np.log2(den_safe.astype(np.float64))).astype(I64) + 1
    # guard against log2 rounding at exact powers of two
    nbits = np.where((U64(1) << nbits.astype(U64)) <= den_safe, nbits + 1, nbits)
    nbits = np.where((U64(1) << (nbits - 1).astype(U64)) > den_safe, nbits - 1, nbits)
    z = (32 - nbits).astype(U64)
    with np.errstate(over="ignore"):
        d16 = (den_safe << z) >> U64(16)
        idx = np.minimum((d16 - U64(0x7FC0)) >> U64(7), U64(256)).astype(I64)
        u = UNR_TABLE[idx].astype(U64) + U64(0x101)
        nr1 = (U64(0x2000080) - d16 * u) >> U64(8)
        nr2 = (U64(0x80) + nr1 * u) >> U64(8)
        raw = num * nr2
        shift = U64(36) - z                                    # 5..36
        mag = (raw + (U64(1) << (shift - U64(1)))) >> shift
    mag = np.minimum(mag, U64(0x7FFFFFFF)).astype(I64)
    out = np.where(neg, -mag, mag)
    return np.where(zero, 0, out).astype(I32)


def project_fixed(pos, cam, width, height):
    """fixed.rs:362-441 for an array of world positions pos[n,3] (f32). Returns sx, sy (int32)."""
    P = fx_from_f32(pos)                                        # from_vec3 of world_pos
    Cp = fx_from_f32(cam.position)
    rel = fx_sub(P, Cp[None, :])
    bx, by, bz = fx_from_f32(cam.basis_x), fx_from_f32(cam.basis_y), fx_from_f32(cam.basis_z)

    def dotf(b):                                               # fixed.rs:311-313
        return fx_add(fx_add(fx_mul(rel[:, 0], b[0]), fx_mul(rel[:, 1], b[1])), fx_mul(rel[:, 2], b[2]))

    cx, cy, cz = dotf(bx), dotf(by), dotf(bz)
    distance = fx_from_f32(F(5.0))
    scale = fx_from_f32(F(4.0))
    viewport_scale = fx_from_f32((F(min(width, height)) / F(2.0)) * F(0.75))
    half_w = I32((width // 2) << 12)
    half_h = I32((height // 2) << 12)
    denom = fx_add(cz, distance)
    absd = np.abs(denom.astype(I64))
    absd = np.where(denom == np.iinfo(I32).min, I64(np.iinfo(I32).min), absd)   # release-mode i32::abs wraps
    near = absd < 256
    px = div_unr(fx_mul(cx, scale), np.where(near, 1, denom))
    py = div_unr(fx_mul(cy, scale), np.where(near, 1, denom))
    sx = fx_add(fx_mul(px, viewport_scale), half_w).astype(I64) >> 12
    sy = fx_add(fx_mul(py, viewport_scale), half_h).astype(I64) >> 12
    sx = np.where(near, I64(half_w) >> 12, sx)
    sy = np.where(near, I64(half_h) >> 12, sy)
    return sx.astype(I32), sy.astype(I32)


# ------------------------------------------------------------------------------------------
# math.rs
# ------------------------------------------------------------------------------------------
def dot3(a, b):                                               # math.rs:23-25, a[n,3] . b[3]
    a = np.asarray(a, dtype=F); b = np.asarray(b, dtype=F)
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def normalize3(a):                                            # math.rs:39-49
    a = np.asarray(a, dtype=F)
    l = np.sqrt(dot3(a, a))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = a / l[..., None]
    return np.where((l == 0)[..., None], F(0), out).astype(F)


def perspective_transform(v, cam):                            # math.rs:103-109
    return np.stack([dot3(v, cam.basis_x), dot3(v, cam.basis_y), dot3(v, cam.basis_z)], axis=-1).astype(F)


# ------------------------------------------------------------------------------------------
# lighting, render.rs:1013-1071 (scalar; called 1-3 times per triangle)
# ------------------------------------------------------------------------------------------
def _acosf_rpoly(z):
    p = F(z * F(F(0.16666586697101593) + F(z * F(F(-0.04274342209100723) + F(z * F(-0.008656363002955914))))))
    q = F(F(z * F(-0.7066296339035034)) + F(1.0))
    return F(p / q)


def ref_acosf(x):
    """f32::acos as the reference's wasm build computes it: compiler_builtins' libm `acosf` (port of musl e_acosf.c),
    docs/bonnie-32.wasm func 2057; called at render.rs:1047.  Scalar, one rounding per operator."""
    x = F(x)
    pio2_hi, pio2_lo = F(1.570796251296997), F(7.549789415861596e-08)
    hx = int(np.asarray(x, dtype=F).view(np.uint32))
    ix = hx & 0x7FFFFFFF
    with np.errstate(invalid="ignore", divide="ignore"):
        if ix >= 0x3F800000:
            if ix == 0x3F800000:
                return F(3.141592502593994) if hx >> 31 else F(0.0)
            return F(F(0.0) / F(x - x))
        if ix < 0x3F000000:
            if ix <= 0x32800000:
                return pio2_hi
            return F(pio2_hi - F(x - F(pio2_lo - F(x * _acosf_rpoly(F(x * x))))))
        if hx >> 31:
            z = F(F(F(1.0) + x) * F(0.5))
            s = F(np.sqrt(z))
            w = F(F(_acosf_rpoly(z) * s) - pio2_lo)
            t = F(pio2_hi - F(s + w))
            return F(t + t)
        z = F(F(F(1.0) - x) * F(0.5))
        s = F(np.sqrt(z))
        df = (np.asarray(s, dtype=F).view(np.uint32) & np.uint32(0xFFFFF000)).view(F)[()]
        c = F(F(z - F(df * df)) / F(s + df))
        w = F(F(_acosf_rpoly(z) * s) + c)
        t = F(df + w)
        return F(t + t)


def shade_multi_light_color(normal, world_pos, lights, ambient):
    normal = np.asarray(normal, dtype=F); world_pos = np.asarray(world_pos, dtype=F)
    tot = [F(ambient), F(ambient), F(ambient)]
    for L in lights:
        if not L.enabled:
            continue
        if L.type == L_DIR:
            neg_dir = np.asarray(L.direction, dtype=F) * F(-1.0)
            n_dot_l = fmax(dot3(normal, neg_dir), F(0.0))
            contribution = F(n_dot_l * F(L.intensity))
        elif L.type == L_POINT:
            to_light = np.asarray(L.position, dtype=F) - world_pos
            dist = np.sqrt(dot3(to_light, to_light))
            if dist > F(L.radius) or dist < F(0.001):
                contribution = F(0.0)
            else:
                att = F(1.0) - (dist / F(L.radius))
                n_dot_l = fmax(dot3(normal, normalize3(to_light)), F(0.0))
                contribution = F(F(F(n_dot_l * F(L.intensity)) * att) * att)
        else:                                                    # Spot, render.rs:1038-1059
            to_light = np.asarray(L.position, dtype=F) - world_pos
            dist = np.sqrt(dot3(to_light, to_light))
            if dist > F(L.radius) or dist < F(0.001):
                contribution = F(0.0)
            else:
                to_surface = normalize3(to_light)
                spot_angle = ref_acosf(dot3(to_surface * F(-1.0), np.asarray(L.direction, dtype=F)))
                if spot_angle > F(L.angle):
                    contribution = F(0.0)
                else:
                    att = F(1.0) - (dist / F(L.radius))
                    edge = F(1.0) - (spot_angle / F(L.angle))
                    n_dot_l = fmax(dot3(normal, to_surface), F(0.0))
                    contribution = F(F(F(F(n_dot_l * F(L.intensity)) * att) * att) * edge)
        lr, lg, lb = (F(c) / F(255.0) for c in L.color)
        tot[0] = F(tot[0] + F(contribution * lr))
        tot[1] = F(tot[1] + F(contribution * lg))
        tot[2] = F(tot[2] + F(contribution * lb))
    return [F(fmin(t, F(1.0))) for t in tot]


# ------------------------------------------------------------------------------------------
# textures
# ------------------------------------------------------------------------------------------
def texels_u16(tex):
    """Texture15.pixels as a u16[h,w] array; indexed inputs go through Clut::lookup
    (types.rs:390-397) per texel, as IndexedAtlas::to_texture15 does (mesh_editor.rs:669-682)."""
    n = tex.width * tex.height
    if tex.format == 0:
        px = np.asarray(tex.pixels, dtype=np.uint16).reshape(-1)[:n]
    else:
        raw = np.asarray(tex.pixels, dtype=np.uint8).reshape(-1)
        if tex.format == 1:
            idx = raw[:n].astype(I64)
        else:
            idx = np.stack([raw & 0xF, raw >> 4], axis=-1).reshape(-1)[:n].astype(I64)
        clut = np.asarray(tex.clut, dtype=np.uint16)
        px = np.where(idx < clut.size, clut[np.minimum(idx, clut.size - 1)], np.uint16(0))
    return px.reshape(tex.height, tex.width).astype(np.uint16)


def rem_euclid_1(u):                                          # core f32::rem_euclid(1.0)
    with np.errstate(invalid="ignore"):
        r = np.fmod(u, F(1.0)).astype(F)
    return np.where(r < 0, r + F(1.0), r).astype(F)


def sample(px, u, v):                                         # types.rs:671-681
    h, w = px.shape
    if w == 0 or h == 0:
        return np.zeros(u.shape, dtype=np.uint16)
    tx = np.minimum(as_usize(rem_euclid_1(u) * F(w)), w - 1)
    ty = np.minimum(as_usize(rem_euclid_1(v) * F(h)), h - 1)
    return px[ty, tx]


DITHER = np.array([[-4, 0, -3, 1], [2, -2, 3, -1], [-3, 1, -4, 0], [3, -1, 2, -2]], dtype=I64)   # render.rs:1150-1155


def expand58(v):                                              # render.rs:1161-1163
    return ((v << 3) | (v >> 2)) & 0xFF


def blend555(f8, b8, mode):                                   # render.rs:1093-1145, arrays [n,3]
    f5 = f8 >> 3
    b5 = b8 >> 3
    if mode == OPAQUE:
        r = f5
    elif mode == AVERAGE:
        r = np.minimum((b5 + f5) // 2, 31)
    elif mode == ADD:
        r = np.minimum(b5 + f5, 31)
    elif mode == SUBTRACT:
        r = np.maximum(b5 - f5, 0)
    elif mode == ADD_QUARTER:
        r = np.minimum(b5 + f5 // 4, 31)
    else:
        r = b5
    return r << 3


# ------------------------------------------------------------------------------------------
# render_mesh_15, render.rs:2302-2572 (wireframe phase not modelled: off in every golden scene)
# ------------------------------------------------------------------------------------------
def render_mesh_15(fb_rgba, fb_z, vertices, faces, textures, camera, settings, fog=None):
    """fb_rgba u8[h,w,4], fb_z f32[h,w] are updated in place. Returns list of face_idx in draw order."""
    tex_px = [texels_u16(t) for t in textures]
    wire = {"back": [], "front": []}
    surfaces = _build_surfaces(fb_z.shape, vertices, faces, textures, camera, settings, fog, rgb888=False, wire=wire)

    # ---- SORT (render.rs:2518-2545) ----
    opaque = [s for s in surfaces if not s["has_tr"]]
    transp = [s for s in surfaces if s["has_tr"]]
    transp = _back_to_front(transp)
    if not settings.use_zbuffer:
        opaque = _back_to_front(opaque)

    # ---- DRAW (render.rs:2547-2572) ----
    if not settings.wireframe_overlay:
        for s in opaque:
            _fill(fb_rgba, fb_z, s, textures, tex_px, settings, skip_z_write=False)
        for s in transp:
            _fill(fb_rgba, fb_z, s, textures, tex_px, settings, skip_z_write=True)
    _wireframe_phase(fb_rgba, fb_z, wire, settings)
    return [s["face_idx"] for s in opaque + transp]


def _back_to_front(lst):
    """stable `sort_by(b.partial_cmp(a).unwrap())` on the centre depth (render.rs:2527-2532 / :2157-2162)."""
    if len(lst) < 2:
        return lst
    keys = np.array([(s["v"][0][2] + s["v"][1][2] + s["v"][2][2]) / F(3.0) for s in lst], dtype=F)
    if np.isnan(keys).any():
        raise ReferencePanic("partial_cmp().unwrap() on NaN")
    order = np.argsort(-keys, kind="stable")                   # descending, ties keep order
    return [lst[i] for i in order]


def _build_surfaces(shape, vertices, faces, textures, camera, settings, fog, rgb888, wire=None):
    """TRANSFORM + CULL phases shared by render_mesh_15 (render.rs:2313-2516) and render_mesh
    (render.rs:1981-2149; no fog, has_transparency = texture blend or editor alpha)."""
    H, W = shape
    pos = np.asarray(vertices["pos"], dtype=F)
    nv = len(pos)
    cpos = np.asarray(camera.position, dtype=F)

    # ---- TRANSFORM (render.rs:2321-2360) ----
    rel = pos - cpos[None, :]
    cam_pos = perspective_transform(rel, camera)
    with np.errstate(all="ignore"):
        if settings.ortho_projection is not None:                       # math.rs:140-148
            zoom, ocx, ocy = (F(x) for x in settings.ortho_projection)
            sx = (cam_pos[:, 0] - ocx) * zoom + (F(W) / F(2.0))
            sy = -(cam_pos[:, 1] - ocy) * zoom + (F(H) / F(2.0))
            sz = cam_pos[:, 2]
        elif settings.use_fixed_point:
            ix, iy = project_fixed(pos, camera, W, H)
            sx, sy = ix.astype(F), iy.astype(F)
            sz = cam_pos[:, 2] + F(5.0)
        else:                                                           # math.rs:117-136
            us = F(5.0) - F(1.0)
            vs = (F(min(W, H)) / F(2.0)) * F(0.75)
            denom = cam_pos[:, 2] + F(5.0)
            tiny = np.abs(denom) < F(0.001)
            sx = np.where(tiny, F(W) / F(2.0), (cam_pos[:, 0] * us) / denom * vs + (F(W) / F(2.0)))
            sy = np.where(tiny, F(H) / F(2.0), (cam_pos[:, 1] * us) / denom * vs + (F(H) / F(2.0)))
            sz = np.where(tiny, cam_pos[:, 2], denom)
    proj = np.stack([sx, sy, sz], axis=-1).astype(F)

    # ---- CULL / BUILD (render.rs:2373-2513), one Python dict per surviving surface ----
    fv = np.asarray(faces["v"], dtype=I64)
    if len(fv) and fv.max() >= nv:
        raise ReferencePanic("vertex index out of bounds")
    flags = np.asarray(faces["flags"], dtype=I64)
    rgba = np.asarray(vertices["rgba"], dtype=I64)
    uvs = np.asarray(vertices["uv"], dtype=F)
    nrm = np.asarray(vertices["normal"], dtype=F)

    surfaces = []
    for fi in range(len(fv)):
        i0, i1, i2 = fv[fi]
        tex_id = int(flags[fi] & 0xFFFF)
        face_blend = int((flags[fi] >> 16) & 7)
        black_tr = bool((flags[fi] >> 19) & 1)
        editor_alpha = int((flags[fi] >> 24) & 0xFF)
        tex = tex_id if (tex_id != 0xFFFF and tex_id < len(textures)) else None
        cz = cam_pos[[i0, i1, i2], 2]
        if settings.ortho_projection is None and (cz <= F(0.1)).any():                 # :2380-2385
            continue
        v1, v2, v3 = proj[i0], proj[i1], proj[i2]
        signed_area = (v2[0] - v1[0]) * (v3[1] - v1[1]) - (v3[0] - v1[0]) * (v2[1] - v1[1])   # :2393
        backface = bool(signed_area <= 0)
        tex_blend = textures[tex].blend_mode if tex is not None else None
        if tex_blend is not None and tex_blend != OPAQUE:                               # :2403-2415 / :2071-2075
            has_tr = True
        elif face_blend != OPAQUE and not rgb888:
            has_tr = True
        else:
            has_tr = editor_alpha < 255
        cols = [tuple(rgba[i]) for i in (i0, i1, i2)]
        if fog is not None:                                                             # :2419-2436
            start, falloff, cull_d, fcol = F(fog[0]), F(fog[1]), F(fog[2]), tuple(fog[3])
            if len(fcol) == 3:
                fcol = fcol + (OPAQUE,)
            if (cz > cull_d).all():
                continue
            newc = []
            for c, z in zip(cols, cz):
                if z <= start:                                                          # :2266-2275
                    f = F(0.0)
                elif falloff <= 0:
                    f = F(1.0)
                else:
                    f = F(fmin((z - start) / falloff, F(1.0)))
                if f <= 0:                                                              # :2279-2293
                    newc.append(c)
                elif f >= 1:
                    newc.append(tuple(int(x) for x in fcol))
                else:
                    inv = F(1.0) - f
                    newc.append(tuple(int(as_u8(F(c[k]) * inv + F(fcol[k]) * f)) for k in range(3)) + (OPAQUE,))
            cols = newc
        order = (0, 1, 2)
        sign = F(1.0)
        if wire is not None:                                                            # :2446-2450, :2509-2511
            if backface and not settings.xray_mode:
                wire["back"].append((v1, v2, v3))
            elif not backface and settings.wireframe_overlay:
                wire["front"].append((v1, v2, v3))
        if backface:
            if not (not settings.backface_cull or settings.xray_mode):                  # :2453
                continue
            order = (0, 2, 1)                                                           # swap v2/v3
            sign = F(-1.0)
        idx = (i0, i1, i2)
        surfaces.append(dict(
            v=[proj[idx[k]] for k in order],
            w=[pos[idx[k]] for k in order],
            wn=[nrm[idx[k]] * sign if backface else nrm[idx[k]] for k in order],
            uv=[uvs[idx[k]] for k in order],
            vc=[cols[k] for k in order],
            face_idx=fi, tex=tex, black_tr=black_tr, has_tr=has_tr, blend=face_blend, editor_alpha=editor_alpha))

    return surfaces


def _fill(fb_rgba, fb_z, s, textures, tex_px, settings, skip_z_write):
    """rasterize_triangle_15, render.rs:1440-1714, vectorised over the bounding box."""
    H, W = fb_z.shape
    v1, v2, v3 = s["v"]
    tex = s["tex"]
    blend_mode = textures[tex].blend_mode if tex is not None else s["blend"]            # :1450-1452

    min_x = int(as_usize(fmax(fmin(fmin(v1[0], v2[0]), v3[0]), F(0.0))))                # :1455-1458
    max_x = int(as_usize(fmin(fmax(fmax(v1[0], v2[0]), v3[0]) + F(1.0), F(W))))
    min_y = int(as_usize(fmax(fmin(fmin(v1[1], v2[1]), v3[1]), F(0.0))))
    max_y = int(as_usize(fmin(fmax(fmax(v1[1], v2[1]), v3[1]) + F(1.0), F(H))))
    if min_x >= max_x or min_y >= max_y:
        return
    third = F(1.0) / F(3.0)
    flat = None
    gour = None
    if settings.shading == SH_FLAT:                                                     # :1466-1472
        center = ((s["w"][0] + s["w"][1]) + s["w"][2]) * third
        wn = normalize3((((s["wn"][0] + s["wn"][1]) + s["wn"][2]) * third)[None, :])[0]
        flat = shade_multi_light_color(wn, center, settings.lights, settings.ambient)
    elif settings.shading == SH_GOURAUD:                                                # :1475-1483
        gour = [shade_multi_light_color(s["wn"][k], s["w"][k], settings.lights, settings.ambient) for k in range(3)]
    vc = s["vc"]
    needs_dither = settings.dithering and (settings.shading == SH_GOURAUD or tex is not None
                                           or vc[0] != vc[1] or vc[1] != vc[2])          # :1487-1492

    with np.errstate(all="ignore"):
        area = (v2[1] - v3[1]) * (v1[0] - v3[0]) + (v3[0] - v2[0]) * (v1[1] - v3[1])     # :1500
        if np.abs(area) < F(0.00001):
            return
        inv_area = F(1.0) / area
        a0 = v2[1] - v3[1]; b0 = v3[0] - v2[0]; a1 = v3[1] - v1[1]; b1 = v1[0] - v3[0]   # :1507-1510
        start_x = F(min_x); start_y = F(min_y)
        w0s = a0 * (start_x - v3[0]) + b0 * (start_y - v3[1])                            # :1517-1518
        w1s = a1 * (start_x - v3[0]) + b1 * (start_y - v3[1])
        ny, nx = max_y - min_y, max_x - min_x

        def grid(ws, a, b):
            col = np.add.accumulate(np.concatenate([[ws], np.full(ny - 1, b, dtype=F)]).astype(F), dtype=F)
            g = np.empty((ny, nx), dtype=F)
            g[:, 0] = col
            if nx > 1:
                g[:, 1:] = a
            return np.add.accumulate(g, axis=1, dtype=F)

        w0 = grid(w0s, a0, b0)
        w1 = grid(w1s, a1, b1)
        bc_x = w0 * inv_area
        bc_y = w1 * inv_area
        bc_z = F(1.0) - bc_x - bc_y
        ERR = F(-0.0001)
        inside = (bc_x >= ERR) & (bc_y >= ERR) & (bc_z >= ERR)                            # :1541-1542
        inv_z1 = F(1.0) / v1[2]; inv_z2 = F(1.0) / v2[2]; inv_z3 = F(1.0) / v3[2]
        inv_z = bc_x * inv_z1 + bc_y * inv_z2 + bc_z * inv_z3
        z = F(1.0) / inv_z
        zb = fb_z[min_y:max_y, min_x:max_x]
        px = fb_rgba[min_y:max_y, min_x:max_x]
        live = inside.copy()
        use_z = settings.use_zbuffer and not settings.xray_mode
        if use_z:
            live &= ~(z >= zb)                                                            # :1553-1560
        uv1, uv2, uv3 = s["uv"]
        if settings.affine_textures:                                                      # :1563-1579
            u = bc_x * uv1[0] + bc_y * uv2[0] + bc_z * uv3[0]
            v = bc_x * uv1[1] + bc_y * uv2[1] + bc_z * uv3[1]
        else:
            uo = bc_x * uv1[0] * inv_z1 + bc_y * uv2[0] * inv_z2 + bc_z * uv3[0] * inv_z3
            vo = bc_x * uv1[1] * inv_z1 + bc_y * uv2[1] * inv_z2 + bc_z * uv3[1] * inv_z3
            u = uo / inv_z
            v = vo / inv_z
        if tex is not None:
            color = sample(tex_px[tex], u, F(1.0) - v).astype(I64)                        # :1582-1586
        else:
            color = np.full((ny, nx), 0x7FFF, dtype=I64)
        r5 = (color >> 10) & 31; g5 = (color >> 5) & 31; b5 = color & 31
        is_black = (r5 == 0) & (g5 == 0) & (b5 == 0)
        transparent = color == 0
        if s["black_tr"]:                                                                 # :1591-1607
            live &= ~is_black
        else:
            color = np.where(transparent, 0x8000, color)
        semi_in = (color & 0x8000) != 0

        tex8 = np.stack([expand58(r5), expand58(g5), expand58(b5)], axis=-1)              # :1613-1615
        vcol = np.stack([as_u8(bc_x * F(vc[0][k]) + bc_y * F(vc[1][k]) + bc_z * F(vc[2][k])) for k in range(3)], axis=-1)
        mod8 = np.minimum((tex8 * vcol) // 128, 255)                                       # :1624-1626
        if settings.shading == SH_NONE:
            shade = [np.full((ny, nx), F(1.0), dtype=F)] * 3
        elif settings.shading == SH_FLAT:
            shade = [np.full((ny, nx), flat[k], dtype=F) for k in range(3)]
        else:
            shade = [bc_x * gour[0][k] + bc_y * gour[1][k] + bc_z * gour[2][k] for k in range(3)]
        shaded = []
        for k in range(3):                                                                # :1643-1645
            sc = shade[k].astype(F)
            sc = np.where(sc < F(0.0), F(0.0), sc)          # f32::clamp keeps NaN
            sc = np.where(sc > F(2.0), F(2.0), sc)
            shaded.append(as_u8(fmin(mod8[..., k].astype(F) * sc, F(255.0))))
        shaded = np.stack(shaded, axis=-1)
        if needs_dither:                                                                  # :1173-1182
            yy, xx = np.meshgrid(np.arange(min_y, max_y), np.arange(min_x, max_x), indexing="ij")
            off = DITHER[yy & 3, xx & 3]
            q = np.clip((shaded + off[..., None]) >> 3, 0, 31)
        else:
            q = shaded >> 3
        all_black = (q == 0).all(axis=-1)                                                 # :1659-1661
        semi = semi_in | all_black
        out8 = expand58(q)                                                                # Color15::r8/g8/b8

        ea = s["editor_alpha"]
        if ea == 0:                                                                       # :1664-1669
            return
        back = px[..., :3].astype(I64)
        if settings.xray_mode:                                                            # :507-526
            new = (out8 + back) // 2
            wmask = live
        else:
            do_blend = semi & (blend_mode != OPAQUE)
            ps1 = np.where(do_blend[..., None], blend555(out8, back, blend_mode), out8)
            if ea < 255:                                                                  # :567-628
                # depth variant rejects on `z >= zbuffer` (:604) — the same predicate as the early
                # test, so nothing further is masked here (differs from `<` only for NaN z)
                wmask = live
                new = (ps1 * ea + back * (255 - ea)) // 255
            else:
                wmask = live & (z < zb) if settings.use_zbuffer else live               # :1684
                new = ps1
            if settings.use_zbuffer and not skip_z_write:
                zb[wmask] = z[wmask]
        px[wmask, 0] = new[wmask, 0]
        px[wmask, 1] = new[wmask, 1]
        px[wmask, 2] = new[wmask, 2]
        px[wmask, 3] = 255


# ------------------------------------------------------------------------------------------
# render_mesh, render.rs:1971-2259 (RGB888 path) — textures: objects with width, height, blend_mode and
# pixels = u8[h*w*4] (r, g, b, blend per texel: struct Color, types.rs:721-726)
# ------------------------------------------------------------------------------------------
def render_mesh(fb_rgba, fb_z, vertices, faces, textures, camera, settings):
    tex_px = [np.asarray(t.pixels, dtype=np.uint8).reshape(t.height, t.width, 4).astype(I64) if t.width * t.height
              else np.zeros((t.height, t.width, 4), dtype=I64) for t in textures]
    wire = {"back": [], "front": []}
    surfaces = _build_surfaces(fb_z.shape, vertices, faces, textures, camera, settings, None, rgb888=True, wire=wire)
    if not settings.use_zbuffer:                                   # :2155-2162, one list
        surfaces = _back_to_front(surfaces)
    if not settings.wireframe_overlay:                             # :2172-2181
        for s in surfaces:
            _fill888(fb_rgba, fb_z, s, tex_px, settings)
    _wireframe_phase(fb_rgba, fb_z, wire, settings)                # :2195-2256, the same phase
    return [s["face_idx"] for s in surfaces]


def blend888(f, b, mode):                                      # Color::blend_with, types.rs:886-930; int arrays [n,3]
    if mode == OPAQUE:
        return f
    if mode == AVERAGE:
        return (b + f) // 2
    if mode == ADD:
        return np.minimum(b + f, 255)
    if mode == SUBTRACT:
        return np.maximum(b - f, 0)
    if mode == ADD_QUARTER:
        return np.minimum(b + f // 4, 255)
    raise AssertionError("Erase never reaches a writer (skipped at render.rs:1350)")


def _fill888(fb_rgba, fb_z, s, tex_px, settings):
    """rasterize_triangle, render.rs:1202-1433, vectorised over the bounding box.  A surface never covers
    a pixel twice, so per-pixel reads of the framebuffer see the state before this surface."""
    H, W = fb_z.shape
    v1, v2, v3 = s["v"]
    tex = s["tex"]
    min_x = int(as_usize(fmax(fmin(fmin(v1[0], v2[0]), v3[0]), F(0.0))))                # :1209-1212
    max_x = int(as_usize(fmin(fmax(fmax(v1[0], v2[0]), v3[0]) + F(1.0), F(W))))
    min_y = int(as_usize(fmax(fmin(fmin(v1[1], v2[1]), v3[1]), F(0.0))))
    max_y = int(as_usize(fmin(fmax(fmax(v1[1], v2[1]), v3[1]) + F(1.0), F(H))))
    if min_x >= max_x or min_y >= max_y:
        return
    third = F(1.0) / F(3.0)
    flat = gour = None
    if settings.shading == SH_FLAT:                                                     # :1220-1226
        center = ((s["w"][0] + s["w"][1]) + s["w"][2]) * third
        wn = normalize3((((s["wn"][0] + s["wn"][1]) + s["wn"][2]) * third)[None, :])[0]
        flat = shade_multi_light_color(wn, center, settings.lights, settings.ambient)
    elif settings.shading == SH_GOURAUD:                                                # :1229-1237
        gour = [shade_multi_light_color(s["wn"][k], s["w"][k], settings.lights, settings.ambient) for k in range(3)]
    vc = s["vc"]
    needs_dither = settings.dithering and (settings.shading == SH_GOURAUD or tex is not None
                                           or vc[0] != vc[1] or vc[1] != vc[2])          # :1241-1246
    with np.errstate(all="ignore"):
        area = (v2[1] - v3[1]) * (v1[0] - v3[0]) + (v3[0] - v2[0]) * (v1[1] - v3[1])     # :1257
        if np.abs(area) < F(0.00001):
            return
        inv_area = F(1.0) / area
        a0 = v2[1] - v3[1]; b0 = v3[0] - v2[0]; a1 = v3[1] - v1[1]; b1 = v1[0] - v3[0]
        w0s = a0 * (F(min_x) - v3[0]) + b0 * (F(min_y) - v3[1])                          # :1278-1279
        w1s = a1 * (F(min_x) - v3[0]) + b1 * (F(min_y) - v3[1])
        ny, nx = max_y - min_y, max_x - min_x

        def grid(ws, a, b):                                                               # incremental stepping :1427-1433
            col = np.add.accumulate(np.concatenate([[ws], np.full(ny - 1, b, dtype=F)]).astype(F), dtype=F)
            g = np.empty((ny, nx), dtype=F)
            g[:, 0] = col
            if nx > 1:
                g[:, 1:] = a
            return np.add.accumulate(g, axis=1, dtype=F)

        bc_x = grid(w0s, a0, b0) * inv_area
        bc_y = grid(w1s, a1, b1) * inv_area
        bc_z = F(1.0) - bc_x - bc_y
        ERR = F(-0.0001)
        live = (bc_x >= ERR) & (bc_y >= ERR) & (bc_z >= ERR)                              # :1303
        inv_z1 = F(1.0) / v1[2]; inv_z2 = F(1.0) / v2[2]; inv_z3 = F(1.0) / v3[2]
        inv_z = bc_x * inv_z1 + bc_y * inv_z2 + bc_z * inv_z3
        z = F(1.0) / inv_z
        zb = fb_z[min_y:max_y, min_x:max_x]
        px = fb_rgba[min_y:max_y, min_x:max_x]
        if settings.use_zbuffer and not settings.xray_mode:
            live &= ~(z >= zb)                                                            # :1313-1320
        uv1, uv2, uv3 = s["uv"]
        if settings.affine_textures:                                                      # :1323-1340
            u = bc_x * uv1[0] + bc_y * uv2[0] + bc_z * uv3[0]
            v = bc_x * uv1[1] + bc_y * uv2[1] + bc_z * uv3[1]
        else:
            uo = bc_x * uv1[0] * inv_z1 + bc_y * uv2[0] * inv_z2 + bc_z * uv3[0] * inv_z3
            vo = bc_x * uv1[1] * inv_z1 + bc_y * uv2[1] * inv_z2 + bc_z * uv3[1] * inv_z3
            u = uo / inv_z
            v = vo / inv_z
        if tex is not None:                                                               # :1343-1347, types.rs:1242-1253
            tp = tex_px[tex]
            th, tw = tp.shape[:2]
            if tw == 0 or th == 0:
                return                                                                    # every sample is TRANSPARENT
            tx = np.minimum(as_usize(rem_euclid_1(u) * F(tw)), tw - 1)
            ty = np.minimum(as_usize(rem_euclid_1(F(1.0) - v) * F(th)), th - 1)
            color = tp[ty, tx]
        else:
            color = np.broadcast_to(np.array([255, 255, 255, OPAQUE], dtype=I64), (ny, nx, 4))
        cblend = color[..., 3]
        live &= cblend != ERASE                                                           # :1350-1354
        vcol = np.stack([as_u8(bc_x * F(vc[0][k]) + bc_y * F(vc[1][k]) + bc_z * F(vc[2][k])) for k in range(3)], axis=-1)
        mod8 = np.minimum((color[..., :3] * vcol) // 128, 255)                             # Color::modulate
        if settings.shading == SH_NONE:
            shade = [np.full((ny, nx), F(1.0), dtype=F)] * 3
        elif settings.shading == SH_FLAT:
            shade = [np.full((ny, nx), flat[k], dtype=F) for k in range(3)]
        else:
            shade = [bc_x * gour[0][k] + bc_y * gour[1][k] + bc_z * gour[2][k] for k in range(3)]
        c8 = np.stack([as_u8(fmin(mod8[..., k].astype(F) * shade[k].astype(F), F(255.0))) for k in range(3)], axis=-1)   # :1074-1081
        if needs_dither:                                                                  # :1186-1197
            yy, xx = np.meshgrid(np.arange(min_y, max_y), np.arange(min_x, max_x), indexing="ij")
            c8 = np.clip((c8 + DITHER[yy & 3, xx & 3][..., None]) >> 3, 0, 31) << 3
        ea = s["editor_alpha"]
        if ea == 0:                                                                       # :1392-1398
            return
        back = px[..., :3].astype(I64)
        ps1 = c8.copy()
        for mode in (AVERAGE, ADD, SUBTRACT, ADD_QUARTER):                                # per-texel blend tag
            m = cblend == mode
            if m.any():
                ps1[m] = blend888(c8[m], back[m], mode)
        if ea < 255:                                                                      # :338-420
            a = F(ea) / F(255.0)
            inv_a = F(1.0) - a
            new = np.stack([as_u8(ps1[..., k].astype(F) * a + back[..., k].astype(F) * inv_a) for k in range(3)], axis=-1)
            wmask = live & ~(z >= zb) if settings.use_zbuffer else live                   # :393
        else:
            new = ps1
            wmask = live & (z < zb) if settings.use_zbuffer else live                     # :425 / :1408
        if settings.use_zbuffer:
            zb[wmask] = z[wmask]
        px[wmask, 0] = new[wmask, 0]
        px[wmask, 1] = new[wmask, 1]
        px[wmask, 2] = new[wmask, 2]
        px[wmask, 3] = 255


# ------------------------------------------------------------------------------------------
# Framebuffer::render_skybox step 1 (sphere pass): render.rs:89-139, rasterize_skybox_triangle :242-299
# ------------------------------------------------------------------------------------------
def render_skybox_mesh(fb_rgba, sky_vertices, faces, camera):
    """sky_vertices: record array with pos f32[3], rgb u8[3]; faces int[nf,3]. No depth; faces in order."""
    H, W = fb_rgba.shape[:2]
    pos = np.asarray(sky_vertices["pos"], dtype=F)
    col = np.asarray(sky_vertices["rgb"], dtype=I64)
    cam_space = perspective_transform(pos - np.asarray(camera.position, dtype=F)[None, :], camera)
    with np.errstate(all="ignore"):
        vs = (F(min(W, H)) / F(2.0)) * F(0.75)                                       # math.rs:117-136
        denom = cam_space[:, 2] + F(5.0)
        tiny = np.abs(denom) < F(0.001)
        sx = np.where(tiny, F(W) / F(2.0), (cam_space[:, 0] * F(4.0)) / denom * vs + (F(W) / F(2.0)))
        sy = np.where(tiny, F(H) / F(2.0), (cam_space[:, 1] * F(4.0)) / denom * vs + (F(H) / F(2.0)))
    behind = cam_space[:, 2] <= F(0.1)
    for i0, i1, i2 in np.asarray(faces, dtype=I64).reshape(-1, 3):
        if behind[i0] or behind[i1] or behind[i2]:
            continue
        p0 = (sx[i0], sy[i0]); p1 = (sx[i1], sy[i1]); p2 = (sx[i2], sy[i2])
        with np.errstate(all="ignore"):
            signed_area = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1])
            if signed_area >= 0:
                continue
            min_x = int(as_usize(fmax(fmin(fmin(p0[0], p1[0]), p2[0]), F(0.0))))
            max_x = int(as_usize(fmin(fmax(fmax(p0[0], p1[0]), p2[0]), F(W) - F(1.0))))
            min_y = int(as_usize(fmax(fmin(fmin(p0[1], p1[1]), p2[1]), F(0.0))))
            max_y = int(as_usize(fmin(fmax(fmax(p0[1], p1[1]), p2[1]), F(H) - F(1.0))))
            if min_x > max_x or min_y > max_y:
                continue
            den = (p1[1] - p2[1]) * (p0[0] - p2[0]) + (p2[0] - p1[0]) * (p0[1] - p2[1])
            if np.abs(den) < F(0.0001):
                continue
            inv = F(1.0) / den
            yy, xx = np.meshgrid(np.arange(min_y, max_y + 1), np.arange(min_x, max_x + 1), indexing="ij")
            px = xx.astype(F) + F(0.5); py = yy.astype(F) + F(0.5)
            w0 = ((p1[1] - p2[1]) * (px - p2[0]) + (p2[0] - p1[0]) * (py - p2[1])) * inv
            w1 = ((p2[1] - p0[1]) * (px - p2[0]) + (p0[0] - p2[0]) * (py - p2[1])) * inv
            w2 = F(1.0) - w0 - w1
            m = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
            out = fb_rgba[min_y:max_y + 1, min_x:max_x + 1]
            for k in range(3):
                v = as_u8(F(col[i0, k]) * w0 + F(col[i1, k]) * w1 + F(col[i2, k]) * w2)
                out[..., k][m] = v[m]
            out[..., 3][m] = 255


def fb_clear(w, h, color):
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    rgba[...] = np.array(list(color[:3]) + [255], dtype=np.uint8)
    z = np.full((h, w), np.finfo(np.float32).max, dtype=np.float32)
    return rgba, z


# ------------------------------------------------------------------------------------------
# Framebuffer::clear_gradient (render.rs:60-77) and the overlay line family (render.rs:684-872)
# ------------------------------------------------------------------------------------------
def fb_clear_gradient(w, h, top, bottom):
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    y = np.arange(h, dtype=np.float32)
    t = y / F(h - 1) if h > 1 else np.zeros(h, dtype=np.float32)          # :64
    t = np.clip(t, F(0.0), F(1.0))                                         # Color::lerp, types.rs:811-820
    inv_t = F(1.0) - t
    for k in range(3):
        rgba[..., k] = as_u8(F(top[k]) * inv_t + F(bottom[k]) * t)[:, None]
    rgba[..., 3] = 0 if (len(top) > 3 and top[3] == ERASE) else 255        # lerp keeps self.blend; to_bytes :829-832
    z = np.full((h, w), np.finfo(np.float32).max, dtype=np.float32)
    return rgba, z


LINE_2D, LINE_2D_ALPHA, LINE_3D, LINE_3D_OVERLAY, LINE_3D_ALPHA, LINE_CIRCLE, LINE_CIRCLE_ALPHA, LINE_FILLED_RECT, LINE_THICK = range(9)


def _area_points(l, w, h):
    """Pixels of the filled primitives (render.rs:631-644, :670-682, :875-938, :954-972) as integer arrays."""
    kind = int(l["kind"])
    x0, y0, x1, y1 = int(l["x0"]), int(l["y0"]), int(l["x1"]), int(l["y1"])
    none = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    if kind in (LINE_CIRCLE, LINE_CIRCLE_ALPHA):
        r = x1
        ys = np.arange(max(y0 - r, 0), min(y0 + r, h - 1) + 1); xs = np.arange(max(x0 - r, 0), min(x0 + r, w - 1) + 1)
        if not len(ys) or not len(xs):
            return none
        yy, xx = np.meshgrid(ys, xs, indexing="ij")
        m = (xx - x0) ** 2 + (yy - y0) ** 2 <= r * r
        return xx[m], yy[m]
    if kind == LINE_FILLED_RECT:
        ys = np.arange(max(min(y0, y1), 0), min(max(y0, y1), h - 1) + 1); xs = np.arange(max(min(x0, x1), 0), min(max(x0, x1), w - 1) + 1)
        if not len(ys) or not len(xs):
            return none
        yy, xx = np.meshgrid(ys, xs, indexing="ij")
        return xx.reshape(-1), yy.reshape(-1)
    # draw_thick_line with thickness > 1: a quad around the segment, tested at pixel centres
    dx, dy = F(x1 - x0), F(y1 - y0)
    ln = np.sqrt(dx * dx + dy * dy)
    if ln < F(0.001):
        return none
    half = F(l["z0"]) * F(0.5)
    px, py = -dy / ln * half, dx / ln * half
    c = [(F(x0) + px, F(y0) + py), (F(x0) - px, F(y0) - py), (F(x1) - px, F(y1) - py), (F(x1) + px, F(y1) + py)]
    cxs = np.array([k[0] for k in c], dtype=F); cys = np.array([k[1] for k in c], dtype=F)
    bx0 = max(int(as_i32(cxs.min(keepdims=True))[0]), 0); bx1 = min(int(as_i32(cxs.max(keepdims=True))[0]), w - 1)
    by0 = max(int(as_i32(cys.min(keepdims=True))[0]), 0); by1 = min(int(as_i32(cys.max(keepdims=True))[0]), h - 1)
    if bx0 > bx1 or by0 > by1:
        return none
    yy, xx = np.meshgrid(np.arange(by0, by1 + 1), np.arange(bx0, bx1 + 1), indexing="ij")
    p0, p1 = xx.astype(F) + F(0.5), yy.astype(F) + F(0.5)
    inside = np.ones(xx.shape, dtype=bool)
    for i in range(4):
        a, b = c[i], c[(i + 1) % 4]
        cross = (b[0] - a[0]) * (p1 - a[1]) - (b[1] - a[1]) * (p0 - a[0])
        inside &= ~(cross < 0)
    return xx[inside], yy[inside]


def _line_points(x0, y0, x1, y1):
    """The pixel sequence and the `step` counter of the reference's Bresenham loop (:768-817; the 2D variants walk the
    same points and have no step counter), as integer arrays."""
    dx, dy = abs(x1 - x0), -abs(y1 - y0)
    sx, sy = (1 if x0 < x1 else -1), (1 if y0 < y1 else -1)
    err, x, y, step = dx + dy, x0, y0, 0
    xs, ys, steps = [], [], []
    while True:
        xs.append(x); ys.append(y); steps.append(step)
        if x == x1 and y == y1:
            break
        e2 = 2 * err
        if e2 >= dy:
            err += dy; x += sx; step += 1
        if e2 <= dx:
            err += dx; y += sy
            if e2 < dy:
                step += 1
    return np.array(xs), np.array(ys), np.array(steps), max(dx, max(-dy, 1))


def draw_lines(fb_rgba, fb_z, lines):
    """lines: records with the fields of abi.LINE_DTYPE; drawn in order into fb_rgba (u8[h,w,4]); fb_z is only read."""
    h, w = fb_z.shape
    for l in lines:
        kind, mode, alpha = int(l["kind"]), int(l["mode"]), int(l["alpha"])
        if kind in (LINE_CIRCLE, LINE_CIRCLE_ALPHA, LINE_FILLED_RECT) or (kind == LINE_THICK and int(as_i32(np.array([l["z0"]], dtype=F))[0]) > 1):
            xs, ys = _area_points(l, w, h)
            steps, total = np.zeros(len(xs), dtype=np.int64), 1
        else:                                                                      # every line kind; draw_thick_line(thickness <= 1) = draw_line
            xs, ys, steps, total = _line_points(int(l["x0"]), int(l["y0"]), int(l["x1"]), int(l["y1"]))
            on = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
            xs, ys, steps = xs[on], ys[on], steps[on]
        if kind in (LINE_3D, LINE_3D_OVERLAY, LINE_3D_ALPHA):
            z0, z1 = F(l["z0"]), F(l["z1"])
            if kind == LINE_3D_ALPHA:
                z0, z1 = z0 * F(0.995), z1 * F(0.995)                      # DEPTH_BIAS :826-828
            t = steps.astype(np.float32) / F(total)                        # step counts < 2^24 are exact in f32
            z = z0 + t * (z1 - z0)
            zb = fb_z[ys, xs]
            ok = (z < zb) if kind == LINE_3D else (z <= zb)
            xs, ys = xs[ok], ys[ok]
        rgb = np.array(l["rgb"], dtype=np.int64)
        back = fb_rgba[ys, xs, :3].astype(np.int64)
        if kind in (LINE_2D_ALPHA, LINE_3D_ALPHA, LINE_CIRCLE_ALPHA):      # set_pixel_alpha :646-667
            out, a = (rgb * alpha + back * (255 - alpha)) // 255, 255
        elif kind == LINE_2D and mode == ERASE:                            # Color::TRANSPARENT, types.rs:920-923
            out, a = np.zeros_like(back), 0
        elif kind == LINE_2D and mode != OPAQUE:                           # set_pixel_blended :313-333
            out, a = blend888(np.broadcast_to(rgb, back.shape), back, mode), 255
        else:                                                              # set_pixel :301-310
            out, a = np.broadcast_to(rgb, back.shape), (0 if int(l["blend"]) == ERASE else 255)
        fb_rgba[ys, xs, :3] = out.astype(np.uint8)
        fb_rgba[ys, xs, 3] = a


# ------------------------------------------------------------------------------------------
# Framebuffer::render_skybox step 2 (stars): render.rs:175-235, from the star direction on
# ------------------------------------------------------------------------------------------
def render_stars(fb_rgba, stars, camera, size):
    """stars: records with dir f32[3], rgb u8[3] (abi.STAR_DTYPE), in the reference's loop order."""
    H, W = fb_rgba.shape[:2]
    d = np.asarray(stars["dir"], dtype=F).reshape(-1, 3)
    with np.errstate(all="ignore"):
        cam_space = perspective_transform(d * F(10000.0), camera)                    # :178
        vs = (F(min(W, H)) / F(2.0)) * F(0.75)                                       # math.rs:117-136
        denom = cam_space[:, 2] + F(5.0)
        tiny = np.abs(denom) < F(0.001)
        sx = as_i32(np.where(tiny, F(W) / F(2.0), (cam_space[:, 0] * F(4.0)) / denom * vs + (F(W) / F(2.0))))
        sy = as_i32(np.where(tiny, F(H) / F(2.0), (cam_space[:, 1] * F(4.0)) / denom * vs + (F(H) / F(2.0))))
    s = int(as_i32(np.array([fmax(F(size), F(1.0))], dtype=F))[0])                   # :204
    rgb = np.asarray(stars["rgb"], dtype=np.uint8).reshape(-1, 3)
    rings = [(1.0, [(0, 0)])]
    if s >= 2:
        rings.append((0.7, [(-1, 0), (1, 0), (0, -1), (0, 1)]))
    if s >= 3:
        rings.append((0.4, [(-2, 0), (2, 0), (0, -2), (0, 2)]))
    for i in np.nonzero(cam_space[:, 2] > F(0.1))[0]:                                # :180, stars in order
        for k, offs in rings:
            c = rgb[i] if k == 1.0 else as_u8(rgb[i].astype(F) * F(k))
            for ox, oy in offs:
                x, y = int(sx[i]) + ox, int(sy[i]) + oy
                if 0 <= x < W and 0 <= y < H:
                    fb_rgba[y, x, :3] = c
                    fb_rgba[y, x, 3] = 255


# ------------------------------------------------------------------------------------------
# render_asset_parts' per-object vertex transform: src/scene.rs:121-160
# ------------------------------------------------------------------------------------------
def place_vertices(vertices, facing, cos_f, sin_f, world_pos):
    """vertices: record array with pos / normal f32[3]; returns the transformed copy (or an unchanged copy when the
    reference's has_transform is false)."""
    out = vertices.copy()
    wp = np.asarray(world_pos, dtype=F)
    if not (abs(F(facing)) > F(0.0001) or (np.abs(wp) > F(0.0001)).any()):
        return out
    c, s = F(cos_f), F(sin_f)
    p = np.asarray(vertices["pos"], dtype=F); n = np.asarray(vertices["normal"], dtype=F)
    out["pos"][:, 0] = (p[:, 0] * c - p[:, 2] * s) + wp[0]
    out["pos"][:, 1] = p[:, 1] + wp[1]
    out["pos"][:, 2] = (p[:, 0] * s + p[:, 2] * c) + wp[2]
    out["normal"][:, 0] = n[:, 0] * c - n[:, 2] * s
    out["normal"][:, 2] = n[:, 0] * s + n[:, 2] * c
    return out


# ------------------------------------------------------------------------------------------
# wireframe phase of render_mesh_15 / render_mesh: render.rs:2574-2635 (and :2195-2256)
# ------------------------------------------------------------------------------------------
_LINE_REC = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"), ("z0", "<f4"), ("z1", "<f4"),
                      ("rgb", "u1", 3), ("blend", "u1"), ("kind", "u1"), ("mode", "u1"), ("alpha", "u1"), ("_pad", "u1")])


def _unique_edges(tris):
    """First occurrence of every end-point pair, in order (the reference searches its list linearly; a set of the
    integer end points keeps the same ones)."""
    seen, out = set(), []
    for v1, v2, v3 in tris:
        for a, b in ((v1, v2), (v2, v3), (v3, v1)):
            pa = (int(as_i32(np.array([a[0]], dtype=F))[0]), int(as_i32(np.array([a[1]], dtype=F))[0]), F(a[2]))
            pb = (int(as_i32(np.array([b[0]], dtype=F))[0]), int(as_i32(np.array([b[1]], dtype=F))[0]), F(b[2]))
            e = (pa, pb) if (pa[0], pa[1]) < (pb[0], pb[1]) else (pb, pa)           # tuple order, as in Rust
            key = (e[0][0], e[0][1], e[1][0], e[1][1])
            if key not in seen:
                seen.add(key)
                out.append(e)
    return out


def _wireframe_phase(fb_rgba, fb_z, wire, settings):
    def lines(edges, kind, rgb):
        ln = np.zeros(len(edges), dtype=_LINE_REC)
        for i, (a, b) in enumerate(edges):
            ln[i] = (a[0], a[1], b[0], b[1], a[2], b[2], rgb, OPAQUE, kind, OPAQUE, 255, 0)
        return ln
    if settings.backface_cull and settings.backface_wireframe:                      # :2577-2603, draw_line_3d
        draw_lines(fb_rgba, fb_z, lines(_unique_edges(wire["back"]), LINE_3D, (80, 80, 100)))
    if settings.wireframe_overlay and wire["front"]:                                # :2606-2633, draw_line
        draw_lines(fb_rgba, fb_z, lines(_unique_edges(wire["front"]), LINE_2D, (200, 200, 220)))
