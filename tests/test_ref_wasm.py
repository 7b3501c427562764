"""The oracle against the reference's own COMPILED code.

tests/golden/ref_wasm/ holds results produced by executing /root/reference/docs/bonnie-32.wasm (the authors' wasm32
release build of the application; `render_mesh_15`, `rasterize_triangle_15`, `project_fixed`, ... are separate named
functions in it) in the interpreter under oracle/wasm/ — see tests/golden/make_ref_wasm.py.  The binary is crate
version 0.1.8 and the source tree 0.1.11; four behavioural differences between them were traced in the decompiled
code (oracle/wasm/DRIFT.md) and are switched back in the oracle by `b32o_set_compat` for this comparison only.
Everything else — transform, snap, near-plane / back-face / fog cull, fog colours, partition + stable sort, edge
stepping, inside test, depth, affine and perspective-correct UV, texel fetch + transparency rules, modulate, Gouraud /
flat multi-light shading, dither, blend modes, x-ray, z-buffer rules, wireframe phase, panics — must agree bit for bit,
framebuffer and z-buffer.
"""
import json
import os

import numpy as np
import pytest

import refbin_cases

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "ref_wasm", "render_mesh_15.json")))
SMALL = {s.name: s for s in refbin_cases.small_scenes()}


@pytest.fixture
def compat_oracle(oracle):
    oracle.lib().b32o_set_compat(refbin_cases.COMPAT_0_1_8)
    yield oracle
    oracle.lib().b32o_set_compat(0)


def _check(oracle, sc):
    rec = FIX["scenes"][sc.name]
    assert rec["inputs"] == refbin_cases.inputs_digest(sc), "scene generator changed: regenerate the fixture"
    rgba, z, tm, rc = oracle.render_scene(sc)
    if "trap" in rec:                       # the reference panicked (NaN sort key / index out of range)
        assert rc != 0, "reference panics, oracle returned OK"
        return
    assert rc == 0
    assert tm["triangles_drawn"] == rec["drawn"]
    a, b = refbin_cases.frame_digest(rgba, z)
    assert a == rec["rgba"], "framebuffer differs from the reference binary"
    assert b == rec["z"], "z-buffer differs from the reference binary"


def test_fixture_covers_every_scene():
    assert set(SMALL) <= set(FIX["scenes"])
    assert len(FIX["wasm_sha256"]) == 64


@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_matches_reference_binary(compat_oracle, name):
    _check(compat_oracle, SMALL[name])


@pytest.mark.parametrize("name", [s.name for s in refbin_cases.big_scenes()])
def test_oracle_matches_reference_binary_100k(compat_oracle, name):
    sc = {s.name: s for s in refbin_cases.big_scenes()}[name]
    if sc.name not in FIX["scenes"]:
        pytest.skip("fixture generated without --big")
    _check(compat_oracle, sc)


FIX8 = json.load(open(os.path.join(HERE, "golden", "ref_wasm", "render_mesh.json")))
SMALL8 = {s.name: s for s in refbin_cases.small_scenes888()}


@pytest.mark.parametrize("name", sorted(SMALL8))
def test_oracle_render_mesh_rgb888_matches_reference_binary(compat_oracle, name):
    """The RGB888 sibling `render_mesh` -> `rasterize_triangle` (render.rs:1971-2259, :1202-1433) of the binary."""
    sc = SMALL8[name]
    rec = FIX8["scenes"][name]
    assert rec["inputs"] == refbin_cases.inputs_digest(sc), "scene generator changed: regenerate the fixture"
    rgba, z, tm, rc = compat_oracle.render_scene888(sc)
    if "trap" in rec:
        assert rc != 0
        return
    assert rc == 0 and tm["triangles_drawn"] == rec["drawn"]
    a, b = refbin_cases.frame_digest(rgba, z)
    assert a == rec["rgba"], "framebuffer differs from the reference binary"
    assert b == rec["z"], "z-buffer differs from the reference binary"


def test_full_frames_kept_for_debugging(compat_oracle):
    fr = np.load(os.path.join(HERE, "golden", "ref_wasm", "frames.npz"))
    for name in ("c1_single_triangle", "c2_1000_tris_64x64_idx8", "gouraud_lights", "mixed_zbuffer"):
        rgba, z, tm, rc = compat_oracle.render_scene(SMALL[name])
        assert np.array_equal(rgba, fr[name + "/rgba"])
        assert np.array_equal(z.view(np.uint32), fr[name + "/z"].view(np.uint32))


def test_compat_switches_are_off_by_default(oracle):
    assert oracle.lib().b32o_get_compat() == 0


# ---- direct calls of named functions of the binary ------------------------------------------------------------------
def test_project_fixed_matches_reference_binary(compat_oracle):
    """fixed::project_fixed (fixed.rs:424-441) on 24 000 vertices x 8 cameras x 3 sizes, incl. non-finite / huge / denormal
    inputs and the |denom| < 256 early-out.  (The binary's divide is exact, 0.1.11's is div_unr: COMPAT_DIV_EXACT.)"""
    import ctypes as C
    import refbin_funcs
    fx = np.load(os.path.join(HERE, "golden", "ref_wasm", "functions.npz"))
    world, cam_idx, size_idx = refbin_funcs.project_inputs()
    cams = [c.to_abi() for c in refbin_funcs.cameras()]
    lib = compat_oracle.lib()
    sx, sy, d = C.c_int32(), C.c_int32(), C.c_float()
    got = np.empty((len(world), 2), np.int32); gd = np.empty(len(world), np.float32)
    for i in range(len(world)):
        wv = (C.c_float * 3)(*[float(x) for x in world[i]])
        w, h = refbin_funcs.SIZES[size_idx[i]]
        lib.b32o_project_fixed(wv, C.byref(cams[cam_idx[i]]), C.c_uint32(w), C.c_uint32(h), C.byref(sx), C.byref(sy), C.byref(d))
        got[i] = (sx.value, sy.value); gd[i] = d.value
    assert np.array_equal(got[:, 0], fx["project_sx"]) and np.array_equal(got[:, 1], fx["project_sy"])
    assert np.array_equal(gd.view(np.uint32), fx["project_depth"].view(np.uint32))
    # the early-out rows really took the early-out: centre of the framebuffer
    for row, early in ((5, True), (6, True), (7, False), (8, True), (9, False)):
        w, h = refbin_funcs.SIZES[size_idx[row]]
        assert (tuple(got[row]) == (w // 2, h // 2)) == early, row


def test_shade_multi_light_color_matches_reference_binary(oracle):
    """render::shade_multi_light_color (render.rs:1013-1071): Directional + Point lights, colours, disabled lights, zero
    radius, dist < 0.001, zero / NaN normals — 6 000 evaluations, bit for bit (no compat switch involved)."""
    import ctypes as C
    from bonnie32_b200 import abi
    import refbin_funcs
    fx = np.load(os.path.join(HERE, "golden", "ref_wasm", "functions.npz"))
    normal, pos, set_idx, ambient = refbin_funcs.shade_inputs()
    sets = []
    for ls in refbin_funcs.light_sets():
        arr = (abi.Light * max(1, len(ls)))()
        for k, l in enumerate(ls):
            arr[k] = l.to_abi()
        sets.append((arr, len(ls)))
    lib = oracle.lib()
    out = (C.c_float * 3)()
    got = np.empty((len(normal), 3), np.float32)
    for i in range(len(normal)):
        n = (C.c_float * 3)(*[float(x) for x in normal[i]]); p = (C.c_float * 3)(*[float(x) for x in pos[i]])
        arr, cnt = sets[set_idx[i]]
        lib.b32o_shade_multi_light(n, p, arr, C.c_uint32(cnt), C.c_float(float(ambient[i])), out)
        got[i] = out[:]
    same = (got.view(np.uint32) == fx["shade"].view(np.uint32)) | (np.isnan(got) & np.isnan(fx["shade"]))
    assert same.all(), np.argwhere(~same)[:5]


def test_sample_level_geometry_matches_reference_binary():
    """The C3 fixtures' room triangles (tests/golden/c3_*.npz, produced by bonnie-32_b200/levels.py from the reference's
    level files) against the reference binary's own `load_level_from_str` + `Room::add_*_to_render_data`
    (src/world/geometry.rs:2839-3352): positions, UVs, normals, colours, indices, black_transparent, bit for bit.
    Texture ids depend on the resolver and are left out of the digest; so are the UVs of rooms that use textures which are
    not 64 texels wide (the UV scale is 32 / texture width, and the binary was driven with a resolver that misses)."""
    import c3
    fix = json.load(open(os.path.join(HERE, "golden", "ref_wasm", "levels.json")))["levels"]
    paths = c3.scene_paths()
    assert len(paths) == 6
    for p in paths:
        sc = c3.load_scene(p)
        all64 = all(t.width == 64 for t in sc.textures)
        want = fix[sc.name]
        assert len(want) == len(sc.rooms)
        for rc, w in zip(sc.rooms, want):
            assert len(rc.vertices) == w["vertices"] and len(rc.faces) == w["faces"]
            d = refbin_cases.geometry_digest(rc.vertices["pos"], rc.vertices["uv"], rc.vertices["normal"], rc.vertices["rgba"], rc.faces["v"],
                                             ((rc.faces["flags"] >> 19) & 1).astype(np.uint8), with_uv=all64)
            assert d == w["sha256" if all64 else "sha256_no_uv"], sc.name


@pytest.mark.parametrize("name,w,h,seed,n", __import__("refbin_prims").CASES, ids=lambda v: v if isinstance(v, str) else None)
def test_overlay_primitives_match_reference_binary(oracle, name, w, h, seed, n):
    """b32o_draw_lines vs the binary's own Framebuffer::draw_line_3d_impl / draw_circle / draw_thick_line, called once per
    primitive in list order (tests/golden/make_ref_wasm_prims.py).  No compat switch is involved (RGB888 writers)."""
    import hashlib
    import json
    import refbin_prims
    fix = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_wasm", "prims.json")))["cases"][name]
    rgba, z = refbin_prims.background(w, h, seed)
    lines = refbin_prims.primitives(w, h, seed, n)
    assert hashlib.sha256(rgba.tobytes() + z.tobytes() + lines.tobytes()).hexdigest() == fix["inputs"]
    z0 = z.copy()
    assert oracle.draw_lines(rgba, z, lines) == 0
    assert np.array_equal(z.view(np.uint32), z0.view(np.uint32))
    assert hashlib.sha256(rgba.tobytes()).hexdigest() == fix["rgba"], "pixels differ from the reference binary"


def _skybox_fixture():
    here = os.path.dirname(__file__)
    meta = json.load(open(os.path.join(here, "golden", "ref_wasm", "skybox.json")))["cases"]
    return meta, np.load(os.path.join(here, "golden", "ref_wasm", "skybox.npz"))


@pytest.mark.parametrize("key", sorted(_skybox_fixture()[0]))
def test_skybox_passes_match_reference_binary(oracle, key):
    """b32o_render_skybox_mesh + b32o_render_stars vs the binary's own Framebuffer::render_skybox (render.rs:81-299) on the
    sphere mesh / star list the binary built for that frame (tests/golden/make_ref_wasm_skybox.py).  No compat switch."""
    import hashlib
    import cases
    meta, arr = _skybox_fixture()
    m = meta[key]
    cam = cases._rotated_camera(m["camera"][0], m["camera"][1], m["camera"][2])
    fb = np.zeros((m["height"], m["width"], 4), np.uint8)                       # Framebuffer::new
    assert oracle.render_skybox_mesh(fb, arr[key + "_verts"], arr[key + "_faces"], cam) == 0
    if m["n_stars"]:
        before = fb.copy()
        assert oracle.render_stars(fb, arr[key + "_stars"], cam, m["star_size"]) == 0
        assert (fb != before).any(), "the star pass drew nothing: the fixture does not exercise it"
    if key + "_frame" in arr:
        bad = (arr[key + "_frame"] != fb).any(-1)
        assert not bad.any(), f"{bad.sum()} pixels differ from the reference binary, first at {np.argwhere(bad)[0][::-1]}"
    assert hashlib.sha256(fb.tobytes()).hexdigest() == m["rgba"], "frame differs from the reference binary"


def test_set_pixel_blended_15_matches_reference_binary(oracle):
    """blend_rgb555 (render.rs:1093-1145) through the binary's Framebuffer::set_pixel_blended_15: 30 000 (Color15, back
    pixel, BlendMode) triples, all six modes.  What survives in the binary under this name is the blend tail only: its
    callers have already decided that the texel blends, so bit 15 is not looked at and the br_table's default arm
    (mode Opaque, never passed in practice) is Average (DRIFT.md item 7).  Its Color15::r8() is `v << 3` (item 3).  What
    this pins is the arithmetic of every blend mode on arbitrary back pixels, incl. the `>> 3` of a non-multiple-of-8 back."""
    import ctypes as C
    import refbin_funcs
    fx = np.load(os.path.join(HERE, "golden", "ref_wasm", "functions.npz"))["blend15"]
    c15, back, mode = refbin_funcs.blend_inputs()
    lib = oracle.lib()
    out = (C.c_uint8 * 3)()
    got = np.empty((len(c15), 4), np.uint8)
    for i in range(len(c15)):
        c = int(c15[i])
        f8 = [((c >> 10) & 31) << 3, ((c >> 5) & 31) << 3, (c & 31) << 3]              # the binary's r8 / g8 / b8
        eff = int(mode[i]) if mode[i] != 0 else 1                                      # the br_table's default arm
        if eff:
            lib.b32o_blend_rgb555(C.c_uint8(f8[0]), C.c_uint8(f8[1]), C.c_uint8(f8[2]), C.c_uint8(int(back[i, 0])), C.c_uint8(int(back[i, 1])),
                                  C.c_uint8(int(back[i, 2])), C.c_uint32(eff), out)
            got[i, :3] = out[:]
        else:
            got[i, :3] = f8
        got[i, 3] = 255
    bad = (got != fx).any(1)
    assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:3].ravel(), got[bad][:3], fx[bad][:3])


def test_framebuffer_clear_matches_reference_binary(oracle):
    """Framebuffer::clear (render.rs:36-45): colour bytes (alpha 0 only for an Erase-blend colour) and depth f32::MAX."""
    from bonnie32_b200 import abi
    fx = np.load(os.path.join(HERE, "golden", "ref_wasm", "functions.npz"))["clear"]
    for row, word in zip(fx, (0x1C161400, 0xFF000005, 0x01020302, 0x00000000)):
        blend, r, g, b = word & 255, (word >> 8) & 255, (word >> 16) & 255, (word >> 24) & 255
        rgba = np.zeros((1, 2, 4), np.uint8); z = np.zeros((1, 2), np.float32)
        oracle.lib().b32o_fb_clear(rgba.ctypes.data, z.ctypes.data, 2, 1, r, g, b, 0 if blend == abi.BLEND_ERASE else 255)
        assert np.array_equal(rgba.reshape(-1), row[:8]) and np.array_equal(z.view(np.uint8).reshape(-1), row[8:])


@pytest.mark.parametrize("name,w,h,seed,n", __import__("refbin_prims").CLIPPED, ids=lambda v: v if isinstance(v, str) else None)
def test_plain_draw_line_matches_reference_binary(oracle, name, w, h, seed, n):
    """Framebuffer::draw_line (render.rs:715-751, B32_LINE_2D) through the one caller the binary keeps as a function,
    draw::draw_3d_line_clipped (draw.rs:12-66): world-space segments in front of an identity camera; the end points the
    oracle is given come from a f32 restatement of world_to_screen (math.rs:503-534) in tests/refbin_prims.py."""
    import hashlib
    import json
    import refbin_prims
    from bonnie32_b200 import abi
    fix = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_wasm", "prims.json")))["clipped"][name]
    rgba, z = refbin_prims.background(w, h, seed)
    p0, p1, rgb, ends = refbin_prims.clipped_segments(w, h, seed, n)
    assert hashlib.sha256(rgba.tobytes() + p0.tobytes() + p1.tobytes() + rgb.tobytes()).hexdigest() == fix["inputs"]
    lines = np.zeros(n, dtype=abi.LINE_DTYPE)
    lines["kind"] = abi.LINE_2D
    lines["mode"] = abi.BLEND_OPAQUE
    lines["x0"], lines["y0"], lines["x1"], lines["y1"] = ends[:, 0, 0], ends[:, 0, 1], ends[:, 1, 0], ends[:, 1, 1]
    lines["rgb"] = rgb
    assert oracle.draw_lines(rgba, z, lines) == 0
    assert hashlib.sha256(rgba.tobytes()).hexdigest() == fix["rgba"], "pixels differ from the reference binary"


# ---- Spot lights (render.rs:1038-1059): f32::acos of the shipped build, the lighting function, whole frames ---------------
SPOT = json.load(open(os.path.join(HERE, "golden", "ref_wasm", "spot.json")))
SPOT15 = {s.name: s for s in refbin_cases.spot_scenes()}
SPOT888 = {s.name: s for s in refbin_cases.spot_scenes888()}


def test_acosf_matches_reference_binary(oracle):
    """`f32::acos` (render.rs:1047) is the `acosf` symbol: compiler_builtins' libm port in the reference's own wasm build.
    The oracle's restatement against that function on 400 000+ arguments (every branch, every float within 64 ulps of the
    branch boundaries, NaN / out-of-domain inputs), bit for bit; the numpy model on a sample."""
    import ctypes as C
    import refbin_funcs
    from oracle import pymodel
    x = refbin_funcs.acosf_inputs()
    want = np.load(os.path.join(HERE, "golden", "ref_wasm", "spot.npz"))["acosf"]
    assert len(want) == len(x)
    got = np.empty_like(x)
    oracle.lib().b32o_acosf(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_uint32(len(x)))
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (x[~same][:5], got[~same][:5], want[~same][:5])
    err = np.abs(got[np.isfinite(got)].astype(np.float64) - np.arccos(x[np.isfinite(got)].astype(np.float64)))
    assert err.max() < 2.5e-7                                  # and it is an arc cosine
    idx = np.concatenate([np.arange(0, 400000, 97), np.arange(400000, len(x))])
    py = np.array([pymodel.ref_acosf(v) for v in x[idx]], np.float32)
    same = (py.view(np.uint32) == want[idx].view(np.uint32)) | (np.isnan(py) & np.isnan(want[idx]))
    assert same.all(), x[idx][~same][:5]


def test_shade_multi_light_color_spot_matches_reference_binary(oracle):
    """shade_multi_light_color with Spot lights in the list (cone test, edge falloff, zero / un-normalised directions whose
    acos is NaN, cones wider than pi, negative angles, disabled lights) — 8 000 evaluations of the binary, bit for bit; the
    numpy model on every 8th."""
    import ctypes as C
    from bonnie32_b200 import abi
    from oracle import pymodel
    import refbin_funcs
    want = np.load(os.path.join(HERE, "golden", "ref_wasm", "spot.npz"))["shade"]
    normal, pos, set_idx, ambient = refbin_funcs.spot_shade_inputs()
    light_sets = refbin_funcs.spot_light_sets()
    sets = []
    for ls in light_sets:
        arr = (abi.Light * max(1, len(ls)))()
        for k, l in enumerate(ls):
            arr[k] = l.to_abi()
        sets.append((arr, len(ls)))
    lib = oracle.lib()
    out = (C.c_float * 3)()
    got = np.empty((len(normal), 3), np.float32)
    for i in range(len(normal)):
        n = (C.c_float * 3)(*[float(v) for v in normal[i]]); p = (C.c_float * 3)(*[float(v) for v in pos[i]])
        arr, cnt = sets[set_idx[i]]
        lib.b32o_shade_multi_light(n, p, arr, C.c_uint32(cnt), C.c_float(float(ambient[i])), out)
        got[i] = out[:]
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), np.argwhere(~same)[:5]
    assert (want != want[:, :1]).any()                 # coloured contributions are in there (a NaN total leaves as 1.0: f32::min)
    with np.errstate(all="ignore"):
        for i in range(0, len(normal), 8):
            py = np.asarray(pymodel.shade_multi_light_color(normal[i], pos[i], light_sets[set_idx[i]], ambient[i]), np.float32)
            ok = (py.view(np.uint32) == want[i].view(np.uint32)) | (np.isnan(py) & np.isnan(want[i]))
            assert ok.all(), i


def test_spot_fixture_covers_every_scene():
    assert set(SPOT["scenes"]) == set(SPOT15) | set(SPOT888)


@pytest.mark.parametrize("name", sorted(SPOT15) + sorted(SPOT888))
def test_oracle_spot_scenes_match_reference_binary(compat_oracle, name):
    """Whole frames lit by Spot lights through the binary's render_mesh_15 / render_mesh: framebuffer and z-buffer."""
    rgb888 = name in SPOT888
    sc = (SPOT888 if rgb888 else SPOT15)[name]
    rec = SPOT["scenes"][name]
    assert rec["inputs"] == refbin_cases.inputs_digest(sc), "scene generator changed: regenerate the fixture"
    rgba, z, tm, rc = compat_oracle.render_scene888(sc) if rgb888 else compat_oracle.render_scene(sc)
    if "trap" in rec:
        assert rc != 0
        return
    assert rc == 0 and tm["triangles_drawn"] == rec["drawn"]
    a, b = refbin_cases.frame_digest(rgba, z)
    assert a == rec["rgba"], "framebuffer differs from the reference binary"
    assert b == rec["z"], "z-buffer differs from the reference binary"
