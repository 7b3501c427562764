"""CPU tests of the oracle: the reference's own exact unit-test facts (fixed.rs:477-548), spec
constants, agreement with the independent numpy model (oracle/pymodel.py) and with the committed
golden fixtures (tests/golden, produced by the numpy model)."""
import hashlib
import json
import os

import numpy as np
import pytest

import bonnie32_b200 as pkg
from bonnie32_b200 import abi, scenes
from oracle import pymodel
import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
HASHES = json.load(open(os.path.join(GOLDEN, "hashes.json")))


# ---- fixed.rs:477-548: the only exact facts the reference's tests pin --------------------------
def test_fixed32_precision(oracle):                    # fixed.rs:478-489
    L = oracle.lib()
    assert L.b32o_fixed_from_f32(1.0) == 4096
    assert L.b32o_fixed_from_f32(0.5) == 2048
    assert L.b32o_fixed_from_f32(0.001) > 0


def test_fixed32_mul(oracle):                          # fixed.rs:492-497
    L = oracle.lib()
    r = L.b32o_fixed_mul(L.b32o_fixed_from_f32(2.0), L.b32o_fixed_from_f32(3.0))
    assert abs(r / 4096.0 - 6.0) < 0.01


def test_unr_division(oracle):                         # fixed.rs:500-531 (tolerances) + survey transcription values
    L = oracle.lib()
    fx = L.b32o_fixed_from_f32
    assert abs(L.b32o_div_unr(fx(10.0), fx(3.0)) / 4096.0 - 10.0 / 3.0) < 0.1
    assert abs(L.b32o_div_unr(fx(10.0), fx(2.0)) / 4096.0 - 5.0) < 0.01
    assert abs(L.b32o_div_unr(fx(-6.0), fx(2.0)) / 4096.0 + 3.0) < 0.01
    assert abs(L.b32o_div_unr(fx(7.5), fx(1.0)) / 4096.0 - 7.5) < 0.1
    assert L.b32o_div_unr(40960, 12288) == 13653
    assert L.b32o_div_unr(40960, 8192) == 20480
    assert L.b32o_div_unr(-24576, 8192) == -12288
    assert L.b32o_div_unr(30720, 4096) == 30720
    assert L.b32o_div_unr(123, 0) == 0


def test_projection_outputs_integers(oracle):          # fixed.rs:534-548
    import ctypes as C
    cam = abi.Camera()
    cam.basis_x[:] = [1, 0, 0]; cam.basis_y[:] = [0, 1, 0]; cam.basis_z[:] = [0, 0, 1]
    w = (C.c_float * 3)(1.234, 2.567, 5.0)
    sx, sy, d = C.c_int32(), C.c_int32(), C.c_float()
    oracle.lib().b32o_project_fixed(w, C.byref(cam), C.c_uint32(320), C.c_uint32(240), C.byref(sx), C.byref(sy), C.byref(d))
    assert -1000 < sx.value < 1000 and -1000 < sy.value < 1000
    # hand computation: x = 160 + floor(1.234*4/10*90) region
    assert (sx.value, sy.value) == (204, 212)


def test_unr_table_formula(oracle):                    # fixed.rs:18-31
    L = oracle.lib()
    t = [L.b32o_unr_table(i) for i in range(257)]
    assert t[:4] == [255, 253, 251, 249] and t[256] == 0
    assert t == [max(0, ((0x40000 // (i + 0x100)) + 1) // 2 - 0x101) for i in range(257)]
    assert t == list(pymodel.UNR_TABLE)


def test_div_unr_matches_numpy_model(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(7)
    num = rng.integers(-2**31, 2**31, 20000, dtype=np.int64)
    den = rng.integers(-2**31, 2**31, 20000, dtype=np.int64)
    den[:2000] = rng.integers(-5000, 5000, 2000)
    num[2000:4000] = rng.integers(-5000, 5000, 2000)
    edge = np.array([0, 1, -1, 255, 256, 257, -256, 2**31 - 1, -2**31, 0x7FC0, 0x8000, 0xFFFF, 0x10000], dtype=np.int64)
    num = np.concatenate([num, np.repeat(edge, len(edge))]); den = np.concatenate([den, np.tile(edge, len(edge))])
    want = pymodel.div_unr(num, den)
    got = np.array([L.b32o_div_unr(int(a), int(b)) for a, b in zip(num, den)], dtype=np.int32)
    assert np.array_equal(got, want)


def test_dither_truth_table(oracle):                   # render.rs:1150-1182, exhaustive 256 x 16
    import ctypes as C
    out = (C.c_uint8 * 3)()
    M = [[-4, 0, -3, 1], [2, -2, 3, -1], [-3, 1, -4, 0], [3, -1, 2, -2]]
    for c in range(256):
        for y in range(4):
            for x in range(4):
                oracle.lib().b32o_dither_and_quantize(C.c_uint8(c), C.c_uint8(255 - c), C.c_uint8(c ^ 0x55), C.c_uint32(x + 4), C.c_uint32(y + 8), out)
                exp = [min(max((v + M[y][x]) >> 3, 0), 31) for v in (c, 255 - c, c ^ 0x55)]
                assert list(out) == exp


def test_blend_truth_table(oracle):                    # render.rs:1093-1145, all 6 modes x 32 x 32 five-bit pairs
    import ctypes as C
    out = (C.c_uint8 * 3)()
    for mode in range(6):
        for f in range(0, 256, 8):
            for b in range(0, 256, 8):
                oracle.lib().b32o_blend_rgb555(C.c_uint8(f), C.c_uint8(f | 7), C.c_uint8(f), C.c_uint8(b), C.c_uint8(b), C.c_uint8(b | 5),
                                               C.c_uint32(mode), out)
                want = pymodel.blend555(np.array([f, f | 7, f]), np.array([b, b, b | 5]), mode)
                assert list(out) == list(want), (mode, f, b)
                f5, b5 = f >> 3, b >> 3
                exp = {0: f5, 1: (b5 + f5) // 2, 2: min(b5 + f5, 31), 3: max(b5 - f5, 0), 4: min(b5 + f5 // 4, 31), 5: b5}[mode] << 3
                assert out[0] == exp


def test_texture_sample_wrap(oracle):                  # types.rs:671-681, X6: tiny negative u wraps to exactly 1.0
    import ctypes as C
    px = np.arange(8 * 4, dtype=np.uint16) + 1
    t = pkg.Texture15(8, 4, px)
    d, keep = t.to_abi()
    f = oracle.lib().b32o_texture_sample
    f.argtypes = [C.POINTER(abi.TexDesc), C.c_float, C.c_float]
    assert f(C.byref(d), 0.0, 0.0) == 1
    assert f(C.byref(d), -1e-9, 0.0) == 8            # rem_euclid -> 1.0 -> clamped to w-1
    assert f(C.byref(d), 0.999, 0.999) == 32
    assert f(C.byref(d), 1.5, -0.25) == px[3 * 8 + 4]
    assert f(C.byref(d), float("nan"), float("inf")) == 1
    got = [f(C.byref(d), float(u), float(v)) for u in np.linspace(-3, 3, 41) for v in np.linspace(-2, 2, 17)]
    uu, vv = np.meshgrid(np.linspace(-3, 3, 41).astype(np.float32), np.linspace(-2, 2, 17).astype(np.float32), indexing="ij")
    want = pymodel.sample(px.reshape(4, 8), uu.reshape(-1), vv.reshape(-1))
    assert got == list(want)


def test_camera_new_basis():                           # camera.rs:21-32, 76-91 and SURVEY §8d
    c = pkg.Camera()
    assert list(c.basis_x) == [-1.0, 0.0, 0.0] and list(c.basis_y) == [0.0, -1.0, 0.0] and list(c.basis_z) == [0.0, 0.0, 1.0]


def test_c1_screen_coordinates(oracle):                # (205,165) (115,165) (160,75)
    for fixed in (True, False):
        sc = scenes.scene_c1(fixed)
        scr, cam = oracle.transform(sc.vertices, sc.camera, sc.settings, 320, 240)
        assert scr.tolist() == [[205.0, 165.0, 8.0], [115.0, 165.0, 8.0], [160.0, 75.0, 8.0]]


# ---- oracle vs the independent numpy model, and vs the committed golden fixtures -----------------
def _digest(rgba, z, order):
    return {"rgba_sha256": hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest(),
            "z_sha256": hashlib.sha256(np.ascontiguousarray(z).tobytes()).hexdigest(),
            "triangles_drawn": len(order),
            "order_sha256": hashlib.sha256(np.asarray(order, dtype=np.uint32).tobytes()).hexdigest()}


SMALL = [scenes.scene_c1(True), scenes.scene_c1(False)] + cases.feature_scenes() + [cases.big_triangle_scene()]


@pytest.mark.parametrize("sc", SMALL, ids=[s.name for s in SMALL])
def test_oracle_equals_numpy_model_and_golden(oracle, sc):
    want, want_z, tm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0
    rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
    order2 = pymodel.render_mesh_15(rgba, z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    assert list(order) == order2
    assert np.array_equal(rgba, want)
    assert np.array_equal(z.view(np.uint32), want_z.view(np.uint32))
    assert _digest(want, want_z, order) == HASHES[sc.name]


SPOT_SMALL = [s for s in cases.spot_scenes() if s.name in ("spot_gouraud_all", "spot_flat_all", "spot_gouraud_mixed_nocull", "spot_mixed_blend_gouraud")]


@pytest.mark.parametrize("sc", SPOT_SMALL, ids=[s.name for s in SPOT_SMALL])
def test_oracle_equals_numpy_model_spot_lights(oracle, sc):
    """Spot lights (render.rs:1038-1059) through both restatements: the C++ oracle and the numpy model (its own ref_acosf)
    draw the same frame; the binary's frames pin both (tests/test_ref_wasm.py)."""
    want, want_z, tm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0
    rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
    order2 = pymodel.render_mesh_15(rgba, z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    assert list(order) == order2
    assert np.array_equal(rgba, want)
    assert np.array_equal(z.view(np.uint32), want_z.view(np.uint32))
    lit, _, _, _ = oracle.render_scene(cases._with(sc, sc.name + "_unlit", lights=[]))
    assert not np.array_equal(lit, want)                   # the lights do reach the frame


def test_scenes_exercise_their_feature(oracle):
    """Guards against vacuous parity: the feature scenes must really hit the code they name."""
    by = {s.name: s for s in cases.feature_scenes()}
    base, _, _, _ = oracle.render_scene(by["painter_idx8"])
    for name in ("nodither", "float_projection", "perspective_correct", "nocull_painter", "xray", "fog",
                 "gouraud_lights", "flat_lights", "mixed_painter", "untextured_vertex_colours"):
        img, _, _, _ = oracle.render_scene(by[name])
        assert not np.array_equal(img, base), name
    # the mixed scene has both passes
    m = by["mixed_zbuffer"]
    blends = (m.faces["flags"] >> 16) & 7
    assert (blends != 0).any() and (blends == 0).any()
    # the large-world scene produces edge values beyond 2^24 (forces the exact incremental path)
    big = by["rotated_camera_large_world"]
    scr, cam = oracle.transform(big.vertices, big.camera, big.settings, 320, 240)
    assert np.abs(scr[:, :2]).max() > 5000


@pytest.mark.parametrize("name", ["c2_1000_tris_64x64_idx8", "c2_1000_tris_64x64_idx8_zbuffer"])
def test_c2_golden_full_framebuffer(oracle, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scenes.scene_c2(use_zbuffer=name.endswith("zbuffer"))
    want, want_z, tm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0 and tm["triangles_drawn"] == len(g["order"])
    assert np.array_equal(order, g["order"])
    assert np.array_equal(want, g["rgba"]) and np.array_equal(want_z.view(np.uint32), g["z"].view(np.uint32))
    # the fixture and the hash file agree
    assert hashlib.sha256(g["rgba"].tobytes()).hexdigest() == HASHES[name]["rgba_sha256"]


@pytest.mark.parametrize("name,zbuf", [("c4_100000_tris_256x256_idx4", False), ("c4_100000_tris_256x256_idx4_zbuffer", True)])
def test_c4_golden_hash(oracle, name, zbuf):
    """BASELINE config 4 at full size: the oracle reproduces the numpy model's framebuffer."""
    sc = scenes.scene_c4(use_zbuffer=zbuf)
    want, want_z, tm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0
    assert _digest(want, want_z, order) == HASHES[name]


def test_oracle_panics_map_to_error_codes(oracle):
    sc = scenes.scene_c2(n_tris=50)
    f = sc.faces.copy(); f["v"][7, 1] = len(sc.vertices)
    rgba, z = pymodel.fb_clear(320, 240, sc.clear)
    rc, _, _ = oracle.render_mesh_15(rgba, z, sc.vertices, f, sc.textures, sc.camera, sc.settings)
    assert rc == abi.B32_ERR_OOB_INDEX
    with pytest.raises(pymodel.ReferencePanic):
        pymodel.render_mesh_15(rgba, z, sc.vertices, f, sc.textures, sc.camera, sc.settings)
    v = cases.nan_depth_vertices(sc, oracle)
    rc, _, _ = oracle.render_mesh_15(rgba, z, v, sc.faces, sc.textures, sc.camera, sc.settings)
    assert rc == abi.B32_ERR_NAN_DEPTH
    with pytest.raises(pymodel.ReferencePanic):
        pymodel.render_mesh_15(rgba, z, v, sc.faces, sc.textures, sc.camera, sc.settings)


def test_splitmix64_reference_values():
    """SplitMix64 known answers (seed 0: first outputs of the published algorithm)."""
    z = []
    x = 0
    for _ in range(3):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        t = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        t = ((t ^ (t >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z.append(t ^ (t >> 31))
    assert z[0] == 0xE220A8397B1DCDAF and z[1] == 0x6E789E6AA1B965F4
    u = scenes.splitmix64_u01(0, 3)
    assert [int(v * 2**24) for v in u] == [v >> 40 for v in z]


# ---- RGB888 sibling: render_mesh / rasterize_triangle (render.rs:1971-2259, 1202-1433) ---------------------
RGB888 = cases.rgb888_scenes()
with open(os.path.join(GOLDEN, "hashes_rgb888.json")) as _f:
    HASHES888 = json.load(_f)


@pytest.mark.parametrize("sc", RGB888, ids=[s.name for s in RGB888])
def test_rgb888_oracle_equals_numpy_model_and_golden(oracle, sc):
    want, want_z, tm, rc, order = oracle.render_scene888(sc, want_order=True)
    assert rc == 0 and tm["triangles_drawn"] == len(order)
    rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
    order2 = pymodel.render_mesh(rgba, z, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
    assert list(order) == order2
    assert np.array_equal(rgba, want)
    assert np.array_equal(z.view(np.uint32), want_z.view(np.uint32))
    assert _digest(want, want_z, order) == HASHES888[sc.name]


def test_rgb888_scenes_exercise_their_feature(oracle):
    by = {s.name: s for s in RGB888}
    base = oracle.render_scene888(by["rgb888_opaque_painter"])[0]
    # RGB888 without dithering keeps all 8 bits; with dithering every written channel is a multiple of 8
    nd = oracle.render_scene888(by["rgb888_opaque_nodither"])[0]
    clear = np.array(list(by["rgb888_opaque_painter"].clear) + [255], np.uint8)
    drawn = (base != clear).any(-1)
    assert (base[drawn][:, :3] % 8 == 0).all() and (nd[(nd != clear).any(-1)][:, :3] % 8 != 0).any()
    for name in ("rgb888_gouraud_lights", "rgb888_flat_lights", "rgb888_mixed_painter",
                 "rgb888_editor_alpha_zbuffer", "rgb888_untextured_vertex_colours"):
        assert not np.array_equal(oracle.render_scene888(by[name])[0], base), name
    # per-texel blend tags really blend: the mixed scene differs from itself with every tag forced to Opaque
    m = by["rgb888_mixed_painter"]
    forced = []
    for t in m.textures8:
        px = np.array(t.pixels, dtype=np.uint8).reshape(-1, 4).copy()
        px[(px[:, 3] != abi.BLEND_ERASE), 3] = abi.BLEND_OPAQUE
        forced.append(type(t)(t.width, t.height, px.reshape(-1)))
    import copy as _copy
    m2 = _copy.copy(m); m2.textures8 = forced
    assert not np.array_equal(oracle.render_scene888(m2)[0], oracle.render_scene888(m)[0])


def test_rgb888_nan_key_panics_only_in_painters_mode(oracle):
    sc = next(s for s in RGB888 if s.name == "rgb888_opaque_painter")
    import dataclasses
    for fi in range(len(sc.faces)):
        v = sc.vertices.copy()
        v["pos"][sc.faces["v"][fi, 0], 2] = np.nan
        bad = dataclasses.replace(sc, vertices=v)
        if oracle.render_scene888(bad)[3] == abi.B32_ERR_NAN_DEPTH:
            zb = dataclasses.replace(bad, settings=dataclasses.replace(bad.settings, use_zbuffer=True))
            assert oracle.render_scene888(zb)[3] == 0          # no sort in z-buffer mode (render.rs:2155)
            return
    raise AssertionError("no face produces a NaN sort key")


# ---- skybox sphere pass (Framebuffer::render_skybox step 1, render.rs:81-139, :242-299) -------------------
SKY = cases.sky_cases()
with open(os.path.join(GOLDEN, "hashes_sky.json")) as _f:
    HASHES_SKY = json.load(_f)


@pytest.mark.parametrize("name,w,h,cam", SKY, ids=[c[0] for c in SKY])
def test_skybox_oracle_equals_numpy_model_and_golden(oracle, name, w, h, cam):
    sv, f = cases.sky_mesh(cam.position)
    want, _ = pymodel.fb_clear(w, h, (0, 0, 0))
    assert oracle.render_skybox_mesh(want, sv, f, cam) == 0
    rgba, _ = pymodel.fb_clear(w, h, (0, 0, 0))
    pymodel.render_skybox_mesh(rgba, sv, f, cam)
    assert np.array_equal(rgba, want)
    assert hashlib.sha256(want.tobytes()).hexdigest() == HASHES_SKY[name]
    covered = (want[..., :3] != 0).any(-1).mean()
    assert covered > 0.95, covered                 # inside the sphere the whole screen is sky


# ---- fuzz: random settings x adversarial geometry; both restatements must agree (incl. on where the reference panics) ----
import fuzz


@pytest.mark.parametrize("rgb888", [False, True], ids=["rgb555", "rgb888"])
def test_fuzz_oracle_equals_numpy_model(oracle, rgb888):
    ok = panics = 0
    for seed in range(24):
        sc = fuzz.fuzz_scene(seed, rgb888)
        want, want_z, tm, rc, order = (oracle.render_scene888 if rgb888 else oracle.render_scene)(sc, want_order=True)
        rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
        try:
            with np.errstate(all="ignore"):
                if rgb888:
                    order2 = pymodel.render_mesh(rgba, z, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
                else:
                    order2 = pymodel.render_mesh_15(rgba, z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        except pymodel.ReferencePanic:
            assert rc in (abi.B32_ERR_NAN_DEPTH, abi.B32_ERR_OOB_INDEX), sc.name
            panics += 1
            continue
        assert rc == 0, sc.name
        assert list(order) == order2, sc.name
        assert np.array_equal(rgba, want), sc.name
        assert np.array_equal(z.view(np.uint32), want_z.view(np.uint32)), sc.name
        ok += 1
    assert ok >= 12


def test_math_rs_reference_facts():                     # math.rs:783-797 (the reference's own unit tests of the vector helpers)
    from bonnie32_b200.raster import _cross
    a = np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    assert abs(float(pymodel.dot3(a, np.array([4.0, 5.0, 6.0], dtype=np.float32))[0]) - 32.0) < 0.001
    c = _cross(np.array([1.0, 0.0, 0.0], np.float32), np.array([0.0, 1.0, 0.0], np.float32))
    assert abs(float(c[2]) - 1.0) < 0.001 and c[0] == 0 and c[1] == 0


def _screen_to_ray(sx, sy, w, h, cam):
    """ray.rs:46-100, restated for the reference's own projection round-trip test."""
    vs = (min(w, h) / 2.0) * 0.75
    us = 5.0 - 1.0
    ndc_x, ndc_y = (sx - w / 2.0) / vs, (sy - h / 2.0) / vs
    d = np.array([ndc_x / us, ndc_y / us, 1.0])
    world = cam.basis_x.astype(np.float64) * d[0] + cam.basis_y.astype(np.float64) * d[1] + cam.basis_z.astype(np.float64) * d[2]
    world /= np.linalg.norm(world)
    o_cam = np.array([ndc_x * 5.0 / us, ndc_y * 5.0 / us, 0.0])
    origin = cam.position.astype(np.float64) + cam.basis_x * o_cam[0] + cam.basis_y * o_cam[1] + cam.basis_z * o_cam[2]
    return origin, world


def ray_roundtrip_distance(project, w=320, h=240):
    """The reference's test_screen_to_ray_roundtrip (ray.rs:333-377): a world point projected with the rasterizer's float
    projection and cast back as a ray passes within 2 units of the point.  `project(vertices, camera, settings, w, h)`
    returns screen positions [n,3]."""
    from bonnie32_b200.raster import Camera
    cam = Camera()
    cam.position = np.array([0.0, 0.0, -100.0], np.float32)
    cam.update_basis()
    world_point = np.array([50.0, 30.0, 100.0])
    v = scenes.make_vertices([tuple(world_point)])
    scr = project(v, cam, scenes.common_settings(use_fixed_point=False), w, h)
    origin, direction = _screen_to_ray(float(scr[0, 0]), float(scr[0, 1]), w, h, cam)
    t = np.dot(world_point - origin, direction)
    return float(np.linalg.norm(origin + t * direction - world_point)), scr[0]


def test_ray_rs_projection_roundtrip(oracle):          # ray.rs:333-377 (tolerance of the reference's own test: 2 units)
    dist, scr = ray_roundtrip_distance(lambda v, c, s, w, h: oracle.transform(v, c, s, w, h)[0])
    assert dist < 2.0, (dist, scr)


# ---- Framebuffer::clear_gradient and the overlay line family: the two restatements agree -------------------
LINES = cases.line_cases()


@pytest.mark.parametrize("name,w,h,seed,lines", [c for c in LINES if c[1] * c[2] <= 320 * 240 and "far" not in c[0]], ids=lambda v: v if isinstance(v, str) else None)
def test_lines_oracle_matches_numpy_model(oracle, name, w, h, seed, lines):
    from oracle import pymodel
    rgba, z = cases.line_background(w, h, seed)
    a = rgba.copy(); b = rgba.copy()
    assert oracle.draw_lines(a, z, lines) == 0
    pymodel.draw_lines(b, z, lines)
    assert np.array_equal(a, b), f"{name}: {(a != b).any(-1).sum()} pixels differ"
    assert not np.array_equal(a, rgba)


def test_line_walk_known_points(oracle):
    """Hand-checked walks of the reference loop (render.rs:719-751): end points included, a point is one pixel, and a
    shallow line steps its minor axis at the half-way error."""
    from bonnie32_b200 import abi, raster
    def drawn(x0, y0, x1, y1):
        rgba = np.zeros((8, 8, 4), np.uint8); z = np.zeros((8, 8), np.float32)
        oracle.draw_lines(rgba, z, raster.make_lines([raster.line_entry(abi.LINE_2D, x0, y0, x1, y1, (255, 255, 255))]))
        ys, xs = np.nonzero(rgba[..., 0])
        return sorted(zip(xs.tolist(), ys.tolist()))
    assert drawn(2, 3, 2, 3) == [(2, 3)]
    assert drawn(0, 0, 3, 3) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert drawn(5, 1, 1, 1) == [(1, 1), (2, 1), (3, 1), (4, 1), (5, 1)]
    # dx = 4, dy = -1: err 3 -> (e2 = 6: x) 2 -> (e2 = 4: x, and 4 <= dx: y) 5 -> 4 -> 3
    assert drawn(0, 0, 4, 1) == [(0, 0), (1, 0), (2, 1), (3, 1), (4, 1)]
    assert drawn(-2, 0, 1, 0) == [(0, 0), (1, 0)]                            # off-screen part is walked, not drawn


@pytest.mark.parametrize("w,h,top,bottom", [(320, 240, (10, 20, 200), (250, 128, 0)), (5, 1, (9, 8, 7), (200, 100, 50)),
                                            (3, 2, (0, 0, 0), (255, 255, 255)), (16, 97, (255, 0, 31, 5), (0, 255, 32))])
def test_clear_gradient_oracle_matches_numpy_model(oracle, w, h, top, bottom):
    from oracle import pymodel
    rgba = np.full((h, w, 4), 7, np.uint8); z = np.zeros((h, w), np.float32)
    oracle.fb_clear_gradient(rgba, z, top, bottom)
    want, want_z = pymodel.fb_clear_gradient(w, h, top, bottom)
    assert np.array_equal(rgba, want) and np.array_equal(z, want_z)
    assert tuple(rgba[0, 0, :3]) == tuple(top[:3]) and (h == 1 or tuple(rgba[-1, 0, :3]) == tuple(bottom[:3]))


@pytest.mark.parametrize("size", [0.5, 2.0, 3.7])
def test_stars_oracle_matches_numpy_model(oracle, size):
    from oracle import pymodel
    name, w, h, cam = cases.sky_cases()[1]
    stars = cases.star_list(cam, w, h, time=1.25)
    a = np.zeros((h, w, 4), np.uint8); b = a.copy()
    assert oracle.render_stars(a, stars, cam, size) == 0
    pymodel.render_stars(b, stars, cam, size)
    assert np.array_equal(a, b)
    lit = int((a[..., 3] == 255).sum())
    assert lit > 5 and (size < 2 or lit > 40)
    # the list builder is deterministic; directions are unit vectors; twinkle dims some stars
    again = cases.star_list(cam, w, h, time=1.25)
    assert stars.tobytes() == again.tobytes()
    assert np.allclose(np.linalg.norm(stars["dir"], axis=1), 1.0, atol=1e-5)
    assert len(np.unique(stars["rgb"], axis=0)) > 10


@pytest.mark.parametrize("facing,pos", [(0.0, (0, 0, 0)), (0.00005, (0.0001, 0, 0)), (1.1, (0, 0, 0)), (-2.7, (3.5, -1.25, 40.0)), (0.0, (0, 0.5, 0))])
def test_place_vertices_oracle_matches_numpy_model(oracle, facing, pos):
    from oracle import pymodel
    from bonnie32_b200 import raster
    sc = scenes.scene_c2(n_tris=50)
    c, s = raster.libm_cosf(facing), raster.libm_sinf(facing)
    a = oracle.place_vertices(sc.vertices, facing, c, s, pos)
    b = pymodel.place_vertices(sc.vertices, facing, c, s, pos)
    assert a.tobytes() == b.tobytes()
    moved = a.tobytes() != np.ascontiguousarray(sc.vertices).tobytes()
    assert moved == (abs(facing) > 0.0001 or any(abs(x) > 0.0001 for x in pos))


# ---- wireframe phase (render.rs:2574-2635): both restatements, every wireframe scene ---------------------------
WIRE = cases.wireframe_scenes()


@pytest.mark.parametrize("sc", WIRE, ids=[s.name for s in WIRE])
def test_wireframe_phase_oracle_equals_numpy_model(oracle, sc):
    want, want_z, tm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0
    rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
    order2 = pymodel.render_mesh_15(rgba, z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    assert list(order) == order2
    bad = (rgba != want).any(-1)
    assert not bad.any(), f"{sc.name}: {bad.sum()} pixels differ"
    assert np.array_equal(z.view(np.uint32), want_z.view(np.uint32))


# ---- filled primitives of the overlay family (draw_circle, draw_circle_alpha, draw_filled_rect, draw_thick_line, draw_rect) ----
PRIMS = cases.prim_cases()


@pytest.mark.parametrize("name,w,h,seed,lines", PRIMS[:3], ids=lambda v: v if isinstance(v, str) else None)
def test_prims_oracle_matches_numpy_model(oracle, name, w, h, seed, lines):
    rgba, z = cases.line_background(w, h, seed)
    a = rgba.copy(); b = rgba.copy()
    assert oracle.draw_lines(a, z, lines) == 0
    pymodel.draw_lines(b, z, lines)
    bad = (a != b).any(-1)
    assert not bad.any(), f"{name}: {bad.sum()} pixels differ, first at {np.argwhere(bad)[0][::-1]}"
    assert len({int(k) for k in lines["kind"]}) == 9


def test_prims_known_shapes(oracle):
    """Hand-checked shapes: a radius-1 circle is a plus, radius 0 one pixel, a negative radius nothing; a filled rectangle is
    inclusive of both corners; draw_rect is its outline; a thick horizontal line of thickness 2 covers the two rows whose
    centres lie inside the quad."""
    from bonnie32_b200 import raster
    def drawn(entries):
        rgba = np.zeros((12, 12, 4), np.uint8); z = np.zeros((12, 12), np.float32)
        assert oracle.draw_lines(rgba, z, raster.make_lines(entries)) == 0
        ys, xs = np.nonzero(rgba[..., 0])
        return sorted(zip(xs.tolist(), ys.tolist()))
    c = (255, 255, 255)
    assert drawn([raster.line_entry(abi.LINE_CIRCLE, 5, 5, 1, 0, c)]) == [(4, 5), (5, 4), (5, 5), (5, 6), (6, 5)]
    assert drawn([raster.line_entry(abi.LINE_CIRCLE, 5, 5, 0, 0, c)]) == [(5, 5)]
    assert drawn([raster.line_entry(abi.LINE_CIRCLE, 5, 5, -1, 0, c)]) == []
    assert drawn([raster.line_entry(abi.LINE_FILLED_RECT, 3, 2, 1, 3, c)]) == [(1, 2), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3)]
    outline = drawn(raster.rect_entries(2, 2, 5, 4, c))
    assert outline == sorted({(x, y) for x in range(2, 6) for y in range(2, 5)} - {(3, 3), (4, 3)})
    # thickness 2 around y = 5: the quad spans y in [4, 6]; pixel centres y + 0.5 inside for rows 4 and 5; x from 2 to 7
    assert drawn([raster.line_entry(abi.LINE_THICK, 2, 5, 8, 5, c, z0=2.0)]) == sorted((x, y) for x in range(2, 8) for y in (4, 5))
    assert drawn([raster.line_entry(abi.LINE_THICK, 2, 5, 8, 5, c, z0=1.0)]) == [(x, 5) for x in range(2, 9)]      # thickness 1 = draw_line
    assert drawn([raster.line_entry(abi.LINE_THICK, 4, 4, 4, 4, c, z0=5.0)]) == []                                   # zero length
