"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import copy
import dataclasses

import numpy as np
import pytest

import bonnie32_b200 as pkg
from bonnie32_b200 import abi, scenes
import cases

pytestmark = pytest.mark.gpu


def render_gpu(ctx, sc, resident=False):
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    if resident:
        ctx.set_textures(sc.textures)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        tm = mesh.render(sc.camera, sc.settings, sc.fog)
        mesh.free()
    else:
        tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    rgba, z = fb.download()
    return rgba, z, tm


def assert_same(sc, got, got_z, tm, want, want_z, otm):
    assert tm["triangles_drawn"] == otm["triangles_drawn"], sc.name
    if not np.array_equal(got, want):
        bad = (got != want).any(axis=-1)
        ys, xs = np.nonzero(bad)
        raise AssertionError(f"{sc.name}: {bad.sum()} pixels differ, first at (x={xs[0]}, y={ys[0]}): "
                             f"got {got[ys[0], xs[0]]} want {want[ys[0], xs[0]]}")
    # bit for bit — except that a NaN depth (render_mesh's editor-alpha writer stores one: `z >= zbuffer` is false for NaN,
    # render.rs:393) may differ in sign / payload: IEEE leaves those to the implementation (x86 produces 0xFFC00000, the
    # GPU 0x7FFFFFFF, wasm either), and no comparison can tell them apart
    zb = (got_z.view(np.uint32) != want_z.view(np.uint32)) & ~(np.isnan(got_z) & np.isnan(want_z))
    assert not zb.any(), f"{sc.name}: {zb.sum()} z-buffer values differ"


FEATURES = cases.feature_scenes()


@pytest.mark.parametrize("sc", FEATURES, ids=[s.name for s in FEATURES])
def test_feature_scene(ctx, oracle, sc):
    want, want_z, otm, rc, order = oracle.render_scene(sc, want_order=True)
    assert rc == 0
    got, got_z, tm = render_gpu(ctx, sc)
    # draw order (opaque pass then transparent pass) must be the reference's
    n = np.zeros(1, dtype=np.uint32)
    buf = np.zeros(max(len(sc.faces), 1), dtype=np.uint32)
    import ctypes as C
    ctx.check(ctx.lib.b32_debug_draw_order(ctx.h, buf.ctypes.data, len(buf), n.ctypes.data_as(C.POINTER(C.c_uint32))))
    assert n[0] == len(order)
    assert np.array_equal(buf[: n[0]], order), sc.name
    assert_same(sc, got, got_z, tm, want, want_z, otm)


@pytest.mark.parametrize("fixed", [True, False])
def test_c1_single_triangle(ctx, oracle, fixed):
    sc = scenes.scene_c1(fixed)
    want, want_z, otm, rc = oracle.render_scene(sc)
    got, got_z, tm = render_gpu(ctx, sc)
    assert tm["triangles_drawn"] == 1          # the reversed copy is culled
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_big_triangles_slow_edge_path(ctx, oracle):
    sc = cases.big_triangle_scene()
    want, want_z, otm, rc = oracle.render_scene(sc)
    got, got_z, tm = render_gpu(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_stepped_surfaces_switch_the_fill_to_the_shared_edge_prefix(ctx, oracle, size):
    """Fixed-point calls whose large surfaces leave the exact-integer range (far off-screen vertices): k_setup tells the
    host through a mapped word, and the next calls on the context — blocking, resident or enqueue-only — run
    k_fill_opaque<.., PRE> (shared edge prefix + conservative box reject) until 8 calls have gone by without such a
    surface.  Every frame equals the oracle, whichever instantiation drew it."""
    big = cases._with(cases.big_triangle_scene(), "big_triangles", width=size[0], height=size[1])
    plain = cases._with(scenes.scene_c2(n_tris=300), "plain", width=size[0], height=size[1])
    want_big = oracle.render_scene(big)
    want_plain = oracle.render_scene(plain)
    assert want_big[3] == 0 and want_plain[3] == 0
    hint = lambda: ctx.lib.b32_debug_prefix_hint(ctx.h)
    for _ in range(9):                                      # whatever earlier tests left behind has aged out after 8 calls
        got, got_z, tm = render_gpu(ctx, plain)
    assert_same(plain, got, got_z, tm, *want_plain[:3])
    assert hint() == 0
    for k in range(3):                                      # call 0 per-pixel replay, calls 1.. the shared prefix
        got, got_z, tm = render_gpu(ctx, big, resident=(k == 2))
        assert_same(big, got, got_z, tm, *want_big[:3])
        assert hint() == 1
    fb = pkg.Framebuffer(big.width, big.height, ctx)
    ctx.set_textures(big.textures)
    mesh = pkg.Mesh(ctx, big.vertices, big.faces)
    for _ in range(12):                                     # enqueue-only frames keep the hint alive by themselves
        mesh.frame_enqueue(big.clear, big.camera, big.settings, big.fog)
    got, got_z = fb.download()
    mesh.free()
    assert np.array_equal(got, want_big[0]) and np.array_equal(got_z.view(np.uint32), want_big[1].view(np.uint32))
    assert hint() == 1
    for k in range(9):                                      # ordinary scenes: drawn by the PRE instantiation while the hint lasts, ...
        got, got_z, tm = render_gpu(ctx, plain)
        assert_same(plain, got, got_z, tm, *want_plain[:3])
        assert hint() == (1 if k < 7 else 0), k             # ... which is 8 calls
    got, got_z, tm = render_gpu(ctx, plain)
    assert_same(plain, got, got_z, tm, *want_plain[:3])


@pytest.mark.parametrize("zbuf", [False, True])
def test_c2_full(ctx, oracle, zbuf):
    sc = scenes.scene_c2(use_zbuffer=zbuf)
    want, want_z, otm, rc = oracle.render_scene(sc)
    got, got_z, tm = render_gpu(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


@pytest.mark.parametrize("zbuf", [False, True])
def test_c4_full_size(ctx, oracle, zbuf):
    """BASELINE config 4 at full size (100k triangles): byte-identical to the oracle."""
    sc = scenes.scene_c4(use_zbuffer=zbuf)
    want, want_z, otm, rc = oracle.render_scene(sc)
    got, got_z, tm = render_gpu(ctx, sc, resident=True)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_indexed_equals_expanded(ctx):
    """Sampling index->CLUT on the device == expanding to Texture15 on the host first."""
    sc = scenes.scene_c4(n_tris=5000)
    a, az, _ = render_gpu(ctx, sc)
    sc2 = copy.copy(sc)
    sc2.textures = [scenes.expand_texture(t) for t in sc.textures]
    b, bz, _ = render_gpu(ctx, sc2)
    assert np.array_equal(a, b) and np.array_equal(az.view(np.uint32), bz.view(np.uint32))


def test_multiple_calls_compose(ctx, oracle):
    """Several render_mesh_15 calls into one framebuffer, in call order (scene.rs:196-259)."""
    parts = [scenes.scene_c2(n_tris=200, seed=s, use_zbuffer=True) for s in (1, 2, 3)]
    fb = pkg.Framebuffer(320, 240, ctx)
    fb.clear(parts[0].clear)
    want = np.empty((240, 320, 4), np.uint8); want_z = np.empty((240, 320), np.float32)
    want[...] = np.array(list(parts[0].clear) + [255], np.uint8); want_z[...] = np.finfo(np.float32).max
    for sc in parts:
        pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        rc, _, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        assert rc == 0
    got, got_z = fb.download()
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


def test_fb_upload_roundtrip(ctx):
    fb = pkg.Framebuffer(64, 48, ctx)
    rng = np.random.default_rng(3)
    px = rng.integers(0, 256, (48, 64, 4), dtype=np.uint8)
    z = rng.random((48, 64), dtype=np.float32)
    fb.upload(px, z)
    a, b = fb.download()
    assert np.array_equal(a, px) and np.array_equal(b, z)
    fb.clear((1, 2, 3))
    a, b = fb.download()
    assert (a == np.array([1, 2, 3, 255], np.uint8)).all() and (b == np.finfo(np.float32).max).all()


def test_transform_parity_random_and_adversarial(ctx, oracle):
    """k_transform vs the oracle on 1M vertices incl. denormals, signed zeros, huge values, NaN/inf,
    and denominators near the |denom|<256 early-out (fixed.rs:406-408)."""
    import ctypes as C
    n = 1_000_000
    u = scenes.splitmix64_u01(42, n * 3).reshape(n, 3)
    pos = ((u - 0.5) * np.array([200.0, 200.0, 120.0])).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-40, -1e-40, 1e5, -1e5, 524287.9, -524288.0, 1e9, -1e9, 3.4e38, np.inf, -np.inf, np.nan,
                        -5.0, -5.05, -4.95, -4.9375, -5.0625, 0.1, 0.0999], dtype=np.float32)
    k = len(special)
    grid = np.stack(np.meshgrid(special, special, special, indexing="ij"), axis=-1).reshape(-1, 3)
    pos[: len(grid)] = grid
    v = scenes.make_vertices(pos)
    cam = cases._rotated_camera(0.37, -1.2, (3.0, -2.0, 1.5))
    for kw in (dict(), dict(use_fixed_point=False), dict(ortho_projection=(7.5, 1.0, -2.0))):
        s = scenes.common_settings(**kw)
        for w, h in ((320, 240), (640, 480), (201, 333)):
            fb = pkg.Framebuffer(w, h, ctx)
            want_s, want_c = oracle.transform(v, cam, s, w, h)
            got_s = np.empty((n, 3), np.float32); got_c = np.empty((n, 3), np.float32)
            ca = cam.to_abi(); sa, keep = s.to_abi()
            ctx.check(ctx.lib.b32_debug_transform(ctx.h, v.ctypes.data, n, C.byref(ca), C.byref(sa), got_s.ctypes.data, got_c.ctypes.data))
            # compare bit patterns; NaN payloads may differ, so canonicalise NaNs
            def canon(a):
                b = a.view(np.uint32).copy(); b[np.isnan(a)] = 0x7FC00000; return b
            assert np.array_equal(canon(got_s), canon(want_s)), (kw, w, h)
            assert np.array_equal(canon(got_c), canon(want_c)), (kw, w, h)


def test_error_codes(ctx, oracle):
    sc = scenes.scene_c2(n_tris=50)
    fb = pkg.Framebuffer(320, 240, ctx)
    fb.clear(sc.clear)
    before, _ = fb.download()
    # out-of-range vertex index: reference panics -> B32_ERR_OOB_INDEX, framebuffer untouched
    f = sc.faces.copy(); f["v"][7, 1] = len(sc.vertices)
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh_15(fb, sc.vertices, f, sc.textures, sc.camera, sc.settings)
    assert e.value.code == abi.B32_ERR_OOB_INDEX
    assert np.array_equal(fb.download()[0], before)
    # a face whose blend bits are not a BlendMode (types.rs: 0..=5): B32_ERR_INVALID, framebuffer untouched
    f = sc.faces.copy(); f["flags"][3] = abi.face_flags(0, 6, False, 255)
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh_15(fb, sc.vertices, f, sc.textures, sc.camera, sc.settings)
    assert e.value.code == abi.B32_ERR_INVALID
    assert np.array_equal(fb.download()[0], before)
    # NaN depth in a sorted pass: reference unwrap() panics -> B32_ERR_NAN_DEPTH
    v = cases.nan_depth_vertices(sc, oracle)
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh_15(fb, v, sc.faces, sc.textures, sc.camera, sc.settings)
    assert e.value.code == abi.B32_ERR_NAN_DEPTH
    assert np.array_equal(fb.download()[0], before)
    rc = oracle.render_scene(dataclasses.replace(sc, vertices=v))[3]
    assert rc == abi.B32_ERR_NAN_DEPTH
    # with the z-buffer on, the opaque pass is not sorted: NaN is legal and must match the oracle
    s2 = copy.copy(sc); s2.vertices = v; s2.settings = scenes.common_settings(use_zbuffer=True)
    want, want_z, otm, rc = oracle.render_scene(s2)
    assert rc == 0
    got, got_z, tm = render_gpu(ctx, s2)
    assert_same(s2, got, got_z, tm, want, want_z, otm)
    # a light type beyond Spot is not a LightType
    bad = pkg.Light.spot((0, 0, 0), (0, 0, 1), 0.5, 10.0, 1.0); bad.type = 3
    s3 = copy.copy(sc); s3.settings = scenes.common_settings(shading=abi.SHADE_GOURAUD, lights=[bad])
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, s3.settings)
    assert e.value.code == abi.B32_ERR_INVALID


def test_empty_inputs(ctx):
    fb = pkg.Framebuffer(320, 240, ctx)
    fb.clear((9, 8, 7))
    s = scenes.common_settings()
    v0 = np.zeros(0, abi.VERTEX_DTYPE); f0 = np.zeros(0, abi.FACE_DTYPE)
    tm = pkg.render_mesh_15(fb, v0, f0, [], pkg.Camera(), s)
    assert tm["triangles_drawn"] == 0
    sc = scenes.scene_c1()
    tm = pkg.render_mesh_15(fb, sc.vertices, f0, [], pkg.Camera(), s)
    assert tm["triangles_drawn"] == 0
    assert (fb.download()[0] == np.array([9, 8, 7, 255], np.uint8)).all()


def test_properties_full_size(ctx):
    """Size-independent properties at BASELINE size (100k triangles)."""
    sc = scenes.scene_c4()
    a, az, tm = render_gpu(ctx, sc, resident=True)
    # determinism: same inputs, same bytes
    b, bz, _ = render_gpu(ctx, sc, resident=True)
    assert np.array_equal(a, b) and np.array_equal(az.view(np.uint32), bz.view(np.uint32))
    # painter's mode never touches the z-buffer
    assert (az == np.finfo(np.float32).max).all()
    # every written pixel is 5-bit expanded with A=255: (v<<3)|(v>>2)
    written = (a[..., :3] != np.array(sc.clear, np.uint8)).any(-1)
    c = a[written][:, :3].astype(np.int32)
    assert ((((c >> 3) << 3) | (c >> 5)) == c).all() and (a[written][:, 3] == 255).all()
    # painter's order == reversing the face list reverses ties only: drawing the scene twice is idempotent
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.upload(a, az)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    assert np.array_equal(fb.download()[0], a)


def test_enqueue_only_path(ctx, oracle):
    """b32_render_mesh_15_enqueue: pass 1 without any host round trip; same bytes as the oracle."""
    for zbuf in (False, True):
        sc = scenes.scene_c4(n_tris=20000, use_zbuffer=zbuf)
        want, want_z, otm, rc = oracle.render_scene(sc)
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        ctx.set_textures(sc.textures)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        for _ in range(3):                       # several frames in flight
            fb.clear(sc.clear)
            mesh.render(sc.camera, sc.settings, sc.fog, enqueue_only=True)
        got, got_z = fb.download()
        mesh.free()
        assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


def test_enqueue_only_errors_surface_at_sync(ctx):
    sc = scenes.scene_c2(n_tris=50)
    f = sc.faces.copy(); f["v"][7, 1] = len(sc.vertices)
    fb = pkg.Framebuffer(320, 240, ctx); fb.clear(sc.clear)
    before = fb.download()[0]
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, f)
    mesh.render(sc.camera, sc.settings, None, enqueue_only=True)
    with pytest.raises(pkg.B32Error) as e:
        ctx.sync()
    assert e.value.code == abi.B32_ERR_OOB_INDEX
    assert np.array_equal(fb.download()[0], before)
    mesh.free()
    # x-ray mode has no pass-1 kernel: the ordered replay reports the reference's panic (found by tests/checks/fuzz_compact.py)
    import cases as _cases
    from oracle import oracle as _orc
    xs = scenes.common_settings(xray_mode=True)
    vnan = _cases.nan_depth_vertices(dataclasses.replace(sc, settings=xs), _orc)
    mesh = pkg.Mesh(ctx, vnan, sc.faces)
    mesh.render(sc.camera, xs, None, enqueue_only=True)
    with pytest.raises(pkg.B32Error) as e:
        ctx.sync()
    assert e.value.code == abi.B32_ERR_NAN_DEPTH
    assert np.array_equal(fb.download()[0], before)
    mesh.free()
    f = sc.faces.copy(); f["flags"][3] = abi.face_flags(0, 7, False, 255)        # not a BlendMode
    mesh = pkg.Mesh(ctx, sc.vertices, f)
    mesh.render(sc.camera, sc.settings, None, enqueue_only=True)
    with pytest.raises(pkg.B32Error) as e:
        ctx.sync()
    assert e.value.code == abi.B32_ERR_INVALID
    assert np.array_equal(fb.download()[0], before)
    mesh.free()


def test_crowded_tile(ctx, oracle):
    """Thousands of surfaces in one screen tile: pass 1 takes them in several windows of 1 024 (no bins, no capacity);
    the ordered pass sorts them in a slice of its global scratch, sized from k_setup's count before the pass runs."""
    n = 6000
    u = scenes.splitmix64_u01(5150, n * 9).reshape(n, 3, 3)
    pos = np.empty((n, 3, 3))
    pos[..., 0] = (u[..., 0] - 0.5) * 0.9          # all inside a few pixels around the screen centre
    pos[..., 1] = (u[..., 1] - 0.5) * 0.9
    pos[..., 2] = 10.0 + 30.0 * u[..., 2]
    v = scenes.make_vertices(pos.reshape(-1, 3), uv=u[..., :2].reshape(-1, 2),
                             rgba=np.concatenate([np.floor(u * 255).reshape(-1, 3), np.zeros((n * 3, 1))], axis=1))
    f = scenes.make_faces(np.arange(n * 3).reshape(n, 3), tex_id=abi.FACE_TEX_NONE)
    # second variant: every other face semi-transparent (pass 2): > 2048 entries in one tile exercises the
    # in-place global-memory tile sort and the sizing of the ordered pass's scratch
    f2 = f.copy()
    f2["flags"][::2] = abi.face_flags(abi.FACE_TEX_NONE, abi.BLEND_AVERAGE, True, 200)
    f2["flags"][1::4] = abi.face_flags(abi.FACE_TEX_NONE, abi.BLEND_ADD, False, 255)
    # third variant: every triangle at the same depth = one walk key for thousands of surfaces (pass 1's depth-ordered
    # windows meet a single over-full key bucket and take it in face order)
    v3 = v.copy()
    v3["pos"][:, 2] = np.float32(20.0)
    # fourth variant: thousands of candidates in the tile but only a few hundred of them semi-transparent: the ordered pass
    # meets a crowded mask row whose ordered entries fit shared memory, and no scratch exists (found by tests/checks/fuzz_big.py)
    f4 = f.copy()
    f4["flags"][::10] = abi.face_flags(abi.FACE_TEX_NONE, abi.BLEND_AVERAGE, True, 255)
    # last variant: a 1920x1080 framebuffer = coarse mask tiles (a fill tile's candidates include its neighbours' surfaces)
    for faces_, xray, verts_, size in ((f, False, v, (320, 240)), (f2, False, v, (320, 240)), (f2, True, v, (320, 240)),
                                       (f, False, v3, (320, 240)), (f4, False, v, (320, 240)), (f2, False, v, (1920, 1080))):
        for zbuf in (False, True):
            sc = scenes.Scene("crowded_tile", verts_, faces_, [], pkg.Camera(),
                              scenes.common_settings(use_zbuffer=zbuf, backface_cull=False, xray_mode=xray), width=size[0], height=size[1])
            want, want_z, otm, rc = oracle.render_scene(sc)
            ctx2 = pkg.Context(0)                     # fresh context: no scratch allocated yet
            got, got_z, tm = render_gpu(ctx2, sc)
            ctx2.close()
            assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_async_host_buffer_path(ctx, oracle):
    """b32_render_mesh_15_ex(ASYNC|ALL_OPAQUE) + b32_fb_download_async from pinned memory == oracle;
    a wrong ALL_OPAQUE assertion is reported at the next sync."""
    import ctypes as C
    sc = scenes.scene_c4(n_tris=20000)
    want, want_z, otm, rc = oracle.render_scene(sc)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    ctx.set_textures(sc.textures)
    lib = ctx.lib
    nvb, nfb, npx = sc.vertices.nbytes, sc.faces.nbytes, sc.width * sc.height * 4
    hv, hf, hp = lib.b32_host_alloc(nvb), lib.b32_host_alloc(nfb), lib.b32_host_alloc(npx)
    C.memmove(hv, sc.vertices.ctypes.data, nvb); C.memmove(hf, sc.faces.ctypes.data, nfb)
    cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi()
    flags = abi.RENDER_ASYNC | abi.RENDER_ALL_OPAQUE
    for _ in range(2):
        fb.clear(sc.clear)
        ctx.check(lib.b32_render_mesh_15_ex(ctx.h, hv, len(sc.vertices), hf, len(sc.faces), C.byref(cam), C.byref(st), None, flags, None))
        ctx.check(lib.b32_fb_download_async(ctx.h, hp, None))
    ctx.sync()
    got = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(sc.height, sc.width, 4)).copy()
    assert np.array_equal(got, want)
    # wrong assertion: a face with a blend mode
    f = sc.faces.copy(); f["flags"][:] = abi.face_flags(0, abi.BLEND_ADD, True, 255)
    C.memmove(hf, f.ctypes.data, nfb)
    ctx.check(lib.b32_render_mesh_15_ex(ctx.h, hv, len(sc.vertices), hf, len(sc.faces), C.byref(cam), C.byref(st), None, flags, None))
    with pytest.raises(pkg.B32Error) as e:
        ctx.sync()
    assert e.value.code == abi.B32_ERR_INVALID
    # ASYNC without the promise: both passes are enqueued, and the blended faces are drawn
    sc2 = cases._with(sc, "async_blended")
    sc2.faces = f
    want2, want2_z, _, rc2 = oracle.render_scene(sc2)
    assert rc2 == 0
    fb.clear(sc.clear)
    ctx.check(lib.b32_render_mesh_15_ex(ctx.h, hv, len(sc.vertices), hf, len(sc.faces), C.byref(cam), C.byref(st), None, abi.RENDER_ASYNC, None))
    ctx.check(lib.b32_fb_download_async(ctx.h, hp, None))
    ctx.sync()
    got2 = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(sc.height, sc.width, 4)).copy()
    assert np.array_equal(got2, want2)
    for h in (hv, hf, hp):
        lib.b32_host_free(h)


WIRE = cases.wireframe_scenes()


@pytest.mark.parametrize("sc", WIRE, ids=[s.name for s in WIRE])
def test_wireframe_phase(ctx, oracle, sc):
    """render.rs:2574-2635 on the device vs the oracle (first-occurrence edge de-duplication, depth-tested
    Bresenham lines, overlay mode draws no solid surfaces)."""
    want, want_z, otm, rc = oracle.render_scene(sc)
    assert rc == 0
    got, got_z, tm = render_gpu(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_wireframe_scenes_draw_lines(oracle):
    by = {s.name: s for s in WIRE}
    plain = copy.copy(by["wire_backface_zbuffer"]); plain.settings = copy.copy(plain.settings); plain.settings.backface_wireframe = False
    a = oracle.render_scene(plain)[0]; b = oracle.render_scene(by["wire_backface_zbuffer"])[0]
    diff = (a != b).any(-1)
    assert diff.sum() > 100 and (b[diff][:, :3] == np.array([80, 80, 100], np.uint8)).all()
    c = oracle.render_scene(by["wire_overlay"])[0]
    assert ((c[..., :3] == np.array([200, 200, 220], np.uint8)).all(-1)).sum() > 100


# ---- RGB888 sibling: b32_render_mesh vs the oracle's render_mesh (render.rs:1971-2259) -------------------
RGB888 = cases.rgb888_scenes()


def render_gpu888(ctx, sc, resident=False):
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    if resident:
        ctx.set_textures_rgb888(sc.textures8)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        tm = mesh.render_rgb888(sc.camera, sc.settings)
        mesh.free()
    else:
        tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
    rgba, z = fb.download()
    return rgba, z, tm


@pytest.mark.parametrize("sc", RGB888, ids=[s.name for s in RGB888])
def test_rgb888_scene(ctx, oracle, sc):
    want, want_z, otm, rc, order = oracle.render_scene888(sc, want_order=True)
    assert rc == 0
    got, got_z, tm = render_gpu888(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


@pytest.mark.parametrize("zbuf,blended", [(False, False), (True, False), (True, True), (False, True)])
def test_rgb888_c2_full(ctx, oracle, zbuf, blended):
    """1 000 triangles through render_mesh: order-free pass (opaque texels) and ordered replay (blend tags)."""
    sc = scenes.scene_c2(use_zbuffer=zbuf)
    sc.settings.use_rgb555 = False
    sc.textures8 = [cases._rng_texture8(31, 64, 64, blend_fraction=0.3 if blended else 0.0)]
    want, want_z, otm, rc = oracle.render_scene888(sc)
    got, got_z, tm = render_gpu888(ctx, sc, resident=True)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_rgb888_and_rgb555_tables_are_independent(ctx, oracle):
    """Both texture tables stay bound: alternating render_mesh / render_mesh_15 calls (the editor's RGB555 toggle)."""
    a = next(s for s in RGB888 if s.name == "rgb888_opaque_zbuffer")
    b = next(s for s in FEATURES if s.name == "zbuffer_idx8")
    for _ in range(2):
        got, got_z, tm = render_gpu888(ctx, a)
        want, want_z, otm, rc = oracle.render_scene888(a)
        assert_same(a, got, got_z, tm, want, want_z, otm)
        got, got_z, tm = render_gpu(ctx, b)
        want, want_z, otm, rc = oracle.render_scene(b)
        assert_same(b, got, got_z, tm, want, want_z, otm)


def test_rgb888_error_codes(ctx, oracle):
    sc = next(s for s in RGB888 if s.name == "rgb888_opaque_painter")
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    before, _ = fb.download()
    f = sc.faces.copy(); f["v"][3, 1] = len(sc.vertices)
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh(fb, sc.vertices, f, sc.textures8, sc.camera, sc.settings)
    assert e.value.code == abi.B32_ERR_OOB_INDEX
    for fi in range(len(sc.faces)):
        v = sc.vertices.copy()
        v["pos"][sc.faces["v"][fi, 0], 2] = np.nan
        if oracle.render_scene888(dataclasses.replace(sc, vertices=v))[3] == abi.B32_ERR_NAN_DEPTH:
            break
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh(fb, v, sc.faces, sc.textures8, sc.camera, sc.settings)
    assert e.value.code == abi.B32_ERR_NAN_DEPTH
    assert np.array_equal(fb.download()[0], before)           # the reference panics before drawing
    zs = dataclasses.replace(sc.settings, use_zbuffer=True)   # no sort in z-buffer mode: the NaN face just draws nothing
    pkg.render_mesh(fb, v, sc.faces, sc.textures8, sc.camera, zs)
    want, want_z, otm, rc = oracle.render_scene888(dataclasses.replace(sc, vertices=v, settings=zs))
    assert rc == 0 and np.array_equal(fb.download()[0], want)


# ---- enqueued frames replayed as a CUDA graph with re-parameterised kernel nodes --------------------------
def test_frame_graph_replay_matches_oracle(ctx, oracle):
    """b32_frame_15_enqueue: frames 3.. of one mesh are graph launches; camera / settings / fog / lights / clear colour
    change every frame and every frame must still equal the oracle.  A second mesh and a resize rebuild the graph."""
    base = cases.feature_scenes(300)
    a = next(s for s in base if s.name == "gouraud_lights")
    b = scenes.scene_c2(n_tris=500, seed=77)
    fb = pkg.Framebuffer(a.width, a.height, ctx)
    ctx.set_textures(a.textures)
    g0 = ctx.graph_launches()
    variants = [dict(), dict(use_zbuffer=False), dict(dithering=False, shading=abi.SHADE_FLAT), dict(affine_textures=False),
                dict(use_fixed_point=False), dict(ambient=0.9), dict(backface_cull=False)]
    for sc, size in ((a, (320, 240)), (b, (320, 240)), (b, (200, 150)), (a, (320, 240))):
        fb.resize(*size)
        ctx.set_textures(sc.textures)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        for i, kw in enumerate(variants):
            st = dataclasses.replace(sc.settings, **kw)
            cam = cases._rotated_camera(0.02 * i, -0.03 * i, (0.1 * i, -0.05 * i, -0.2 * i))
            fog = (10.0, 30.0, 50.0 + i, (90, 110, 130)) if i % 3 == 2 else None
            clear = (20 + i, 22, 28 + 2 * i)
            mesh.frame_enqueue(clear, cam, st, fog)
            got, got_z = fb.download()
            want = np.empty((size[1], size[0], 4), np.uint8); want_z = np.empty((size[1], size[0]), np.float32)
            want[...] = np.array(list(clear) + [255], np.uint8); want_z[...] = np.finfo(np.float32).max
            rc, otm, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, cam, st, fog)
            assert rc == 0
            assert np.array_equal(got, want), (sc.name, size, i)
            assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (sc.name, size, i)
        mesh.free()
    assert ctx.graph_launches() - g0 >= 4 * (len(variants) - 1) - 4      # all but the first frame(s) of each topology


def test_frame_graph_reports_errors_at_sync(ctx):
    """A NaN sort key in an enqueued (graph) frame surfaces at the next sync, as for plain enqueues."""
    sc = scenes.scene_c2(n_tris=300)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    ctx.set_textures(sc.textures)
    v = sc.vertices.copy()
    mesh = pkg.Mesh(ctx, v, sc.faces)
    for _ in range(3):
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings)
    ctx.sync()
    f = sc.faces.copy(); f["v"][5, 2] = len(v) + 7
    bad = pkg.Mesh(ctx, v, f)
    for _ in range(4):
        bad.frame_enqueue(sc.clear, sc.camera, sc.settings)
    with pytest.raises(pkg.B32Error) as e:
        ctx.sync()
    assert e.value.code == abi.B32_ERR_OOB_INDEX
    mesh.free(); bad.free()


# ---- skybox sphere pass ---------------------------------------------------------------------------------------
SKY = cases.sky_cases()


@pytest.mark.parametrize("name,w,h,cam", SKY, ids=[c[0] for c in SKY])
def test_skybox_mesh(ctx, oracle, name, w, h, cam):
    sv, f = cases.sky_mesh(cam.position)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.clear((0, 0, 0))
    _, z0 = fb.download()
    fb.render_skybox_mesh(sv, f, cam)
    got, got_z = fb.download()
    want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
    assert oracle.render_skybox_mesh(want, sv, f, cam) == 0
    bad = (got != want).any(-1)
    assert not bad.any(), f"{name}: {bad.sum()} pixels differ"
    assert np.array_equal(got_z.view(np.uint32), z0.view(np.uint32))     # the sky never touches the z-buffer


def test_skybox_then_scene_composes(ctx, oracle):
    """Game frame order (src/game/renderer.rs:91-137): clear, skybox, then the rooms over it."""
    name, w, h, cam = SKY[1]
    sv, f = cases.sky_mesh(cam.position)
    sc = scenes.scene_c2(n_tris=400, use_zbuffer=True)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.clear((0, 0, 0))
    fb.render_skybox_mesh(sv, f, cam)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    got, got_z = fb.download()
    want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
    want_z = np.full((h, w), np.finfo(np.float32).max, np.float32)
    oracle.render_skybox_mesh(want, sv, f, cam)
    rc, _, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    assert rc == 0 and np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))
    with pytest.raises(pkg.B32Error) as e:
        fb.render_skybox_mesh(sv, np.array([[0, 1, len(sv)]], np.uint32), cam)
    assert e.value.code == abi.B32_ERR_OOB_INDEX


# ---- fuzz: random settings x adversarial geometry (indexed meshes, degenerate / huge / near-plane / non-finite triangles,
#      zero-sized and out-of-range textures, odd framebuffer sizes) ----------------------------------------------------
import fuzz


@pytest.mark.parametrize("rgb888", [False, True], ids=["rgb555", "rgb888"])
def test_fuzz_gpu_equals_oracle(ctx, oracle, rgb888):
    ok = panics = 0
    for seed in range(60):
        sc = fuzz.fuzz_scene(seed, rgb888)
        want, want_z, otm, rc = (oracle.render_scene888 if rgb888 else oracle.render_scene)(sc)
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        fb.clear(sc.clear)
        before = fb.download()[0]
        try:
            if rgb888:
                tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
            else:
                tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        except pkg.B32Error as e:
            assert e.code == rc, (sc.name, e.code, rc)
            assert np.array_equal(fb.download()[0], before), sc.name       # the reference panics before it draws
            panics += 1
            continue
        assert rc == 0, sc.name
        got, got_z = fb.download()
        assert_same(sc, got, got_z, tm, want, want_z, otm)
        ok += 1
    assert ok >= 30 and ok + panics == 60


@pytest.mark.parametrize("seed,n_tris", [(4249, 400), (1295, 1500), (4689, 1500), (5202, 1500), (2079, 120)])
def test_fuzz_regressions_nan_depth_in_render_mesh(ctx, oracle, seed, n_tris):
    """Found by tests/checks/fuzz_extended.py (10 000 scenes): render_mesh's editor-alpha writer stores a NaN depth
    (`z >= zbuffer` is false for NaN, render.rs:393) and the next such fragment replaces it with ANY finite depth, so a
    pixel's depth can rise — the ordered replay must not reject a fragment early against a depth taken before that."""
    sc = fuzz.fuzz_scene(seed, True, n_tris=n_tris)
    want, want_z, otm, rc = oracle.render_scene888(sc)
    assert rc == 0
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
    got, got_z = fb.download()
    assert_same(sc, got, got_z, tm, want, want_z, otm)


@pytest.mark.parametrize("seed", [115649, 109344])
def test_fuzz_regression_never_cull_keys_walk_first(ctx, oracle, seed):
    """Found by tests/checks/fuzz_extended.py (40 000 scenes): ortho + z-buffer frames, where some surfaces claim no depth
    bound (walk key 0xFFFFFFFF) and others do, differed from run to run in about every second run — a no-bound surface that
    shared the top key bucket with bounded ones could land behind them and be skipped by their bound.  Repeated, because
    the order inside a bucket comes from shared-memory atomics."""
    sc = fuzz.fuzz_scene(seed, False, n_tris=1500)
    want, want_z, otm, rc = oracle.render_scene(sc)
    assert rc == 0
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    for rep in range(10):
        fb.clear(sc.clear)
        tm = mesh.render(sc.camera, sc.settings, sc.fog)
        got, got_z = fb.download()
        assert_same(sc, got, got_z, tm, want, want_z, otm)
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
        got, got_z = fb.download()
        assert_same(sc, got, got_z, tm, want, want_z, otm)
    mesh.free()


def test_two_devices_in_one_process(oracle):
    """A host thread that holds contexts on two GPUs: every entry point selects its context's device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = scenes.scene_c2(n_tris=400, use_zbuffer=True)
    want, want_z, otm, rc = oracle.render_scene(sc)
    ctxs = [pkg.Context(0), pkg.Context(1)]
    fbs = [pkg.Framebuffer(sc.width, sc.height, c) for c in ctxs]
    for _ in range(2):                                   # interleaved calls on the two devices
        for c, fb in zip(ctxs, fbs):
            fb.clear(sc.clear)
            tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
        for c, fb in zip(ctxs, fbs):
            got, got_z = fb.download()
            assert_same(sc, got, got_z, tm, want, want_z, otm)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("w,h", [(640, 480), (1024, 768), (1920, 1080)])
def test_large_framebuffers_all_passes(ctx, oracle, w, h):
    """Both passes, the wireframe phase and the RGB888 replay at the editor / game framebuffer sizes (src/game/renderer.rs:34-49)."""
    by = {s.name: s for s in cases.feature_scenes(200)}
    for name in ("mixed_zbuffer", "mixed_painter", "gouraud_lights"):
        sc = cases._with(by[name], f"{name}_{w}x{h}", width=w, height=h, backface_wireframe=(name == "mixed_zbuffer"))
        want, want_z, otm, rc = oracle.render_scene(sc)
        assert rc == 0
        got, got_z, tm = render_gpu(ctx, sc)
        assert_same(sc, got, got_z, tm, want, want_z, otm)
        if not sc.settings.backface_wireframe:                 # the same frame enqueued (clear folded in, both passes, graph replay)
            fb = pkg.Framebuffer(w, h, ctx)
            ctx.set_textures(sc.textures)
            mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
            for _ in range(3):
                mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
            got, got_z = fb.download()
            mesh.free()
            assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (sc.name, "enqueued")
    m = next(s for s in RGB888 if s.name == "rgb888_mixed_zbuffer")
    sc = cases._with(m, f"rgb888_mixed_zbuffer_{w}x{h}", width=w, height=h)
    want, want_z, otm, rc = oracle.render_scene888(sc)
    got, got_z, tm = render_gpu888(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_enqueued_c4_at_1080p_stays_small(oracle):
    """An enqueued 100k-triangle frame at 1920x1080 (8 160 fill tiles): round 1's worst-case bins needed 26 GB per context
    here; the tile masks need a few MB.  Device memory taken by a fresh context for this frame stays below 200 MB (the
    framebuffer, the resident mesh and the work buffers together), and the frame equals the oracle."""
    import torch
    sc = cases._with(scenes.scene_c4(), "c4_1080p", width=1920, height=1080)
    want, want_z, otm, rc = oracle.render_scene(sc)
    assert rc == 0
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    c = pkg.Context(0)
    fb = pkg.Framebuffer(sc.width, sc.height, c)
    c.set_textures(sc.textures)
    mesh = pkg.Mesh(c, sc.vertices, sc.faces)
    for _ in range(3):
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
    got, got_z = fb.download()
    free1, _ = torch.cuda.mem_get_info(0)
    mesh.free()
    c.close()
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))
    assert free0 - free1 < 200 << 20, f"{(free0 - free1) >> 20} MB taken"


def test_wireframe_absurd_edge_is_reported_not_walked(ctx, oracle):
    """A back-face edge whose end point saturates to +-2^31 would be billions of Bresenham steps (the reference walks them
    all): the device skips that edge and the call says so; everything else of the frame equals the oracle without it."""
    sc = cases.grid_mesh_scene(nx=6, ny=4)
    sc.settings.use_fixed_point = False                       # float projection: screen coordinates can be anything
    v = sc.vertices.copy()
    v["pos"][0] = (-3.0e30, 0.0, 14.0)                         # projects to ~6e31 -> `as i32` saturates; the face stays a back face
    bad = dataclasses.replace(sc, vertices=v)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    with pytest.raises(pkg.B32Error) as e:
        pkg.render_mesh_15(fb, bad.vertices, bad.faces, bad.textures, bad.camera, bad.settings)
    assert e.value.code == abi.B32_ERR_UNSUPPORTED
    got, _ = fb.download()
    assert ((got[..., :3] == [80, 80, 100]).all(-1)).sum() > 100      # the other edges were drawn


def test_ray_rs_projection_roundtrip_on_device(ctx):
    """The reference's own projection round-trip test (ray.rs:333-377) against the device's float projection."""
    import ctypes as C
    from test_oracle import ray_roundtrip_distance

    def project(v, cam, st, w, h):
        pkg.Framebuffer(w, h, ctx)
        scr = np.empty((len(v), 3), np.float32); cs = np.empty((len(v), 3), np.float32)
        ca = cam.to_abi(); sa, keep = st.to_abi()
        ctx.check(ctx.lib.b32_debug_transform(ctx.h, v.ctypes.data, len(v), C.byref(ca), C.byref(sa), scr.ctypes.data, cs.ctypes.data))
        return scr

    dist, scr = ray_roundtrip_distance(project)
    assert dist < 2.0, (dist, scr)


@pytest.mark.parametrize("zbuf", [True, False])
def test_sparse_calls_compose_over_existing_depth(ctx, oracle, zbuf):
    """Many small calls (a few surfaces per tile each) into one framebuffer: every call must honour the colour and depth
    the earlier ones left (a walk that stops early on a stale bound would not)."""
    fb = pkg.Framebuffer(320, 240, ctx)
    clear = scenes.CLEAR_COLOR
    fb.clear(clear)
    want = np.empty((240, 320, 4), np.uint8); want_z = np.empty((240, 320), np.float32)
    want[...] = np.array(list(clear) + [255], np.uint8); want_z[...] = np.finfo(np.float32).max
    for k in range(8):
        sc = scenes.scene_c2(n_tris=30 + 10 * k, seed=900 + k, use_zbuffer=zbuf)
        sc.vertices["pos"][:, :2] *= np.float32(0.4)            # crowd the surfaces into the middle tiles, various depths
        pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
        rc, _, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
        assert rc == 0
    got, got_z = fb.download()
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


# ---- Framebuffer::clear_gradient and the overlay line family (render.rs:60-77, :684-872) ---------------------
LINES = cases.line_cases()


@pytest.mark.parametrize("name,w,h,seed,lines", LINES, ids=lambda v: v if isinstance(v, str) else None)
def test_draw_lines(ctx, oracle, name, w, h, seed, lines):
    rgba, z = cases.line_background(w, h, seed)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    fb.draw_lines(lines)
    got, got_z = fb.download()
    want = rgba.copy()
    assert oracle.draw_lines(want, z, lines) == 0
    bad = (got != want).any(-1)
    assert not bad.any(), f"{name}: {bad.sum()} pixels differ, first at {np.argwhere(bad)[0][::-1]}"
    assert np.array_equal(got_z.view(np.uint32), z.view(np.uint32))          # lines never write depth


def test_draw_lines_one_by_one_equals_one_list(ctx, oracle):
    """The reference's per-line methods (one device pass each) and one list give the same framebuffer."""
    from bonnie32_b200 import raster
    name, w, h, seed, lines = LINES[3]
    rgba, z = cases.line_background(w, h, seed)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    for l in lines[:120]:
        c = (*l["rgb"].tolist(), int(l["blend"]))
        a = (int(l["x0"]), int(l["y0"]), int(l["x1"]), int(l["y1"]))
        k = int(l["kind"])
        if k == abi.LINE_2D and l["mode"] == abi.BLEND_OPAQUE: fb.draw_line(*a, c)
        elif k == abi.LINE_2D: fb.draw_line_blended(*a, c, int(l["mode"]))
        elif k == abi.LINE_2D_ALPHA: fb.draw_line_alpha(*a, c, int(l["alpha"]))
        elif k == abi.LINE_3D: fb.draw_line_3d(a[0], a[1], float(l["z0"]), a[2], a[3], float(l["z1"]), c)
        elif k == abi.LINE_3D_OVERLAY: fb.draw_line_3d_overlay(a[0], a[1], float(l["z0"]), a[2], a[3], float(l["z1"]), c)
        else: fb.draw_line_3d_alpha(a[0], a[1], float(l["z0"]), a[2], a[3], float(l["z1"]), c, int(l["alpha"]))
    got, _ = fb.download()
    want = rgba.copy()
    oracle.draw_lines(want, z, lines[:120])
    assert np.array_equal(got, want)


def test_draw_lines_over_rendered_scene(ctx, oracle):
    """Editor frame order: render the mesh, then depth-tested overlay lines against ITS z-buffer, no download between."""
    sc = scenes.scene_c2(n_tris=600, use_zbuffer=True)
    lines = cases.random_lines(sc.width, sc.height, 500, 99, kinds=(2, 3, 4))
    lines["z0"] = np.linspace(3.0, 60.0, len(lines), dtype=np.float32); lines["z1"] = lines["z0"][::-1]
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    fb.draw_lines(lines)
    got, got_z = fb.download()
    want, want_z, _, rc = oracle.render_scene(sc)
    assert rc == 0 and oracle.draw_lines(want, want_z, lines) == 0
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


def test_draw_lines_errors(ctx):
    from bonnie32_b200 import raster
    fb = pkg.Framebuffer(64, 48, ctx)
    fb.draw_lines(raster.make_lines([]))                                       # empty list: nothing to do
    bad_kind = raster.make_lines([raster.line_entry(17, 0, 0, 5, 5, (1, 2, 3))])
    with pytest.raises(pkg.B32Error) as e:
        fb.draw_lines(bad_kind)
    assert e.value.code == abi.B32_ERR_INVALID
    bad_mode = raster.make_lines([raster.line_entry(abi.LINE_2D, 0, 0, 5, 5, (1, 2, 3), mode=9)])
    with pytest.raises(pkg.B32Error) as e:
        fb.draw_lines(bad_mode)
    assert e.value.code == abi.B32_ERR_INVALID
    far = raster.make_lines([raster.line_entry(abi.LINE_2D, 0, 0, abi.LINE_MAX_COORD + 1, 5, (1, 2, 3))])
    with pytest.raises(pkg.B32Error) as e:
        fb.draw_lines(far)
    assert e.value.code == abi.B32_ERR_UNSUPPORTED
    before, _ = fb.download()
    assert not before.any()                                                    # rejected lists draw nothing


@pytest.mark.parametrize("w,h,top,bottom", [(320, 240, (10, 20, 200), (250, 128, 0)), (5, 1, (9, 8, 7), (200, 100, 50)),
                                            (641, 479, (0, 0, 0), (255, 255, 255)), (16, 97, (255, 0, 31, 5), (0, 255, 32))])
def test_clear_gradient(ctx, oracle, w, h, top, bottom):
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(np.full((h, w, 4), 9, np.uint8), np.zeros((h, w), np.float32))
    fb.clear_gradient(top, bottom)
    got, got_z = fb.download()
    want = np.empty((h, w, 4), np.uint8); want_z = np.empty((h, w), np.float32)
    oracle.fb_clear_gradient(want, want_z, top, bottom)
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


@pytest.mark.parametrize("kind", [abi.LINE_2D, abi.LINE_3D, abi.LINE_3D_ALPHA], ids=["2d", "3d", "3d_alpha"])
def test_draw_lines_every_slope(ctx, oracle, kind):
    """The device computes pixel k of a line in closed form; the reference walks an error term.  Every (dx, dy) of a
    41 x 41 neighbourhood, each in its own cell of a large framebuffer, plus the same slopes scaled by 7 and by 1000
    (64-bit path) through a small window."""
    R, cell = 20, 44
    side = (2 * R + 1) * cell
    ln = np.zeros((2 * R + 1) ** 2, dtype=abi.LINE_DTYPE)
    dy, dx = np.divmod(np.arange(len(ln)), 2 * R + 1)
    cx, cy = dx * cell + cell // 2, dy * cell + cell // 2
    ln["x0"], ln["y0"], ln["x1"], ln["y1"] = cx, cy, cx + dx - R, cy + dy - R
    ln["z0"], ln["z1"] = 4.0, 12.0
    ln["rgb"] = (np.arange(len(ln))[:, None] * np.array([7, 13, 29]) % 255 + 1).astype(np.uint8)
    ln["kind"], ln["alpha"] = kind, 200
    rgba = np.zeros((side, side, 4), np.uint8); rgba[..., 3] = 255
    z = np.full((side, side), 8.0, np.float32)
    for scale, (w, h) in ((1, (side, side)), (7, (side, side)), (1000, (97, 61))):
        l2 = ln.copy()
        if scale > 1:
            l2["x1"] = l2["x0"] + (l2["x1"] - l2["x0"]) * scale; l2["y1"] = l2["y0"] + (l2["y1"] - l2["y0"]) * scale
        if scale == 1000:
            l2["x0"] = 48 - (ln["x1"] - ln["x0"]) * 333; l2["y0"] = 30 - (ln["y1"] - ln["y0"]) * 333
            l2["x1"] = l2["x0"] + (ln["x1"] - ln["x0"]) * scale; l2["y1"] = l2["y0"] + (ln["y1"] - ln["y0"]) * scale
            l2["kind"] = abi.LINE_2D_ALPHA if kind == abi.LINE_2D else kind       # overlapping: make the order matter
        bg, bz = rgba[:h, :w].copy(), z[:h, :w].copy()
        fb = pkg.Framebuffer(w, h, ctx)
        fb.upload(bg, bz)
        fb.draw_lines(l2)
        got, _ = fb.download()
        want = bg.copy()
        oracle.draw_lines(want, bz, l2)
        bad = (got != want).any(-1)
        assert not bad.any(), f"scale {scale}: {bad.sum()} pixels differ, first at {np.argwhere(bad)[0][::-1]}"
        assert (want[..., :3] != 0).any()


@pytest.mark.parametrize("size", [1.0, 2.0, 3.0])
@pytest.mark.parametrize("name,w,h,cam", SKY[:4], ids=[s[0] for s in SKY[:4]])
def test_render_stars(ctx, oracle, name, w, h, cam, size):
    stars = cases.star_list(cam, w, h, time=0.75, count=2000 if w > 320 else 400)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.clear((3, 4, 5))
    fb.render_stars(stars, cam, size)
    got, got_z = fb.download()
    want = np.zeros((h, w, 4), np.uint8); want[...] = (3, 4, 5, 255)
    assert oracle.render_stars(want, stars, cam, size) == 0
    bad = (got != want).any(-1)
    assert not bad.any(), f"{name}: {bad.sum()} pixels differ"
    assert (got_z == np.finfo(np.float32).max).all()


def test_game_frame_without_downloads(ctx, oracle):
    """The game's frame (src/game/renderer.rs:91-179 + the debug lines of :1018-1047): clear, skybox sphere, stars,
    room mesh, depth-tested lines — all on the device, one download at the end."""
    name, w, h, cam = SKY[1]
    sv, f = cases.sky_mesh(cam.position)
    stars = cases.star_list(cam, w, h, time=3.0)
    sc = scenes.scene_c2(n_tris=500, use_zbuffer=True)
    lines = cases.random_lines(w, h, 60, 5, kinds=(2,))
    fb = pkg.Framebuffer(w, h, ctx)
    fb.clear((0, 0, 0))
    fb.render_skybox_mesh(sv, f, cam)
    fb.render_stars(stars, cam, 2.0)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    fb.draw_lines(lines)
    got, got_z = fb.download()
    want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
    want_z = np.full((h, w), np.finfo(np.float32).max, np.float32)
    oracle.render_skybox_mesh(want, sv, f, cam)
    oracle.render_stars(want, stars, cam, 2.0)
    rc, _, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    assert rc == 0 and oracle.draw_lines(want, want_z, lines) == 0
    assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


def test_two_host_threads_two_contexts(oracle):
    """One context per thread (include/b32_raster.h: a context is not thread-safe, contexts are independent): two host
    threads render different scenes at the same time, each into its own context."""
    import threading
    scs = [scenes.scene_c2(n_tris=800, use_zbuffer=False), cases.feature_scenes(300)[3]]
    out = [None, None]

    def work(i):
        c = pkg.Context(0)
        for _ in range(20):
            out[i] = render_gpu(c, scs[i])
        c.close()

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for i, sc in enumerate(scs):
        want, want_z, otm, rc = oracle.render_scene(sc)
        assert rc == 0 and out[i] is not None
        assert_same(sc, out[i][0], out[i][1], out[i][2], want, want_z, otm)


# ---- placed asset parts: render_asset_parts' per-object transform on the device (src/scene.rs:109-169) ----------
@pytest.mark.parametrize("rgb888", [False, True], ids=["rgb555", "rgb888"])
def test_render_placed_parts(ctx, oracle, rgb888):
    """A 'room' call, then the same resident part mesh placed four times (facing / position per object, one of them
    with no transform), enqueue-only where the call allows it: framebuffer and z-buffer equal the oracle drawing
    host-transformed copies, as the reference does."""
    from bonnie32_b200 import raster
    room = scenes.scene_c2(n_tris=300, use_zbuffer=True)
    part = cases.rgb888_scenes(120)[0] if rgb888 else cases.feature_scenes(120)[1]
    settings = dataclasses.replace(part.settings, use_zbuffer=True, backface_cull=False)
    placements = [(0.0, (0.0, 0.0, 0.0)), (0.7, (1.5, 0.25, 6.0)), (-2.2, (-2.0, -0.5, 12.0)), (3.1, (0.0, 1.0, 3.0))]
    fb = pkg.Framebuffer(room.width, room.height, ctx)
    fb.clear(room.clear)
    want, want_z = orc_fb(room)
    if rgb888:
        ctx.set_textures_rgb888(part.textures8)
    else:
        pkg.render_mesh_15(fb, room.vertices, room.faces, room.textures, room.camera, room.settings)
        rc, _, _ = oracle.render_mesh_15(want, want_z, room.vertices, room.faces, room.textures, room.camera, room.settings)
        assert rc == 0
        ctx.set_textures(part.textures)
    mesh = pkg.Mesh(ctx, part.vertices, part.faces)
    drawn, before = [], want.copy()
    for k, (facing, pos) in enumerate(placements):
        tm = mesh.render_placed(room.camera, settings, facing, pos, rgb888=rgb888, enqueue_only=(k % 2 == 1))
        v = oracle.place_vertices(part.vertices, facing, raster.libm_cosf(facing), raster.libm_sinf(facing), pos)
        if rgb888:
            rc, otm, _ = oracle.render_mesh(want, want_z, v, part.faces, part.textures8, room.camera, settings)
        else:
            rc, otm, _ = oracle.render_mesh_15(want, want_z, v, part.faces, part.textures, room.camera, settings)
        assert rc == 0
        if tm is not None:
            assert tm["triangles_drawn"] == otm["triangles_drawn"]
            drawn.append(tm["triangles_drawn"])
    got, got_z = fb.download()
    mesh.free()
    assert sum(drawn) > 0 and (want != before).any(-1).sum() > 500          # the placed parts are on screen
    bad = (got != want).any(-1)
    assert not bad.any(), f"{bad.sum()} pixels differ"
    assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))


def orc_fb(sc):
    rgba = np.empty((sc.height, sc.width, 4), np.uint8)
    a = 0 if (len(sc.clear) > 3 and sc.clear[3] == abi.BLEND_ERASE) else 255
    rgba[...] = (*sc.clear[:3], a)
    return rgba, np.full((sc.height, sc.width), np.finfo(np.float32).max, np.float32)


def test_download_view_equals_download(ctx):
    sc = scenes.scene_c2(n_tris=300, use_zbuffer=True)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings)
    a, az = fb.download()
    b, bz = fb.download_view(want_z=True)
    assert np.array_equal(a, b) and np.array_equal(az.view(np.uint32), bz.view(np.uint32))
    fb.resize(64, 48); fb.clear((1, 2, 3))
    c, _ = fb.download_view()
    assert c.shape == (48, 64, 4) and (c == (1, 2, 3, 255)).all()


# ---- filled primitives of the overlay family: draw_circle(_alpha), draw_filled_rect, draw_thick_line, draw_rect ----------
PRIMS = cases.prim_cases()


@pytest.mark.parametrize("name,w,h,seed,lines", PRIMS, ids=lambda v: v if isinstance(v, str) else None)
def test_draw_prims(ctx, oracle, name, w, h, seed, lines):
    rgba, z = cases.line_background(w, h, seed)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    fb.draw_lines(lines)
    got, got_z = fb.download()
    want = rgba.copy()
    assert oracle.draw_lines(want, z, lines) == 0
    bad = (got != want).any(-1)
    assert not bad.any(), f"{name}: {bad.sum()} pixels differ, first at {np.argwhere(bad)[0][::-1]}"
    assert np.array_equal(got_z.view(np.uint32), z.view(np.uint32))


def test_draw_prims_reference_methods(ctx, oracle):
    """The Framebuffer methods with the reference's names and arguments, one device pass each."""
    from bonnie32_b200 import raster
    w, h = 96, 72
    rgba, z = cases.line_background(w, h, 5)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    fb.draw_filled_rect(70, 60, 10, 20, (10, 200, 30))
    fb.draw_circle(30, 30, 12, (250, 10, 10))
    fb.draw_circle_alpha(50, 40, 20, (0, 0, 255), 100)
    fb.draw_thick_line(5, 65, 90, 8, 5, (255, 255, 0))
    fb.draw_thick_line(90, 65, 5, 30, 1, (0, 255, 255))
    fb.draw_rect(3, 3, 92, 68, (255, 255, 255))
    got, _ = fb.download()
    entries = [raster.line_entry(abi.LINE_FILLED_RECT, 70, 60, 10, 20, (10, 200, 30)), raster.line_entry(abi.LINE_CIRCLE, 30, 30, 12, 0, (250, 10, 10)),
               raster.line_entry(abi.LINE_CIRCLE_ALPHA, 50, 40, 20, 0, (0, 0, 255), alpha=100),
               raster.line_entry(abi.LINE_THICK, 5, 65, 90, 8, (255, 255, 0), z0=5.0), raster.line_entry(abi.LINE_THICK, 90, 65, 5, 30, (0, 255, 255), z0=1.0)]
    entries += raster.rect_entries(3, 3, 92, 68, (255, 255, 255))
    want = rgba.copy()
    assert oracle.draw_lines(want, z, raster.make_lines(entries)) == 0
    assert np.array_equal(got, want)
    for bad, code in ((raster.line_entry(abi.LINE_CIRCLE, 5, 5, 40000, 0, (1, 2, 3)), abi.B32_ERR_UNSUPPORTED),
                      (raster.line_entry(abi.LINE_THICK, 0, 0, 9, 9, (1, 2, 3), z0=2.5), abi.B32_ERR_INVALID),
                      (raster.line_entry(9, 0, 0, 9, 9, (1, 2, 3)), abi.B32_ERR_INVALID)):
        with pytest.raises(pkg.B32Error) as e:
            fb.draw_lines(raster.make_lines([bad]))
        assert e.value.code == code


# ---- round 2: both passes enqueued, binning without bins, C5, the reference binary itself -------------------
def test_enqueued_frames_with_pass2_replay_as_graphs(ctx, oracle):
    """render.rs:2561-2569 without a host round trip: frames whose meshes / textures hold semi-transparent surfaces (all
    blend modes, editor alpha, x-ray) are enqueued with b32_frame_15_enqueue — pass 1 + the ordered replay in one
    re-parameterised CUDA graph — and every frame equals the oracle while camera, settings and clear colour change."""
    by = {s.name: s for s in cases.feature_scenes(300)}
    g0 = ctx.graph_launches()
    for name in ("mixed_zbuffer", "mixed_painter", "mixed_xray", "mixed_zbuffer_nocull_gouraud"):
        sc = by[name]
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        ctx.set_textures(sc.textures)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        for i in range(6):
            st = dataclasses.replace(sc.settings, dithering=(i % 2 == 0), affine_textures=(i % 3 != 1))
            cam = cases._rotated_camera(0.015 * i, -0.02 * i, (0.1 * i, -0.05 * i, -0.2 * i))
            clear = (20 + i, 22, 28 + 2 * i)
            mesh.frame_enqueue(clear, cam, st, None)
            if i % 2 == 1:                                       # frames stay in flight in between
                got, got_z = fb.download()
                want = np.empty((sc.height, sc.width, 4), np.uint8); want_z = np.empty((sc.height, sc.width), np.float32)
                want[...] = np.array(list(clear) + [255], np.uint8); want_z[...] = np.finfo(np.float32).max
                rc, otm, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, cam, st, None)
                assert rc == 0
                assert np.array_equal(got, want), (name, i)
                assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (name, i)
        mesh.free()
    assert ctx.graph_launches() - g0 >= 4 * 3


def test_folded_clear_is_ordered_before_the_fill(ctx, oracle):
    """The frame's clear rides in k_setup and k_fill_opaque is launched behind it with programmatic dependent launch: the
    fill must not read the framebuffer before k_setup has completed.  Large framebuffer, small mesh (k_setup is a handful of
    CTAs and finishes its faces long before its share of the clear), a different camera and clear colour every frame, graph
    replay, many frames."""
    sc = scenes.scene_c2(n_tris=600, use_zbuffer=True)
    w, h = 1280, 960
    fb = pkg.Framebuffer(w, h, ctx)
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    for i in range(24):
        cam = cases._rotated_camera(0.01 * (i % 5), 0.02 * (i % 7) - 0.05, (0.3 * (i % 3), 0.0, -1.0 * (i % 4)))
        clear = (10 * (i % 20), 255 - 9 * i, 28 + i)
        mesh.frame_enqueue(clear, cam, sc.settings, None)
        if i % 4 == 3 or i < 3:
            got, got_z = fb.download()
            want = np.empty((h, w, 4), np.uint8); want_z = np.empty((h, w), np.float32)
            want[...] = np.array(list(clear) + [255], np.uint8); want_z[...] = np.finfo(np.float32).max
            rc, otm, _ = oracle.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, cam, sc.settings, None)
            assert rc == 0
            assert np.array_equal(got, want), i
            assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), i
    mesh.free()


def test_c5_all_eight_frames(ctx, oracle):
    """BASELINE config 5: the eight independent 100k-triangle frames, each bit-exact against the oracle AND against the
    committed golden hashes, rendered with frames in flight on one context (enqueued, one download per frame)."""
    import hashlib, json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hashes.json")))
    for k in range(8):
        sc = scenes.scene_c5(k)
        want, want_z, otm, rc = oracle.render_scene(sc)
        assert rc == 0
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        ctx.set_textures(sc.textures)
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, None)
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, None)
        got, got_z = fb.download()
        mesh.free()
        assert np.array_equal(got, want), k
        assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), k
        assert hashlib.sha256(got.tobytes()).hexdigest() == gold[sc.name]["rgba_sha256"]


def test_c5_frames_on_two_devices(oracle):
    """Different C5 frames on different GPUs of one box (one context per device), checked frame by frame."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctxs = [pkg.Context(d) for d in range(2)]
    jobs = []
    for k in range(4):
        c = ctxs[k % 2]
        sc = scenes.scene_c5(k)
        fb = pkg.Framebuffer(sc.width, sc.height, c)
        c.set_textures(sc.textures)
        mesh = pkg.Mesh(c, sc.vertices, sc.faces)
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, None)
        jobs.append((sc, fb, mesh))
        if k % 2 == 1:                              # both devices busy: collect the pair
            for sc_, fb_, mesh_ in jobs:
                got, got_z = fb_.download()
                want, want_z, otm, rc = oracle.render_scene(sc_)
                assert np.array_equal(got, want), sc_.name
                assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), sc_.name
                mesh_.free()
            jobs = []
    for c in ctxs:
        c.close()


def _expand_shl3(rgba_wasm, clear):
    """The 0.1.8 binary writes 5-bit channels as `v << 3`, the 0.1.11 source as `(v << 3) | (v >> 2)` (oracle/wasm/DRIFT.md
    item 3).  For frames without blending a written pixel is recognisable (alpha 255, low three bits 0 in every channel,
    unlike the clear colour), so the 0.1.11 bytes follow from the binary's by re-expansion."""
    out = rgba_wasm.copy()
    is_clear = (rgba_wasm[..., :3] == np.array(clear[:3], np.uint8)).all(-1)
    v = rgba_wasm[..., :3] >> 3
    out[..., :3] = np.where(is_clear[..., None], rgba_wasm[..., :3], (v << 3) | (v >> 2))
    return out


REFBIN_DIRECT = ["float_projection", "c2_1000_float", "rotated_camera_large_world_float", "ortho", "c4_100000_float_nodither"]


@pytest.mark.parametrize("name", REFBIN_DIRECT)
def test_gpu_against_reference_binary_directly(ctx, name):
    """CUDA vs the reference's own compiled code with NO oracle in between, on the scenes where the 0.1.8 binary and the
    0.1.11 source differ only by the 5->8-bit expansion of written pixels (float / ortho projection: no fixed-point divide;
    textured or vertex-coloured faces: the dither rule agrees; opaque faces only: no blending, no transparency rule).
    Fixtures: tests/golden/ref_wasm (sha256 of the binary's framebuffer + z-buffer; full frames where kept)."""
    import hashlib, json, os
    import refbin_cases
    here = os.path.dirname(__file__)
    fix = json.load(open(os.path.join(here, "golden", "ref_wasm", "render_mesh_15.json")))["scenes"]
    direct = np.load(os.path.join(here, "golden", "ref_wasm", "direct_frames.npz"))
    sc = {s.name: s for s in refbin_cases.small_scenes() + refbin_cases.big_scenes()}[name]
    assert fix[name]["inputs"] == refbin_cases.inputs_digest(sc)
    assert (sc.clear[0] & 7) and sc.settings.use_zbuffer is not None
    got, got_z, tm = render_gpu(ctx, sc)
    assert tm["triangles_drawn"] == fix[name]["drawn"]
    assert hashlib.sha256(np.ascontiguousarray(got_z, "<f4").tobytes()).hexdigest() == fix[name]["z"], "z-buffer differs from the reference binary"
    want = _expand_shl3(direct[name], sc.clear)
    assert hashlib.sha256(direct[name].tobytes()).hexdigest() == fix[name]["rgba"]
    assert np.array_equal(got, want), "framebuffer differs from the reference binary (after 5->8-bit re-expansion)"


def test_compact_marshalling_variants(ctx, oracle):
    """b32_render_mesh_15_ex with B32_VTX_NO_NORMAL / B32_FACES_IMPLICIT / B32_FACES_UNIFORM: fewer bytes across PCIe, same framebuffer.
    Blocking and enqueued; C4-style soup (both flags), an indexed mesh (no-normal only), lit scenes (implicit only)."""
    import ctypes as C
    lib = ctx.lib
    by = {s.name: s for s in cases.feature_scenes(300)}
    grid = cases.grid_mesh_scene(nx=20, ny=12, flip=False)
    grid.settings.shading = abi.SHADE_NONE; grid.settings.backface_wireframe = False
    for sc, expect in ((scenes.scene_c4(n_tris=30000), abi.VTX_NO_NORMAL | abi.FACES_IMPLICIT), (by["zbuffer_idx8"], abi.VTX_NO_NORMAL | abi.FACES_IMPLICIT),
                       (by["gouraud_lights"], abi.FACES_IMPLICIT), (by["mixed_zbuffer"], abi.VTX_NO_NORMAL | abi.FACES_IMPLICIT), (grid, abi.VTX_NO_NORMAL)):
        want, want_z, otm, rc = oracle.render_scene(sc)
        assert rc == 0
        v, f, flags = abi.compact_buffers(sc.vertices, sc.faces, sc.settings.shading == abi.SHADE_NONE)
        uniform = bool(expect & abi.FACES_IMPLICIT) and bool((sc.faces["flags"] == sc.faces["flags"][0]).all())
        assert flags == ((expect & ~abi.FACES_IMPLICIT) | abi.FACES_UNIFORM if uniform else expect), sc.name
        assert len(f) == (1 if uniform else len(sc.faces))
        layouts = [(f, flags)]
        if uniform:                                            # the same soup with one flags word per face
            layouts.append((np.ascontiguousarray(sc.faces["flags"]), (flags & ~abi.FACES_UNIFORM) | abi.FACES_IMPLICIT))
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        ctx.set_textures(sc.textures)
        cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi()
        fog = pkg.raster.fog_to_abi(sc.fog)
        for f_, flags_ in layouts:
            for asyn in (0, abi.RENDER_ASYNC):
                fb.clear(sc.clear)
                tm = abi.Timings()
                ctx.check(lib.b32_render_mesh_15_ex(ctx.h, v.ctypes.data, len(v), f_.ctypes.data, len(sc.faces), C.byref(cam), C.byref(st),
                                                    C.byref(fog) if fog is not None else None, flags_ | asyn, C.byref(tm)))
                got, got_z = fb.download()
                assert np.array_equal(got, want), (sc.name, asyn, flags_)
                assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (sc.name, asyn, flags_)
                if not asyn:
                    assert tm.triangles_drawn == otm["triangles_drawn"]
    # the promise is checked
    sc = by["gouraud_lights"]
    v, f, flags = abi.compact_buffers(sc.vertices, sc.faces, True)
    cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi()
    nf = len(sc.faces)
    assert lib.b32_render_mesh_15_ex(ctx.h, v.ctypes.data, len(v), f.ctypes.data, nf, C.byref(cam), C.byref(st), None, flags, None) == abi.B32_ERR_INVALID
    fl = np.ascontiguousarray(sc.faces["flags"])
    assert lib.b32_render_mesh_15_ex(ctx.h, v.ctypes.data, len(v) - 1, fl.ctypes.data, nf, C.byref(cam), C.byref(st), None, abi.FACES_IMPLICIT, None) == abi.B32_ERR_OOB_INDEX
    assert lib.b32_render_mesh_15_ex(ctx.h, v.ctypes.data, len(v) - 1, fl.ctypes.data, nf, C.byref(cam), C.byref(st), None, abi.FACES_UNIFORM, None) == abi.B32_ERR_OOB_INDEX


def test_skybox_against_reference_binary_directly(ctx):
    """b32_render_skybox_mesh + b32_render_stars vs the frames the reference's own compiled Framebuffer::render_skybox
    produced (tests/golden/ref_wasm/skybox.*: mesh and star list as the binary built them), with no oracle in between."""
    import hashlib, json, os
    here = os.path.dirname(__file__)
    meta = json.load(open(os.path.join(here, "golden", "ref_wasm", "skybox.json")))["cases"]
    arr = np.load(os.path.join(here, "golden", "ref_wasm", "skybox.npz"))
    for key, m in sorted(meta.items()):
        cam = cases._rotated_camera(m["camera"][0], m["camera"][1], m["camera"][2])
        fb = pkg.Framebuffer(m["width"], m["height"], ctx)
        fb.upload(np.zeros((m["height"], m["width"], 4), np.uint8), np.full((m["height"], m["width"]), np.finfo(np.float32).max, np.float32))
        fb.render_skybox_mesh(arr[key + "_verts"], arr[key + "_faces"], cam)
        if m["n_stars"]:
            fb.render_stars(arr[key + "_stars"], cam, m["star_size"])
        got, _ = fb.download()
        if key + "_frame" in arr:
            bad = (arr[key + "_frame"] != got).any(-1)
            assert not bad.any(), (key, int(bad.sum()), np.argwhere(bad)[0][::-1])
        assert hashlib.sha256(got.tobytes()).hexdigest() == m["rgba"], key


def test_frame_timings_of_enqueued_frames(oracle):
    """b32_ctx_frame_timings / b32_frame_timings: RasterTimings (types.rs:1499-1514) for frames that were only enqueued —
    exact triangles_drawn, kernel times from the device's own clock, no effect on the pixels, never blocks."""
    import ctypes as C
    c = pkg.Context(0)
    try:
        lib = c.lib
        tm = abi.Timings()
        assert lib.b32_frame_timings(c.h, C.byref(tm)) == abi.B32_ERR_INVALID         # not enabled yet
        c.check(lib.b32_ctx_frame_timings(c.h, 1))
        c.check(lib.b32_frame_timings(c.h, C.byref(tm)))
        assert tm.triangles_drawn == 0 and tm.draw_ms == 0.0                       # nothing has finished
        by = {s.name: s for s in cases.feature_scenes(400)}
        for sc in (scenes.scene_c2(n_tris=700), by["mixed_zbuffer"], by["xray"]):
            want, want_z, otm, rc = oracle.render_scene(sc)
            fb = pkg.Framebuffer(sc.width, sc.height, c)
            c.set_textures(sc.textures)
            mesh = pkg.Mesh(c, sc.vertices, sc.faces)
            for _ in range(4):                                                    # plain launches, then graph replays
                mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
            got, got_z = fb.download()
            assert np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), sc.name
            c.check(lib.b32_frame_timings(c.h, C.byref(tm)))
            assert tm.triangles_drawn == otm["triangles_drawn"], sc.name
            assert 0.0 < tm.cull_ms < 5.0 and 0.0 < tm.draw_ms < 50.0, (sc.name, tm.cull_ms, tm.draw_ms)
            mesh.free()
        c.check(lib.b32_ctx_frame_timings(c.h, 0))
    finally:
        c.close()


# ---- Spot lights (render.rs:1038-1059; acos = the shipped build's libm acosf, k_setup<true>) -------------------------------
SPOT15 = cases.spot_scenes()
SPOT888 = cases.spot_scenes888()


@pytest.mark.parametrize("sc", SPOT15 + SPOT888, ids=[s.name for s in SPOT15 + SPOT888])
def test_spot_light_scene(ctx, oracle, sc):
    rgb888 = not sc.settings.use_rgb555
    want, want_z, otm, rc = (oracle.render_scene888 if rgb888 else oracle.render_scene)(sc)
    assert rc == 0
    got, got_z, tm = (render_gpu888 if rgb888 else render_gpu)(ctx, sc)
    assert_same(sc, got, got_z, tm, want, want_z, otm)


def test_spot_lights_at_scale_blocking_and_enqueued(ctx, oracle):
    """30 000 Gouraud-lit triangles under the five spot lights: host buffers, resident mesh, and enqueued frames that
    alternate between a Spot and a Point light list (the two k_setup instantiations under one frame topology)."""
    sc = scenes.scene_c4(n_tris=30000)
    un = scenes.splitmix64_u01(91, len(sc.vertices) * 3).reshape(-1, 3)
    sc.vertices = sc.vertices.copy()
    sc.vertices["normal"] = (2.0 * un - 1.0).astype(np.float32)
    spot = cases._with(sc, "c4_30000_spot", shading=abi.SHADE_GOURAUD, lights=cases.spot_lights(), ambient=0.1)
    point = cases._with(sc, "c4_30000_point", shading=abi.SHADE_GOURAUD, lights=[pkg.Light.point((0.0, 0.0, 0.0), 70.0, 1.8)], ambient=0.1)
    wants = {}
    for s in (spot, point):
        want, want_z, otm, rc = oracle.render_scene(s)
        assert rc == 0
        wants[s.name] = (want, want_z, otm)
        for resident in (False, True):
            got, got_z, tm = render_gpu(ctx, s, resident=resident)
            assert_same(s, got, got_z, tm, want, want_z, otm)
    assert not np.array_equal(wants[spot.name][0], wants[point.name][0])
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    for k in range(8):
        s = spot if k % 2 == 0 else point
        mesh.frame_enqueue(sc.clear, s.camera, s.settings, s.fog)
        got, got_z = fb.download()                         # syncs; frames 3.. are graph launches
        assert np.array_equal(got, wants[s.name][0]), (k, s.name)
        assert np.array_equal(got_z.view(np.uint32), wants[s.name][1].view(np.uint32)), (k, s.name)
    mesh.free()


def test_spot_lights_against_reference_binary_directly(ctx):
    """CUDA vs the reference's own compiled code with no oracle in between: flat shading under Directional + Spot + Point
    lights, float projection, no dither, opaque faces (the conditions of test_gpu_against_reference_binary_directly)."""
    import hashlib, json, os
    import refbin_cases
    here = os.path.dirname(__file__)
    name = "spot_flat_mixed_float_nodither"
    fix = json.load(open(os.path.join(here, "golden", "ref_wasm", "spot.json")))["scenes"][name]
    frame = np.load(os.path.join(here, "golden", "ref_wasm", "spot.npz"))["frame/" + name]
    sc = {s.name: s for s in refbin_cases.spot_scenes()}[name]
    assert fix["inputs"] == refbin_cases.inputs_digest(sc)
    got, got_z, tm = render_gpu(ctx, sc)
    assert tm["triangles_drawn"] == fix["drawn"]
    assert hashlib.sha256(np.ascontiguousarray(got_z, "<f4").tobytes()).hexdigest() == fix["z"], "z-buffer differs from the reference binary"
    assert hashlib.sha256(frame.tobytes()).hexdigest() == fix["rgba"]
    assert np.array_equal(got, _expand_shl3(frame, sc.clear)), "framebuffer differs from the reference binary (after 5->8-bit re-expansion)"
