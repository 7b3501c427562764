"""TEST INFRASTRUCTURE (uses the oracle as the checker; lives under tests/ for that reason).  Small end-to-end run for compute-sanitizer (GPU box): every kernel of the library on small scenes, checked vs the oracle."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import cases
from bonnie32_b200 import scenes

ctx = pkg.Context(0)
bad = 0
by = {s.name: s for s in cases.feature_scenes(100)}
for name in ("painter_idx8", "zbuffer_idx8", "mixed_zbuffer", "xray", "gouraud_lights", "fog", "rotated_camera_large_world_float"):
    sc = by[name]
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    got, gz = fb.download()
    want, wz, _, rc = orc.render_scene(sc)
    ok = rc == 0 and np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32))
    print(name, "OK" if ok else "MISMATCH"); bad += not ok
for sc in cases.wireframe_scenes(60)[:2] + [s for s in cases.rgb888_scenes(100) if s.name in ("rgb888_opaque_zbuffer", "rgb888_mixed_painter")]:
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
    if sc.textures8 is not None:
        pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
        want, wz, _, rc = orc.render_scene888(sc)
    else:
        pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        want, wz, _, rc = orc.render_scene(sc)
    got, gz = fb.download()
    ok = rc == 0 and np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32))
    print(sc.name, "OK" if ok else "MISMATCH"); bad += not ok
# float projection with semi-transparent surfaces and x-ray: the shared-edge-prefix instantiations of both fill kernels
for name in ("mixed_zbuffer", "xray"):
    sc = cases._with(by[name], name + "_float", use_fixed_point=False)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    got, gz = fb.download()
    want, wz, _, rc = orc.render_scene(sc)
    ok = rc == 0 and np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32))
    print(sc.name, "OK" if ok else "MISMATCH"); bad += not ok
# Spot lights: the k_setup<true> instantiation (acos path, the one kernel with a stack frame)
for sc in [s for s in cases.spot_scenes(100) if s.name in ("spot_gouraud_all", "spot_flat_mixed_float_nodither", "spot_mixed_blend_gouraud")]:
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    got, gz = fb.download()
    want, wz, _, rc = orc.render_scene(sc)
    ok = rc == 0 and np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32))
    print(sc.name, "OK" if ok else "MISMATCH"); bad += not ok
# C4-style atlas (TMA-staged mask), enqueue + graph replay, skybox
sc = scenes.scene_c4(n_tris=3000)
fb = pkg.Framebuffer(sc.width, sc.height, ctx); ctx.set_textures(sc.textures)
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(4):
    mesh.frame_enqueue(sc.clear, sc.camera, sc.settings)
got, gz = fb.download()
want, wz, _, rc = orc.render_scene(sc)
ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)) and ctx.graph_launches() >= 2
print("c4_3000_graph_replay", "OK" if ok else "MISMATCH"); bad += not ok
name, w, h, cam = cases.sky_cases()[1]
sv, f = cases.sky_mesh(cam.position)
fb = pkg.Framebuffer(w, h, ctx); fb.clear((0, 0, 0)); fb.render_skybox_mesh(sv, f, cam)
got, _ = fb.download()
want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
orc.render_skybox_mesh(want, sv, f, cam)
ok = np.array_equal(got, want); print(name, "OK" if ok else "MISMATCH"); bad += not ok
# overlay lines (all kinds, blended stacks) and the gradient clear
for name, w, h, seed, lines in [c for c in cases.line_cases() if c[0] in ("lines_all_kinds_dense", "lines_star_alpha", "lines_far_endpoints")]:
    rgba, z = cases.line_background(w, h, seed)
    fb = pkg.Framebuffer(w, h, ctx); fb.upload(rgba, z); fb.draw_lines(lines)
    got, _ = fb.download()
    want = rgba.copy(); orc.draw_lines(want, z, lines)
    ok = np.array_equal(got, want); print(name, "OK" if ok else "MISMATCH"); bad += not ok
fb = pkg.Framebuffer(61, 47, ctx); fb.clear_gradient((10, 20, 200), (250, 128, 0))
got, gz = fb.download()
want = np.empty((47, 61, 4), np.uint8); wz = np.empty((47, 61), np.float32); orc.fb_clear_gradient(want, wz, (10, 20, 200), (250, 128, 0))
ok = np.array_equal(got, want) and np.array_equal(gz, wz); print("clear_gradient", "OK" if ok else "MISMATCH"); bad += not ok
# star pass and a placed part
name, w, h, cam = cases.sky_cases()[1]
stars = cases.star_list(cam, w, h, time=0.5)
fb = pkg.Framebuffer(w, h, ctx); fb.clear((0, 0, 0)); fb.render_stars(stars, cam, 3.0)
got, _ = fb.download()
want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
orc.render_stars(want, stars, cam, 3.0)
ok = np.array_equal(got, want); print("stars", "OK" if ok else "MISMATCH"); bad += not ok
from bonnie32_b200 import raster
sc = by["zbuffer_idx8"]
fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear); ctx.set_textures(sc.textures)
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
mesh.render_placed(sc.camera, sc.settings, 0.6, (0.5, 0.0, 2.0))
got, gz = fb.download()
v = orc.place_vertices(sc.vertices, 0.6, raster.libm_cosf(0.6), raster.libm_sinf(0.6), (0.5, 0.0, 2.0))
want = np.empty((sc.height, sc.width, 4), np.uint8); want[...] = (*sc.clear[:3], 255)
wz = np.full((sc.height, sc.width), np.finfo(np.float32).max, np.float32)
orc.render_mesh_15(want, wz, v, sc.faces, sc.textures, sc.camera, sc.settings)
ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)); print("placed_part", "OK" if ok else "MISMATCH"); bad += not ok
# filled overlay primitives
name, w, h, seed, lines = cases.prim_cases()[1]
rgba, z = cases.line_background(w, h, seed)
fb = pkg.Framebuffer(w, h, ctx); fb.upload(rgba, z); fb.draw_lines(lines)
got, _ = fb.download()
want = rgba.copy(); orc.draw_lines(want, z, lines)
ok = np.array_equal(got, want); print(name, "OK" if ok else "MISMATCH"); bad += not ok
# round 2: both passes enqueued (graph replay), compact marshalling, coarse mask tiles, several candidate windows per tile
import ctypes as C
from bonnie32_b200 import abi
sc = by["mixed_zbuffer"]
fb = pkg.Framebuffer(sc.width, sc.height, ctx); ctx.set_textures(sc.textures)
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(4):
    mesh.frame_enqueue(sc.clear, sc.camera, sc.settings)
got, gz = fb.download()
want, wz, _, rc = orc.render_scene(sc)
ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)); print("mixed_enqueued_pass2", "OK" if ok else "MISMATCH"); bad += not ok
sc = scenes.scene_c4(n_tris=2000)
v, f, flags = abi.compact_buffers(sc.vertices, sc.faces, True)
fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear); ctx.set_textures(sc.textures)
cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi()
want, wz, _, rc = orc.render_scene(sc)
fl = np.ascontiguousarray(sc.faces["flags"])
for name, f_, flags_ in (("compact_marshalling_uniform", f, flags), ("compact_marshalling_implicit", fl, (flags & ~abi.FACES_UNIFORM) | abi.FACES_IMPLICIT)):
    fb.clear(sc.clear)
    ctx.check(ctx.lib.b32_render_mesh_15_ex(ctx.h, v.ctypes.data, len(v), f_.ctypes.data, len(sc.faces), C.byref(cam), C.byref(st), None, flags_, None))
    got, gz = fb.download()
    ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)); print(name, "OK" if ok else "MISMATCH"); bad += not ok
# enqueued frames that publish their timings (host-mapped status ring)
ctx.check(ctx.lib.b32_ctx_frame_timings(ctx.h, 1))
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(4):
    mesh.frame_enqueue(sc.clear, sc.camera, sc.settings)
got, gz = fb.download()
tmx = abi.Timings(); ctx.check(ctx.lib.b32_frame_timings(ctx.h, C.byref(tmx)))
ok = np.array_equal(got, want) and tmx.triangles_drawn > 0 and tmx.draw_ms > 0; print("frame_timings_enqueued", "OK" if ok else "MISMATCH"); bad += not ok
ctx.check(ctx.lib.b32_ctx_frame_timings(ctx.h, 0))
sc = cases._with(by["mixed_zbuffer"], "mixed_1920x1080", width=1920, height=1080)
fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
got, gz = fb.download()
want, wz, _, rc = orc.render_scene(sc)
ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)); print("coarse_mask_tiles_1080p", "OK" if ok else "MISMATCH"); bad += not ok
n = 2600                                                    # > 2 windows of 1 024 candidates in the centre tiles; > 2 048 ordered entries
u = scenes.splitmix64_u01(5150, n * 9).reshape(n, 3, 3)
pos = np.empty((n, 3, 3)); pos[..., 0] = (u[..., 0] - 0.5) * 0.9; pos[..., 1] = (u[..., 1] - 0.5) * 0.9; pos[..., 2] = 10.0 + 30.0 * u[..., 2]
vv = scenes.make_vertices(pos.reshape(-1, 3), uv=u[..., :2].reshape(-1, 2), rgba=np.concatenate([np.floor(u * 255).reshape(-1, 3), np.zeros((n * 3, 1))], axis=1))
ff = scenes.make_faces(np.arange(n * 3).reshape(n, 3), tex_id=abi.FACE_TEX_NONE)
for xray in (False, True):
    sc = scenes.Scene("crowded", vv, ff, [], pkg.Camera(), scenes.common_settings(use_zbuffer=True, backface_cull=False, xray_mode=xray))
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
    pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    got, gz = fb.download()
    want, wz, _, rc = orc.render_scene(sc)
    ok = np.array_equal(got, want) and np.array_equal(gz.view(np.uint32), wz.view(np.uint32)); print("crowded_tile_xray" if xray else "crowded_tile_windows", "OK" if ok else "MISMATCH"); bad += not ok
print("mismatches:", bad)
sys.exit(1 if bad else 0)
