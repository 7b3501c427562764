"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Fuzz of b32_render_mesh_placed (one part of render_asset_parts,
src/scene.rs:109-169): random parts rotated and moved on the device, both colour paths, blocking and enqueued (the enqueued
RGB888 call is the only enqueue path of render_mesh), several parts composed into one frame.
usage (GPU box): python tests/checks/fuzz_placed.py [n_frames] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
from bonnie32_b200 import raster
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
first = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
ctx = pkg.Context(0)
t0 = time.time(); bad = parts = refused = 0
for seed in range(first, first + n):
    rng = np.random.default_rng(seed ^ 0x9A)
    rgb888 = bool(rng.random() < 0.5)
    enq = bool(rng.random() < 0.5)
    base = fuzz.fuzz_scene(seed, rgb888, n_tris=int(rng.choice([30, 120, 400])))
    base.settings.backface_wireframe = False; base.settings.wireframe_overlay = False
    w, h = base.width, base.height
    fb = pkg.Framebuffer(w, h, ctx); fb.clear(base.clear)
    want = np.empty((h, w, 4), np.uint8); want[...] = np.array(list(base.clear[:3]) + [255], np.uint8)
    want_z = np.full((h, w), np.finfo(np.float32).max, np.float32)
    if rgb888: ctx.set_textures_rgb888(base.textures8)
    else: ctx.set_textures(base.textures)
    failed = False
    meshes = []
    for k in range(int(rng.integers(1, 4))):
        sc = fuzz.fuzz_scene(seed * 7 + k, rgb888, n_tris=int(rng.choice([30, 120, 400])))
        pos = sc.vertices["pos"]; pos[~np.isfinite(pos)] = np.float32(0.5)
        facing = float(rng.random() * 6.3 - 3.0); wp = (rng.normal(size=3) * np.array([1.0, 1.0, 4.0])).astype(np.float32)
        v = orc.place_vertices(sc.vertices, facing, raster.libm_cosf(facing), raster.libm_sinf(facing), wp)
        bw, bz = want.copy(), want_z.copy()
        if rgb888: rc, otm, _ = orc.render_mesh(want, want_z, v, sc.faces, base.textures8, base.camera, base.settings)
        else: rc, otm, _ = orc.render_mesh_15(want, want_z, v, sc.faces, base.textures, base.camera, base.settings, base.fog)
        if rc != 0: want[...] = bw; want_z[...] = bz
        mesh = pkg.Mesh(ctx, sc.vertices, sc.faces); meshes.append(mesh)
        code = 0
        try:
            mesh.render_placed(base.camera, base.settings, facing, wp, None if rgb888 else base.fog, rgb888=rgb888, enqueue_only=enq)
            if enq: ctx.sync()
        except pkg.B32Error as e:
            code = e.code
        parts += 1
        if code == pkg.abi.B32_ERR_UNSUPPORTED and enq: refused += 1; failed = True; break
        if code != rc:
            print("MISMATCH (error code)", seed, k, "rgb888" if rgb888 else "rgb555", "enq" if enq else "blocking", code, rc); bad += 1; failed = True; break
    if not failed:
        got, got_z = fb.download()
        zs = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        if not (np.array_equal(got, want) and zs):
            print("MISMATCH seed", seed, "rgb888" if rgb888 else "rgb555", "enq" if enq else "blocking", "pixels", int((got != want).any(-1).sum())); bad += 1
    for m in meshes: m.free()
print(f"seeds {first}..{first + n - 1}: {n} frames, {parts} placed parts (both colour paths, blocking and enqueued), {refused} refused")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
