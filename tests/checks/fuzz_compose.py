"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Composition fuzz: several random calls (tests/fuzz.py scenes, RGB555
and RGB888 mixed) drawn one after another into ONE framebuffer without a clear in between — every call must honour the colour
and depth (NaN depths included) the earlier ones left.  Compared with the oracle after every call.
usage (GPU box): python tests/checks/fuzz_compose.py [n_groups] [first_seed] [calls_per_group]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
first = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
per = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = pkg.Context(0)
t0 = time.time()
bad = calls = panics = nan_groups = 0
for gidx in range(n):
    rng = np.random.default_rng(gidx + first)
    w, h = [(320, 240), (200, 150), (333, 77), (64, 64), (640, 480)][int(rng.integers(0, 5))]
    fb = pkg.Framebuffer(w, h, ctx)
    clear = tuple(int(x) for x in rng.integers(0, 256, 3))
    fb.clear(clear)
    want = np.empty((h, w, 4), np.uint8); want[...] = np.array(list(clear) + [255], np.uint8)
    want_z = np.full((h, w), np.finfo(np.float32).max, np.float32)
    for c in range(per):
        seed = first + gidx * per + c
        rgb888 = bool(rng.random() < 0.5)
        sc = fuzz.fuzz_scene(seed, rgb888, n_tris=int(rng.choice([30, 120, 400])))
        sc.settings.backface_wireframe = False; sc.settings.wireframe_overlay = False
        before_w, before_z = want.copy(), want_z.copy()
        if rgb888:
            rc, otm, _ = orc.render_mesh(want, want_z, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
        else:
            rc, otm, _ = orc.render_mesh_15(want, want_z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        try:
            if rgb888:
                tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
            else:
                tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
            err = 0
        except pkg.B32Error as e:
            err = e.code
        if rc != 0:                                        # the reference panics: nothing may have been drawn
            want[...] = before_w; want_z[...] = before_z
            panics += 1
        calls += 1
        got, got_z = fb.download()
        zsame = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        if err != rc or not np.array_equal(got, want) or not zsame or (rc == 0 and tm["triangles_drawn"] != otm["triangles_drawn"]):
            print("MISMATCH group", gidx, "call", c, "seed", seed, "rgb888" if rgb888 else "rgb555", "rc", rc, err,
                  "pixels", int((got != want).any(-1).sum()), "z", int((~((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z)))).sum()))
            bad += 1
            break
    nan_groups += int(np.isnan(want_z).any())
print(f"groups {first}..{first + n - 1} x {per} calls: {calls} calls compared after each ({panics} reference panics), "
      f"{nan_groups} groups ended with NaN depths in the z-buffer")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
