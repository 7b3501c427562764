"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Fuzz of the Spot-light path (k_setup<true>, render.rs:1038-1059)
against the oracle: the scenes of tests/fuzz.py with 1-4 random Spot lights spliced into their light lists (random cone
angles incl. 0 and > pi, directions a hair off unit length so that |dot| > 1 reaches acos, disabled ones), Gouraud and
flat shading, both colour paths, blocking and — for RGB555 — enqueued three times (graph replay).
usage (GPU box): python tests/checks/fuzz_spot.py [n_scenes_per_path] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from bonnie32_b200 import abi
from oracle import oracle as orc
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
first = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ctx = pkg.Context(0)
t0 = time.time()
bad = 0


def add_spots(sc, seed):
    rng = np.random.default_rng(990000 + seed)
    ls = list(sc.settings.lights)
    for _ in range(int(rng.integers(1, 5))):
        l = pkg.Light.spot(rng.normal(size=3) * 8 + np.array([0, 0, 10.0]), rng.normal(size=3),
                           float(rng.choice([0.0, 0.2, 0.7, 1.5, 2.5, 3.3]) * rng.random() if rng.random() < 0.9 else -0.3),
                           float(rng.choice([0.0, 5.0, 30.0, 90.0, 1e6])), float(rng.random() * 3))
        if rng.random() < 0.3:
            l.direction = (l.direction * np.float32(1.0 + rng.random() * 1e-6)).astype(np.float32)
        if rng.random() < 0.05:
            l.direction = np.zeros(3, np.float32)
        l.color = tuple(int(x) for x in rng.integers(0, 256, size=3))
        l.enabled = bool(rng.random() < 0.9)
        ls.insert(int(rng.integers(0, len(ls) + 1)), l)
    sc.settings.lights = ls
    sc.settings.shading = int(abi.SHADE_GOURAUD if rng.random() < 0.6 else abi.SHADE_FLAT)
    return sc


for rgb888 in (False, True):
    ok = panics = enq = 0
    for seed in range(first, first + n):
        nt = int(np.random.default_rng(seed ^ 0x5EED).choice([30, 120, 120, 400, 1500]))
        sc = add_spots(fuzz.fuzz_scene(seed, rgb888, n_tris=nt), seed)
        want, want_z, otm, rc = (orc.render_scene888 if rgb888 else orc.render_scene)(sc)
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        fb.clear(sc.clear)
        try:
            if rgb888:
                tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
            else:
                tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        except pkg.B32Error as e:
            if e.code != rc:
                print("MISMATCH (error path)", "rgb888" if rgb888 else "rgb555", seed, e.code, rc); bad += 1
            panics += 1
            continue
        got, got_z = fb.download()
        zsame = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        if not (rc == 0 and np.array_equal(got, want) and zsame and tm["triangles_drawn"] == otm["triangles_drawn"]):
            print("MISMATCH", "rgb888" if rgb888 else "rgb555", seed, nt); bad += 1
            continue
        ok += 1
        wire = (sc.settings.backface_cull and sc.settings.backface_wireframe) or sc.settings.wireframe_overlay
        if not rgb888 and not wire:
            mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
            for _ in range(3):
                mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
            got, got_z = fb.download()
            if not (np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))):
                print("MISMATCH (enqueued)", seed, nt); bad += 1
            mesh.free(); enq += 1
    print(f"{'rgb888' if rgb888 else 'rgb555'}: seeds {first}..{first + n - 1}: {ok} identical spot-lit frames (framebuffer + z-buffer + drawn count), "
          f"{panics} reference panics reported as the same error code, {enq} also enqueued x3")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
