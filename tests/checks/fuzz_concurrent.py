"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Concurrency fuzz: several contexts (streams) with frames of different
random scenes in flight at the same time, as bench.py and a multi-viewport editor have them; every context's framebuffer is
compared with the oracle afterwards.  usage (GPU box): python tests/checks/fuzz_concurrent.py [n_rounds] [first_seed] [n_contexts]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
first = int(sys.argv[2]) if len(sys.argv) > 2 else 900000
nctx = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctxs = [pkg.Context(0) for _ in range(nctx)]
t0 = time.time(); bad = frames = skipped = 0
for rnd in range(n):
    jobs = []
    for k, c in enumerate(ctxs):
        seed = first + rnd * nctx + k
        rng = np.random.default_rng(seed ^ 0xC0)
        sc = fuzz.fuzz_scene(seed, False, n_tris=int(rng.choice([120, 400, 1500, 5000])))
        sc.settings.backface_wireframe = False; sc.settings.wireframe_overlay = False
        pos = sc.vertices["pos"]; pos[~np.isfinite(pos)] = np.float32(1.5)
        want, want_z, otm, rc = orc.render_scene(sc)
        if rc != 0:
            skipped += 1
            continue
        fb = pkg.Framebuffer(sc.width, sc.height, c)
        c.set_textures(sc.textures)
        jobs.append((c, fb, pkg.Mesh(c, sc.vertices, sc.faces), sc, want, want_z))
    for rep in range(3):                                    # all contexts enqueue before anybody waits
        for c, fb, mesh, sc, _, _ in jobs:
            mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
    for c, fb, mesh, sc, want, want_z in jobs:
        try:
            got, got_z = fb.download()
        except pkg.B32Error as e:                           # an enqueue-only frame may refuse a tile that needs the global sort scratch
            if e.code != pkg.abi.B32_ERR_UNSUPPORTED: raise
            skipped += 1; mesh.free(); continue
        zs = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        frames += 1
        if not (np.array_equal(got, want) and zs):
            print("MISMATCH round", rnd, "scene", sc.name, (sc.width, sc.height)); bad += 1
        mesh.free()
print(f"rounds {n} x {nctx} contexts: {frames} frames compared ({skipped} skipped: reference panics / refused), 3 enqueues each, all contexts in flight together")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
