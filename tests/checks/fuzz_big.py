"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Fuzz of the shapes the ordinary fuzz does not reach: thousands of
triangles in small framebuffers (crowded tiles: the scratch-slice path, over-full key buckets, the ordered pass's global sort)
and in 1280x720 / 1920x1080 ones (coarse mask tiles), both colour paths, blocking and (RGB555) enqueued.
usage (GPU box): python tests/checks/fuzz_big.py [n_scenes] [first_seed] [triangle counts, e.g. 1500,3000,4000]
(counts of at most 4 096 get no crowded-tile scratch: their crowded tiles take the windows-in-list-order route)"""
import dataclasses, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
first = int(sys.argv[2]) if len(sys.argv) > 2 else 700000
counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4500, 6000, 12000]
t0 = time.time(); bad = ok = panics = 0
for seed in range(first, first + n):
    rng = np.random.default_rng(seed ^ 0xB16)
    rgb888 = bool(rng.random() < 0.35)
    nt = int(rng.choice(counts))
    w, h = [(64, 64), (200, 150), (320, 240), (1280, 720), (1920, 1080)][int(rng.integers(0, 5))]
    sc = dataclasses.replace(fuzz.fuzz_scene(seed, rgb888, n_tris=nt), width=w, height=h)
    sc.settings.backface_wireframe = False; sc.settings.wireframe_overlay = False
    if rng.random() < 0.85:                                 # with thousands of triangles nearly every scene holds a non-finite vertex
        pos = sc.vertices["pos"]; pos[~np.isfinite(pos)] = np.float32(1.5)          # (= a reference panic): keep most scenes drawable
    if rng.random() < 0.5:                                  # pile the surfaces into the middle of the screen
        sc.vertices["pos"][:, :2] *= np.float32(rng.choice([0.05, 0.3]))
    want, want_z, otm, rc = (orc.render_scene888 if rgb888 else orc.render_scene)(sc)
    ctx = pkg.Context(0)                                    # fresh context: scratch buffers sized by this scene alone
    try:
        fb = pkg.Framebuffer(w, h, ctx); fb.clear(sc.clear)
        try:
            if rgb888: tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
            else: tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        except pkg.B32Error as e:
            if e.code != rc: print("MISMATCH (error code)", seed, e.code, rc); bad += 1
            panics += 1
            continue
        def same(got, got_z):
            zs = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
            return rc == 0 and np.array_equal(got, want) and zs
        if not same(*fb.download()) or tm["triangles_drawn"] != otm["triangles_drawn"]:
            print("MISMATCH", "rgb888" if rgb888 else "rgb555", seed, nt, (w, h)); bad += 1
            continue
        if not rgb888:
            ctx.set_textures(sc.textures)
            mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
            for _ in range(2):
                mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
            try:
                if not same(*fb.download()):
                    print("MISMATCH (enqueued)", seed, nt, (w, h)); bad += 1
                    continue
            except pkg.B32Error as e:                       # an enqueue-only frame may refuse a tile that needs the global sort scratch
                if e.code != pkg.abi.B32_ERR_UNSUPPORTED: raise
            mesh.free()
        ok += 1
    finally:
        ctx.close()
print(f"seeds {first}..{first + n - 1}: {ok} identical frames, {panics} reference panics, triangle counts {counts}, 64x64 .. 1920x1080")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
