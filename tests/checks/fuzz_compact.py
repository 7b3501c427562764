"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Fuzz of the compact marshalling layouts of b32_render_mesh_15_ex
(B32_VTX_NO_NORMAL / B32_FACES_IMPLICIT / B32_FACES_UNIFORM): random soups, blocking and ASYNC, against the oracle on the
full records.  usage (GPU box): python tests/checks/fuzz_compact.py [n_scenes] [first_seed]"""
import ctypes as C, dataclasses, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
from bonnie32_b200 import abi
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
first = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
ctx = pkg.Context(0); lib = ctx.lib
t0 = time.time(); bad = ok = panics = 0
seen = {}
for seed in range(first, first + n):
    rng = np.random.default_rng(seed ^ 0xCC)
    sc = fuzz.fuzz_scene(seed, False, n_tris=int(rng.choice([30, 120, 400, 1500])))
    f = sc.faces.copy(); f["v"] = np.arange(len(f) * 3, dtype=np.uint32).reshape(-1, 3)          # an unindexed soup
    if rng.random() < 0.4: f["flags"][:] = f["flags"][0]                                          # ... with one flags word
    sc = dataclasses.replace(sc, faces=f)
    sc.settings.backface_wireframe = False; sc.settings.wireframe_overlay = False
    want, want_z, otm, rc = orc.render_scene(sc)
    v_, f_, flags = abi.compact_buffers(sc.vertices, sc.faces, sc.settings.shading == abi.SHADE_NONE)
    seen[flags] = seen.get(flags, 0) + 1
    fb = pkg.Framebuffer(sc.width, sc.height, ctx); ctx.set_textures(sc.textures)
    cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi(); fog = pkg.raster.fog_to_abi(sc.fog)
    for asyn in (0, abi.RENDER_ASYNC):
        fb.clear(sc.clear)
        code = lib.b32_render_mesh_15_ex(ctx.h, v_.ctypes.data, len(v_), f_.ctypes.data, len(sc.faces), C.byref(cam), C.byref(st),
                                         C.byref(fog) if fog is not None else None, flags | asyn, None)
        try:
            ctx.sync()                                      # errors of an enqueue-only call surface here
        except pkg.B32Error as e:
            code = code or e.code
        got, got_z = fb.download()
        if rc != 0 or code != 0:
            clear_px = np.array(list(sc.clear[:3]) + [255], np.uint8)
            if code != rc or not (got == clear_px).all(): print("MISMATCH (error path)", seed, asyn, code, rc); bad += 1
            panics += 1
            continue
        zs = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        if not (np.array_equal(got, want) and zs):
            print("MISMATCH seed", seed, "async" if asyn else "blocking", "flags", flags); bad += 1
        else: ok += 1
print(f"seeds {first}..{first + n - 1}: {ok} identical frames (blocking + async), {panics} reference panics; layouts used {seen}")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
