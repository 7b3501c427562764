"""TEST INFRASTRUCTURE (uses the oracle as the checker).  A long fuzz run of the CUDA path against the oracle, beyond the
60 + 60 scenes of tests/test_gpu_parity.py::test_fuzz_gpu_equals_oracle: random settings x adversarial geometry (tests/fuzz.py),
both colour paths, 30-1500 triangles per scene, every non-panicking scene also through the enqueued (CUDA-graph) path.
usage (GPU box): python tests/checks/fuzz_extended.py [n_scenes_per_path] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = pkg.Context(0)
t0 = time.time()
bad = 0
for rgb888 in (False, True):
    ok = panics = enq = nan_frames = 0
    for seed in range(first, first + n):
        nt = int(np.random.default_rng(seed ^ 0x5EED).choice([30, 120, 120, 400, 1500]))
        sc = fuzz.fuzz_scene(seed, rgb888, n_tris=nt)
        want, want_z, otm, rc = (orc.render_scene888 if rgb888 else orc.render_scene)(sc)
        fb = pkg.Framebuffer(sc.width, sc.height, ctx)
        fb.clear(sc.clear)
        before = fb.download()[0]
        try:
            if rgb888:
                tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
            else:
                tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
        except pkg.B32Error as e:
            if e.code != rc or not np.array_equal(fb.download()[0], before):
                print("MISMATCH (error path)", "rgb888" if rgb888 else "rgb555", seed, e.code, rc); bad += 1
            panics += 1
            continue
        got, got_z = fb.download()
        # z-buffer bit for bit, except the sign / payload of a NaN (IEEE leaves them to the implementation)
        zsame = ((got_z.view(np.uint32) == want_z.view(np.uint32)) | (np.isnan(got_z) & np.isnan(want_z))).all()
        nan_frames += int(np.isnan(want_z).any())
        same = rc == 0 and np.array_equal(got, want) and zsame and tm["triangles_drawn"] == otm["triangles_drawn"]
        if not same:
            print("MISMATCH", "rgb888" if rgb888 else "rgb555", seed, nt); bad += 1
            continue
        ok += 1
        wire = (sc.settings.backface_cull and sc.settings.backface_wireframe) or sc.settings.wireframe_overlay
        if not rgb888 and not wire:                        # the same frame enqueued (clear folded in, both passes, graph replay)
            mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
            for _ in range(3):
                mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
            got, got_z = fb.download()
            if not (np.array_equal(got, want) and np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))):
                print("MISMATCH (enqueued)", seed, nt); bad += 1
            mesh.free(); enq += 1
    print(f"{'rgb888' if rgb888 else 'rgb555'}: seeds {first}..{first + n - 1}: {ok} identical frames (framebuffer + z-buffer + drawn count), "
          f"{panics} reference panics reported as the same error code, {enq} also enqueued x3, {nan_frames} frames with NaN depths in the z-buffer")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
