"""TEST INFRASTRUCTURE.  One fuzz scene rendered many times (blocking and enqueued): how often does it differ from the oracle?
usage: python tests/checks/fuzz_repeat.py rgb555|rgb888 seed n_tris [repeats]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz
rgb888 = sys.argv[1] == "rgb888"; seed = int(sys.argv[2]); nt = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
sc = fuzz.fuzz_scene(seed, rgb888, n_tris=nt)
ctx = pkg.Context(0)
want, want_z, otm, rc = (orc.render_scene888 if rgb888 else orc.render_scene)(sc)
fb = pkg.Framebuffer(sc.width, sc.height, ctx)
def cmp(got, got_z):
    zb = (got_z.view(np.uint32) != want_z.view(np.uint32)) & ~(np.isnan(got_z) & np.isnan(want_z))
    return int((got != want).any(-1).sum()), int(zb.sum())
res_b, res_e = [], []
for _ in range(reps):
    fb.clear(sc.clear)
    if rgb888: pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
    else: pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    res_b.append(cmp(*fb.download()))
if not rgb888:
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    for _ in range(reps):
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
        res_e.append(cmp(*fb.download()))
print("seed", seed, "size", sc.width, sc.height, "env", {k: v for k, v in os.environ.items() if k.startswith("B32_")})
print(" blocking:", sum(1 for r in res_b if r != (0, 0)), "of", reps, "differ", sorted(set(res_b))[:6])
print(" enqueued:", sum(1 for r in res_e if r != (0, 0)), "of", len(res_e), "differ", sorted(set(res_e))[:6])
