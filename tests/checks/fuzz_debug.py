"""TEST INFRASTRUCTURE.  One fuzz scene in detail: usage python tests/checks/fuzz_debug.py rgb888|rgb555 seed n_tris"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import fuzz
rgb888 = sys.argv[1] == "rgb888"; seed = int(sys.argv[2]); nt = int(sys.argv[3])
sc = fuzz.fuzz_scene(seed, rgb888, n_tris=nt)
ctx = pkg.Context(0)
want, want_z, otm, rc = (orc.render_scene888 if rgb888 else orc.render_scene)(sc)
fb = pkg.Framebuffer(sc.width, sc.height, ctx); fb.clear(sc.clear)
tm = pkg.render_mesh(fb, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings) if rgb888 else pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
got, got_z = fb.download()
s = sc.settings
print("seed", seed, "size", sc.width, sc.height, "rc", rc, "drawn", tm["triangles_drawn"], otm["triangles_drawn"])
print({k: getattr(s, k) for k in ("affine_textures", "use_zbuffer", "shading", "backface_cull", "backface_wireframe", "dithering", "wireframe_overlay", "use_fixed_point", "xray_mode")}, "ortho", s.ortho_projection is not None)
bad = (got != want).any(-1); badz = (got_z.view(np.uint32) != want_z.view(np.uint32)) & ~(np.isnan(got_z) & np.isnan(want_z))
print("pixels differ", int(bad.sum()), "z differ", int(badz.sum()))
ys, xs = np.nonzero(bad | badz)
for y, x in list(zip(ys, xs))[:12]:
    print((x, y), "got", got[y, x], got_z[y, x], "want", want[y, x], want_z[y, x])
fl = sc.faces["flags"]
print("blend modes", np.bincount((fl >> 16) & 7, minlength=6), "editor_alpha<255", int(((fl >> 24) < 255).sum()), "tex ids", np.unique(fl & 0xFFFF))
print("tex blends", [int(t.blend_mode) for t in (sc.textures8 if rgb888 else sc.textures)])

# ---- smallest prefix of the face list that still differs, then the last face of that prefix --------------------------
import dataclasses
def differs(m):
    s2 = dataclasses.replace(sc, faces=sc.faces[:m].copy())
    w, wz, _, rc2 = (orc.render_scene888 if rgb888 else orc.render_scene)(s2)
    fb.clear(sc.clear)
    if rgb888: pkg.render_mesh(fb, s2.vertices, s2.faces, s2.textures8, s2.camera, s2.settings)
    else: pkg.render_mesh_15(fb, s2.vertices, s2.faces, s2.textures, s2.camera, s2.settings, s2.fog)
    g_, gz_ = fb.download()
    zb = (gz_.view(np.uint32) != wz.view(np.uint32)) & ~(np.isnan(gz_) & np.isnan(wz))
    return (g_ != w).any() or zb.any()
if bad.any() or badz.any():
    lo, hi = 0, len(sc.faces)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if differs(mid): hi = mid
        else: lo = mid
    fi = hi - 1
    f = sc.faces[fi]
    print("first differing prefix:", hi, "faces; face", fi, "flags: tex", int(f["flags"]) & 0xFFFF, "blend", (int(f["flags"]) >> 16) & 7, "black_tr", (int(f["flags"]) >> 19) & 1, "alpha", int(f["flags"]) >> 24)
    for k in f["v"]:
        v = sc.vertices[k]; print("   v", k, v["pos"], v["uv"], v["normal"], v["rgba"])
    print("lights", [(int(l.type), l.position, l.direction, l.radius, l.intensity, l.color, l.enabled) for l in s.lights], "ambient", s.ambient)

# ---- the same frame enqueued (sparse fill shape, folded clear, ordered pass behind pass 1) ------------------------------
if not rgb888:
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    ctx.set_textures(sc.textures)
    for _ in range(3):
        mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, sc.fog)
    g2, gz2 = fb.download()
    b2 = (g2 != want).any(-1); bz2 = (gz2.view(np.uint32) != want_z.view(np.uint32)) & ~(np.isnan(gz2) & np.isnan(want_z))
    print("ENQUEUED: pixels differ", int(b2.sum()), "z differ", int(bz2.sum()))
    ys, xs = np.nonzero(b2 | bz2)
    for y, x in list(zip(ys, xs))[:8]:
        print("  ", (x, y), "got", g2[y, x], gz2[y, x], "want", want[y, x], want_z[y, x])
