"""Device time of the ordered replay (semi-transparent surfaces, x-ray) on float-projection scenes, where every surface
replays the reference's rounded edge additions.  usage (GPU box): python tools/perf_ordered_float.py
(B32_NO_EDGE_PREFIX=1 for the per-fragment replay)"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
import cases

ctx = pkg.Context(0)
by = {s.name: s for s in cases.feature_scenes(1000)}
for name in ("mixed_zbuffer", "mixed_painter", "xray"):
    for w, h in ((320, 240), (640, 480), (1920, 1080)):
        for fixed in (True, False):
            sc = cases._with(by[name], name, width=w, height=h, use_fixed_point=fixed)
            fb = pkg.Framebuffer(w, h, ctx)
            best = None
            for _ in range(5):
                fb.clear(sc.clear)
                tm = pkg.render_mesh_15(fb, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
                best = tm if best is None or tm["draw_ms"] < best["draw_ms"] else best
            print(f"{name:16s} {w:4d}x{h:<4d} {'fixed' if fixed else 'float'}  cull {best['cull_ms'] * 1e3:7.1f} us  draw (pass 1 + ordered replay) {best['draw_ms'] * 1e3:8.1f} us  drawn {best['triangles_drawn']}")
