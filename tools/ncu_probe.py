"""Blocking calls on the shapes other than C4, for ncu launch lists / captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe.csv python tools/ncu_probe.py
    ncu --set full --clock-control none --import-source on -k regex:k_fill_ordered -c 2 -o gpurun_out/ord python tools/ncu_probe.py ordered"""
import copy
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
import cases
from bonnie32_b200 import scenes, abi
what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = pkg.Context(0)


def run(sc, w=None, h=None, n=2):
    fb = pkg.Framebuffer(w or sc.width, h or sc.height, ctx)
    ctx.set_textures(sc.textures)
    mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
    for _ in range(n):
        fb.clear(sc.clear); mesh.render(sc.camera, sc.settings, sc.fog)
    ctx.sync(); mesh.free()


sc = scenes.scene_c4()
t = copy.copy(sc); t.faces = sc.faces.copy(); t.faces["flags"] = abi.face_flags(0, abi.BLEND_AVERAGE, True, 255)
x = copy.copy(sc); x.settings = copy.copy(sc.settings); x.settings.xray_mode = True
if what in ("all", "ordered"):
    run(t); run(x)
if what == "all":
    run(scenes.scene_c1(), n=3); run(scenes.scene_c2(), n=3)
if what in ("all", "bigtri"):
    run(cases.big_triangle_scene(), 1920, 1080, n=2)
if what == "float":         # float-projection level-like scene: every surface stepped -> the PRE instantiations
    f = cases._with(cases.feature_scenes(1000)[0], "c2_float", use_fixed_point=False, width=640, height=480)
    run(f, 640, 480, n=2)
    run(cases.big_triangle_scene(), 1920, 1080, n=3)     # fixed point: call 1 per-pixel replay, calls 2, 3 the shared prefix
if what == "1m":
    run(scenes.scene_c4(n_tris=1_000_000), n=2)
