"""One blocking C4 frame (dense fill shape) and a few enqueued ones (sparse shape), for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:"k_setup|k_fill_opaque" -s 4 -c 4 -o gpurun_out/x python tools/ncu_c4.py"""
import sys
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
sc = pkg.scenes.scene_c4()
ctx = pkg.Context(0)
fb = pkg.Framebuffer(320, 240, ctx)
ctx.set_textures(sc.textures)
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(3):                       # blocking: k_setup, k_fill_opaque<dense>
    fb.clear(sc.clear); mesh.render(sc.camera, sc.settings)
for _ in range(3):                       # enqueued frames: k_setup (+ clear), k_fill_opaque<sparse>
    mesh.frame_enqueue(sc.clear, sc.camera, sc.settings, None)
ctx.sync()
