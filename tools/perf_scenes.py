#!/usr/bin/env python
"""Device time of the resident render call on scenes other than the headline C4 frame (GPU box only).
Prints one line per scene: kernel-group times from b32_debug_kernel_times and the blocking call's wall time."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
import c3, cases
from bonnie32_b200 import scenes, abi


ONLY = sys.argv[1:]          # optional name filters: run only the scenes whose name contains one of them


def run(ctx, name, sc_list, w, h, clear, reps=20):
    """sc_list: list of (vertices, faces, camera, settings, fog) calls composing one frame."""
    if ONLY and not any(o in name for o in ONLY):
        return
    fb = pkg.Framebuffer(w, h, ctx)
    meshes = [pkg.Mesh(ctx, v, f) for v, f, *_ in sc_list]
    times = np.zeros(4); wall = []
    kt = (C.c_float * 7)()
    tris = sum(len(f) for _, f, *_ in sc_list)
    drawn = 0
    for r in range(reps + 3):
        fb.clear(clear); ctx.sync()
        t0 = time.perf_counter()
        d = 0
        for m, (_, _, cam, st, fog) in zip(meshes, sc_list):
            tm = m.render(cam, st, fog); d += tm["triangles_drawn"]
            if r >= 3:
                ctx.lib.b32_debug_kernel_times(ctx.h, kt, 7)
                times += np.array(list(kt)[:4])
        if r >= 3:
            wall.append(time.perf_counter() - t0)
        drawn = d
    times /= reps
    print(f"{name:42s} {w}x{h} calls={len(sc_list):2d} tris={tris:8d} drawn={drawn:7d} setup+bin {times[0]*1e3:7.1f} us  fill {times[1]*1e3:7.1f} us  "
          f"obin {times[2]*1e3:6.1f} us  ofill {times[3]*1e3:7.1f} us  wall/frame {np.median(wall)*1e6:8.1f} us  {tris/np.median(wall)/1e6:8.1f} Mtri/s")
    for m in meshes:
        m.free()


def main():
    ctx = pkg.Context(0)
    def one(sc): return [(sc.vertices, sc.faces, sc.camera, sc.settings, sc.fog)]
    for sc in [scenes.scene_c1(), scenes.scene_c2(), scenes.scene_c2(use_zbuffer=True), scenes.scene_c4(), scenes.scene_c4(use_zbuffer=True)]:
        ctx.set_textures(sc.textures)
        run(ctx, sc.name + ("_z" if sc.settings.use_zbuffer else ""), one(sc), sc.width, sc.height, sc.clear)
    sc = scenes.scene_c4(); ctx.set_textures(sc.textures)
    run(ctx, "c4_640x480", one(sc), 640, 480, sc.clear)
    run(ctx, "c4_1920x1080", one(sc), 1920, 1080, sc.clear)
    sc = scenes.scene_c4(n_tris=1_000_000); ctx.set_textures(sc.textures)
    run(ctx, "c4_1M_tris", one(sc), 320, 240, sc.clear, reps=5)
    sc = scenes.scene_c4(n_tris=1_000_000, use_zbuffer=True)
    run(ctx, "c4_1M_tris_z", one(sc), 320, 240, sc.clear, reps=5)
    # ordered replay at scale: x-ray (every surface in draw order), all faces semi-transparent (pass 2 only)
    import copy
    sc = scenes.scene_c4(); ctx.set_textures(sc.textures)
    x = copy.copy(sc); x.settings = copy.copy(sc.settings); x.settings.xray_mode = True
    run(ctx, "c4_xray", one(x), 320, 240, sc.clear, reps=5)
    t = copy.copy(sc); t.faces = sc.faces.copy(); t.faces["flags"] = abi.face_flags(0, abi.BLEND_AVERAGE, True, 255)
    run(ctx, "c4_all_transparent_painter", one(t), 320, 240, sc.clear, reps=5)
    t2 = copy.copy(t); t2.settings = copy.copy(sc.settings); t2.settings.use_zbuffer = True
    run(ctx, "c4_all_transparent_zbuffer", one(t2), 320, 240, sc.clear, reps=5)
    sc10 = scenes.scene_c4(n_tris=10_000); ctx.set_textures(sc10.textures)
    t3 = copy.copy(sc10); t3.faces = sc10.faces.copy(); t3.faces["flags"] = abi.face_flags(0, abi.BLEND_ADD, True, 255)
    run(ctx, "c4_10k_all_transparent", one(t3), 320, 240, sc.clear, reps=5)
    sc = cases.big_triangle_scene(); ctx.set_textures(sc.textures)
    run(ctx, "big_triangles", one(sc), 320, 240, sc.clear)
    run(ctx, "big_triangles_1920x1080", one(sc), 1920, 1080, sc.clear)
    by = {s.name: s for s in cases.feature_scenes(1000)}
    for n in ("mixed_painter", "mixed_zbuffer", "xray", "gouraud_lights"):
        sc = by[n]; ctx.set_textures(sc.textures)
        run(ctx, "feat1000_" + n, one(sc), sc.width, sc.height, sc.clear)
    for p in c3.scene_paths():
        lv = c3.load_scene(p); ctx.set_textures(lv.textures)
        for mode, kw in c3.MODES.items():
            calls = [(rc.vertices, rc.faces, lv.camera, lv.settings(rc.ambient, **kw), rc.fog) for rc in lv.rooms]
            run(ctx, f"c3_{lv.name}_{mode}", calls, lv.width, lv.height, lv.clear)
            if mode == "zbuffer":
                run(ctx, f"c3_{lv.name}_{mode}_640x480", calls, 640, 480, lv.clear)
                # float projection: no surface has integer edge values, every inside test replays the rounded additions
                fcalls = [(rc.vertices, rc.faces, lv.camera, lv.settings(rc.ambient, use_fixed_point=False, **kw), rc.fog) for rc in lv.rooms]
                run(ctx, f"c3_{lv.name}_{mode}_float", fcalls, lv.width, lv.height, lv.clear)
                run(ctx, f"c3_{lv.name}_{mode}_float_640x480", fcalls, 640, 480, lv.clear)


if __name__ == "__main__":
    main()
