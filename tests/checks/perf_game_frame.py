"""TEST INFRASTRUCTURE (uses the oracle as the checker; lives under tests/ for that reason).  The game's frame (src/game/renderer.rs:91-179: clear, skybox sphere + stars, the level's rooms, debug lines, read-back
for present) on the sample levels at the game's 640x480: device path (wall clock, one download per frame) against the CPU
oracle doing the same calls on one core.  GPU box only; fixtures from tests/golden (no reference tree needed)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry
pkg = entry.load_package()
entry.build_oracle()
from oracle import oracle as orc
from bonnie32_b200 import levels
import c3, cases

W, H = 640, 480
ctx = pkg.Context(0)
fb = pkg.Framebuffer(W, H, ctx)
for path in c3.scene_paths():
    sc = c3.load_scene(path)
    cam = sc.camera
    sv, sf = cases.sky_mesh(cam.position)
    stars = cases.star_list(cam, W, H, time=1.0, count=400)
    lines = cases.random_lines(W, H, 64, 7, kinds=(2,))
    lines["z0"] = lines["z1"] = 200.0
    lr = levels.LevelRenderer(ctx, sc)

    def frame():
        fb.clear((0, 0, 0))
        fb.render_skybox_mesh(sv, sf, cam)
        fb.render_stars(stars, cam, 2.0)
        lr.render(fb, cam, clear=False)
        fb.draw_lines(lines)
        return fb.download_view()[0]                       # pinned destination, as a game loop would keep one

    for _ in range(5): got = frame()
    t0 = time.perf_counter()
    for _ in range(50): got = frame()
    dev = (time.perf_counter() - t0) / 50

    def oracle_frame():
        rgba = np.zeros((H, W, 4), np.uint8); rgba[..., 3] = 255
        z = np.full((H, W), np.finfo(np.float32).max, np.float32)
        orc.render_skybox_mesh(rgba, sv, sf, cam)
        orc.render_stars(rgba, stars, cam, 2.0)
        for rc in sc.rooms:
            orc.render_mesh_15(rgba, z, rc.vertices, rc.faces, sc.textures, cam, sc.settings(rc.ambient), rc.fog)
        orc.draw_lines(rgba, z, lines)
        return rgba
    want = oracle_frame()
    t0 = time.perf_counter()
    for _ in range(3): oracle_frame()
    cpu = (time.perf_counter() - t0) / 3
    tris = sum(len(rc.faces) for rc in sc.rooms)
    print(f"{sc.name:12s} {len(sc.rooms)} rooms {tris:5d} triangles + {len(sf)} sky faces: device frame {dev * 1e6:7.1f} us ({1 / dev:6.0f} fps)   "
          f"oracle, 1 core {cpu * 1e3:6.2f} ms ({1 / cpu:4.0f} fps)   identical: {np.array_equal(got, want)}")
    lr.close()
