"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Skybox fuzz: random sphere + mountain meshes around random cameras,
random star lists and sizes, b32_render_skybox_mesh + b32_render_stars vs the oracle's two passes.
usage (GPU box): python tests/checks/fuzz_sky.py [n_frames] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
first = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
ctx = pkg.Context(0)
t0 = time.time(); bad = 0
for seed in range(first, first + n):
    rng = np.random.default_rng(seed)
    w, h = [(320, 240), (640, 480), (200, 150), (333, 77), (64, 64)][int(rng.integers(0, 5))]
    pos = rng.normal(size=3) * np.array([200.0, 50.0, 200.0]) * float(rng.choice([0.0, 1.0, 30.0]))
    cam = cases._rotated_camera(float(rng.normal() * 0.8), float(rng.random() * 6.3), pos)
    sv, sf = cases.sky_mesh(cam.position, seed=seed, h_segments=int(rng.choice([12, 48])), v_segments=int(rng.choice([8, 32])),
                            n_mountains=int(rng.choice([0, 7, 40])))
    if rng.random() < 0.3:                                     # the mesh need not be centred on the camera: partly behind it
        sv["pos"] += (rng.normal(size=3) * 4000.0).astype(np.float32)
    stars = cases.star_list(cam, w, h, float(rng.random() * 100), seed=seed, count=int(rng.choice([0, 50, 400])),
                            horizon=float(rng.choice([0.3, 0.5, 0.9])), twinkle_speed=float(rng.choice([0.0, 1.5])))
    size = float(rng.choice([0.5, 1.0, 2.0, 3.0, 7.5]))
    fb = pkg.Framebuffer(w, h, ctx)
    clear = tuple(int(x) for x in rng.integers(0, 256, 3))
    fb.clear(clear)
    fb.render_skybox_mesh(sv, sf, cam)
    if len(stars): fb.render_stars(stars, cam, size)
    got, _ = fb.download()
    want = np.empty((h, w, 4), np.uint8); want[...] = np.array(list(clear) + [255], np.uint8)
    orc.render_skybox_mesh(want, sv, sf, cam)
    if len(stars): orc.render_stars(want, stars, cam, size)
    if not np.array_equal(got, want):
        print("MISMATCH seed", seed, (w, h), "pixels", int((got != want).any(-1).sum())); bad += 1
print(f"seeds {first}..{first + n - 1}: {n} skybox frames (sphere + mountains, stars)")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
