"""TEST INFRASTRUCTURE (uses the oracle as the checker).  Overlay-primitive fuzz: random lists of every b32_line kind (2D lines in
all blend modes, alpha lines, depth-tested lines, circles, alpha circles, filled rectangles, thick lines) over random
backgrounds, b32_draw_lines vs b32o_draw_lines.
usage (GPU box): python tests/checks/fuzz_lines.py [n_lists] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import oracle as orc
from bonnie32_b200 import abi
import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
first = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
ctx = pkg.Context(0)
t0 = time.time(); bad = prims = 0
for seed in range(first, first + n):
    rng = np.random.default_rng(seed)
    w, h = [(320, 240), (64, 48), (37, 23), (200, 150), (640, 480)][int(rng.integers(0, 5))]
    cnt = int(rng.choice([5, 40, 200, 600]))
    ln = cases.random_lines(w, h, cnt, seed, kinds=tuple(range(9)), spread=float(rng.choice([1.1, 1.5, 3.0])))
    circ = (ln["kind"] == abi.LINE_CIRCLE) | (ln["kind"] == abi.LINE_CIRCLE_ALPHA)
    ln["x1"] = np.where(circ, rng.integers(-2, 30, cnt), ln["x1"])                  # radius
    thick = ln["kind"] == abi.LINE_THICK
    ln["z0"] = np.where(thick, rng.integers(-1, 8, cnt).astype(np.float32), ln["z0"])   # thickness (<= 1: plain line)
    rgba, z = cases.line_background(w, h, seed)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    fb.draw_lines(ln)
    got, got_z = fb.download()
    want = rgba.copy()
    rc = orc.draw_lines(want, z, ln)
    prims += cnt
    if rc != 0 or not np.array_equal(got, want) or not np.array_equal(got_z.view(np.uint32), z.view(np.uint32)):
        print("MISMATCH seed", seed, "size", w, h, "prims", cnt, "pixels", int((got != want).any(-1).sum())); bad += 1
print(f"seeds {first}..{first + n - 1}: {n} lists, {prims} primitives of all nine kinds")
print(f"mismatches: {bad}   ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
