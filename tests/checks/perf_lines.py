"""TEST INFRASTRUCTURE (uses the oracle as the checker; lives under tests/ for that reason).  Time b32_draw_lines against the CPU oracle on the test line lists (blocking call, pageable host list)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry
pkg = entry.load_package()
entry.build_oracle()
from oracle import oracle as orc
import cases

ctx = pkg.Context(0)
for name, w, h, n, kinds in [("2D, all blend modes", 640, 480, 2000, (0,)), ("3D strict/overlay", 640, 480, 2000, (2, 3)), ("all kinds", 640, 480, 2000, (0, 1, 2, 3, 4)),
                             ("all kinds", 320, 240, 500, (0, 1, 2, 3, 4)), ("editor grid (3D alpha)", 640, 480, 200, (4,))]:
    lines = cases.random_lines(w, h, n, 5, kinds=kinds)
    rgba, z = cases.line_background(w, h, 3)
    fb = pkg.Framebuffer(w, h, ctx)
    fb.upload(rgba, z)
    for _ in range(3): fb.draw_lines(lines)
    l0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    for _ in range(20): fb.draw_lines(lines)
    dt = (time.perf_counter() - t0) / 20
    launches = (ctx.kernel_launches() - l0) / 20
    want = rgba.copy()
    t0 = time.perf_counter()
    for _ in range(5): orc.draw_lines(want, z, lines)
    ot = (time.perf_counter() - t0) / 5
    print(f"{name:24s} {w}x{h} {n:5d} lines: device call {dt * 1e6:8.1f} us ({launches:.0f} kernels)   oracle (1 core) {ot * 1e6:8.1f} us")
